"""Build libwcsdr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each .cu is compiled to an object in parallel, then linked into
`wavecap-sdr_b200/libwcsdr_b200.so` (git-ignored, but it travels with the gpurun snapshot).
Objects are rebuilt only when the source or a header is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
ROOT = PKG.parent
LIB = PKG / "libwcsdr_b200.so"
OBJ = CSRC / "build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"] + (["-DWC_DEV_ABLATE", "-DWC_DEV"] if os.environ.get("WC_DEV_ABLATE") else []) \
        + (["-DWC_DEV"] if os.environ.get("WC_DEV") else [])   # dev build: WC_* env switches + superseded kernel variants
# per-file extras: the P25 kernels replay the reference's float32/float64 operation order, so the
# compiler must not contract a*b+c into one fused multiply-add there (explicit fma() calls stay fused)
EXTRA = {"p25.cu": ["-fmad=false"], "cqpsk.cu": ["-fmad=false"], "discdemod.cu": ["-fmad=false"]}


def _newer(a: Path, b: Path) -> bool:
    return (not b.exists()) or a.stat().st_mtime > b.stat().st_mtime


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + list((ROOT / "include").glob("*.h"))
    newest_hdr = max((h.stat().st_mtime for h in hdrs), default=0.0)

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = OBJ / (src.stem + ".o")
        if force or _newer(src, obj) or obj.stat().st_mtime < newest_hdr:
            cmd = [NVCC, *ARCH, *FLAGS, *EXTRA.get(src.name, []), "-c", str(src), "-o", str(obj)]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{p.stdout}\n{p.stderr}")
            (OBJ / (src.stem + ".ptxas.log")).write_text(p.stderr)
            return obj, p.stderr
        return obj, ""

    with ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    if force or any(_newer(o, LIB) for o in objs):
        # shared cudart: one CUDA runtime per process (torch's, loaded first by _native.lib()) instead of a second,
        # statically embedded copy; the rpath covers a process that loads the library without torch
        cmd = [NVCC, *ARCH, "-shared", "--cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
               "-o", str(LIB), *map(str, objs)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
