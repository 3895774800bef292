"""Recipe: make the UNMODIFIED reference runnable on the GPU box as the checker / CPU baseline.

Test infrastructure only. The reference is pure Python, so "building" it is staging its own files where they can
travel: `/root/reference/backend/wavecapsdr/**/*.py` and `/root/reference/backend/benchmark_dsp.py` are copied
byte for byte into `oracle/_ref/backend/` — git-ignored (no reference source ever enters the history), NOT
gpurun-ignored (so the copy ships with the snapshot like the built .so). `__graft_entry__.build()` runs this when
`/root/reference` is present (the build container); on the GPU box only the staged copy is used:

  * tests/test_reference_benchmark_gpu.py runs the reference's own backend/benchmark_dsp.py against
    wavecap_sdr_b200.install();
  * bench.py --impl reference / cpu_baseline time the reference's own PolyphaseChannelizer.process + quadrature_demod
    on the box's host cores (`cpu_baseline.kind: "reference"`); without the copy they fall back to the oracle port.

Nothing under wavecap-sdr_b200/ may import from here.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/backend"
DST = os.path.join(ROOT, "oracle", "_ref", "backend")


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "wavecapsdr", "dsp", "channelizer.py"))


def build(verbose: bool = False) -> bool:
    """Stage the reference when /root/reference exists; returns whether a staged copy is available afterwards."""
    if not os.path.isdir(os.path.join(SRC, "wavecapsdr")):
        return staged()
    n = 0
    for base, _dirs, files in os.walk(os.path.join(SRC, "wavecapsdr")):
        if "__pycache__" in base:
            continue
        rel = os.path.relpath(base, SRC)
        for f in files:
            if not f.endswith((".py", ".json", ".yaml", ".yml")):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), os.path.join(DST, rel, f))
            n += 1
    shutil.copyfile(os.path.join(SRC, "benchmark_dsp.py"), os.path.join(DST, "benchmark_dsp.py"))
    if verbose:
        print(f"staged {n + 1} reference files under {DST}")
    return True


def load():
    """Import the staged reference (SURVEY §8c recipe: trunking before capture). Returns the wavecapsdr package."""
    import logging

    if not staged():
        raise RuntimeError("oracle/_ref is not staged (run oracle/build_ref.py in the build container)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.dont_write_bytecode = True
    if DST not in sys.path:
        sys.path.insert(0, DST)
    logging.disable(logging.CRITICAL)
    import wavecapsdr.trunking  # noqa: F401
    import wavecapsdr.capture  # noqa: F401
    return sys.modules["wavecapsdr"]


if __name__ == "__main__":
    print("staged" if build(verbose=True) else "reference not available")
