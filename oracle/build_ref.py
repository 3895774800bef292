"""Recipe: make the UNMODIFIED reference runnable on the GPU box as the checker / CPU baseline.

Test infrastructure only. The reference is pure Python, so "building" it is packing its own files where they can
travel: `/root/reference/backend/wavecapsdr/**/*.py` (+ its json/yaml data files) and
`/root/reference/backend/benchmark_dsp.py` go, byte for byte, into ONE archive `oracle/_ref/reference_backend.tar` —
git-ignored (no reference source ever enters the history or sits in the tree as source files), NOT gpurun-ignored (the
archive ships with the snapshot like the built .so). `__graft_entry__.build()` runs this when `/root/reference` is present
(the build container). At run time `load()` unpacks the archive into a scratch directory under the system temp dir
(once per archive version; real files, so numba's `cache=True` kernels work) and puts it on `sys.path`:

  * tests/test_reference_benchmark_gpu.py runs the reference's own backend/benchmark_dsp.py against
    wavecap_sdr_b200.install() and compares the rebound functions with the original ones executed on the box's CPU;
  * bench.py --impl reference / cpu_baseline time the reference's own functions on the box's host cores
    (`cpu_baseline.kind: "reference"`); without the archive they fall back to the oracle port.

Nothing under wavecap-sdr_b200/ may import from here.
"""
from __future__ import annotations

import hashlib
import io
import os
import sys
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/backend"
ARCHIVE = os.path.join(ROOT, "oracle", "_ref", "reference_backend.tar")
_loaded_dir: str | None = None
# The reference's own test suite (backend/tests/test_*.py and tests/unit/test_*.py), run unmodified against
# install() by tests/test_reference_unit_tests_gpu.py. Left out: the hardware-marked integration directory and the four
# files that import the FastAPI app (its `slowapi` dependency is not in this image, so they cannot be collected either way).
REFERENCE_TESTS_EXCLUDED = ("tests/test_captures_channels.py", "tests/test_config_reload.py", "tests/test_trunking_api.py",
                            "tests/test_trunking_voice_api.py")
# the subset that exercises functions install() rebinds directly (kept as a list for the report)
REFERENCE_TESTS_OF_THE_PATH = (
    "tests/unit/test_dsp_core.py", "tests/unit/test_fm_demod.py", "tests/unit/test_fft_backends.py",
    "tests/unit/test_pack_functions.py", "tests/test_p25_dsp.py", "tests/test_p25_bch.py", "tests/test_reference_fec.py",
    "tests/test_tsbk_chain.py", "tests/test_tsbk_decoding.py", "tests/test_tsbk_decoder_roundtrip.py",
    "tests/test_p25_message_assertions.py",
)


def _reference_tests() -> tuple[str, ...]:
    import glob

    found = sorted(glob.glob(os.path.join(SRC, "tests", "test_*.py")) + glob.glob(os.path.join(SRC, "tests", "unit", "test_*.py")))
    rel = tuple(os.path.relpath(p, SRC) for p in found)
    return tuple(r for r in rel if r not in REFERENCE_TESTS_EXCLUDED)


def staged() -> bool:
    return os.path.isfile(ARCHIVE)


def build(verbose: bool = False) -> bool:
    """Pack the reference when /root/reference exists; returns whether an archive is available afterwards."""
    if not os.path.isdir(os.path.join(SRC, "wavecapsdr")):
        return staged()
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    members = []
    for base, dirs, files in os.walk(os.path.join(SRC, "wavecapsdr")):
        dirs[:] = sorted(d for d in dirs if d != "__pycache__")
        for f in sorted(files):
            if f.endswith((".py", ".json", ".yaml", ".yml")):
                members.append(os.path.join(base, f))
    members.append(os.path.join(SRC, "benchmark_dsp.py"))
    # the reference's own unit tests of this path (property and known-answer tests): run unmodified against
    # install() by tests/test_reference_unit_tests_gpu.py
    for rel in _reference_tests() + ("tests/conftest.py", "tests/unit/__init__.py"):
        path = os.path.join(SRC, rel)
        if os.path.isfile(path):
            members.append(path)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as tar:        # deterministic: sorted members, zeroed metadata
        for path in members:
            info = tar.gettarinfo(path, arcname=os.path.join("backend", os.path.relpath(path, SRC)))
            info.mtime, info.uid, info.gid, info.uname, info.gname = 0, 0, 0, "", ""
            with open(path, "rb") as fh:
                tar.addfile(info, fh)
    data = buf.getvalue()
    if not (os.path.isfile(ARCHIVE) and open(ARCHIVE, "rb").read() == data):
        with open(ARCHIVE, "wb") as fh:
            fh.write(data)
    if verbose:
        print(f"packed {len(members)} reference files into {ARCHIVE} ({len(data)} bytes)")
    return True


def unpacked_dir() -> str:
    """<tmp>/wcsdr_b200_ref_<digest>/backend, unpacking the archive there on first use (atomic rename: safe under many
    concurrent worker processes)."""
    global _loaded_dir
    if _loaded_dir and os.path.isdir(_loaded_dir):
        return _loaded_dir
    if not staged():
        raise RuntimeError("oracle/_ref/reference_backend.tar is missing (run oracle/build_ref.py in the build container)")
    with open(ARCHIVE, "rb") as fh:
        digest = hashlib.sha1(fh.read()).hexdigest()[:16]
    final = os.path.join(tempfile.gettempdir(), f"wcsdr_b200_ref_{digest}")
    if not os.path.isdir(os.path.join(final, "backend", "wavecapsdr")):
        tmp = tempfile.mkdtemp(prefix="wcsdr_b200_ref_unpack_")
        with tarfile.open(ARCHIVE) as tar:
            tar.extractall(tmp, filter="data")
        try:
            os.rename(tmp, final)
        except OSError:                                      # another process won the race
            import shutil

            shutil.rmtree(tmp, ignore_errors=True)
    _loaded_dir = os.path.join(final, "backend")
    return _loaded_dir


def benchmark_script() -> str:
    return os.path.join(unpacked_dir(), "benchmark_dsp.py")


def reference_test_paths(only_the_path: bool = False) -> list[str]:
    """the packed reference test modules (all of them, or the subset that calls rebound functions directly)"""
    d = unpacked_dir()
    if only_the_path:
        rels = REFERENCE_TESTS_OF_THE_PATH
    else:
        import glob

        rels = sorted(os.path.relpath(p, d) for p in glob.glob(os.path.join(d, "tests", "test_*.py")) +
                      glob.glob(os.path.join(d, "tests", "unit", "test_*.py")))
    return [p for p in (os.path.join(d, rel) for rel in rels) if os.path.isfile(p)]


def load():
    """Import the packed reference (SURVEY §8c recipe: trunking before capture). Returns the wavecapsdr package."""
    import logging

    d = unpacked_dir()
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.dont_write_bytecode = True
    if d not in sys.path:
        sys.path.insert(0, d)
    logging.disable(logging.CRITICAL)
    import wavecapsdr.trunking  # noqa: F401
    import wavecapsdr.capture  # noqa: F401
    return sys.modules["wavecapsdr"]


if __name__ == "__main__":
    print("packed" if build(verbose=True) else "reference not available")
