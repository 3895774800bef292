"""GPU: the reference's OWN unit tests of this path, unmodified, against wavecap_sdr_b200.install().

SURVEY §4: the reference's hot-path tests are property and known-answer tests (soft-clip bounds, RMS target, quadrature of a
tone, resampler lengths, AGC direction, wbfm/nbfm dtype / range / length, de-emphasis and MPX attenuation, FFT backend
registry and tone peaks, wire packers, C4FMDemodulator construction / empty input / reset, BCH / trellis / TSBK chains,
framer assertions). They travel in oracle/_ref/reference_backend.tar (packed byte for byte by oracle/build_ref.py in the
build container, git-ignored) and run here in a subprocess whose pytest plugin (tests/ref_install_plugin.py) calls install()
before they are imported — so `from wavecapsdr.dsp.fm import wbfm_demod` in those files resolves to the CUDA path. The same
files are first run with the reference untouched: whatever passes there must pass here."""
import os
import re
import subprocess
import sys

import pytest

from conftest import parity_note
from oracle import build_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref.staged(), reason="oracle/_ref not staged")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(paths, install: bool):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, env.get("PYTHONPATH", "")])
    env["WC_REF_NO_INSTALL"] = "0" if install else "1"
    work = os.path.dirname(os.path.dirname(paths[0]))
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_install_plugin", *paths, "-q", "-p", "no:cacheprovider", "-c", os.devnull,
           "--rootdir", work, "-rfE", "--tb=short", "-m", "not hardware", "-o", "addopts="]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=work, timeout=900)
    out = r.stdout + r.stderr
    summary = [ln for ln in r.stdout.splitlines() if re.search(r"\d+ (passed|failed|error)", ln)]
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|errors?|skipped|xfailed|xpassed)", summary[-1] if summary else "")}
    failed = sorted(set(re.findall(r"^(?:FAILED|ERROR) (\S+)", out, flags=re.M)))
    return r.returncode, counts, failed, out


def _check(paths, label, floor):
    rc0, base, failed0, out0 = _run(paths, install=False)
    assert base.get("passed", 0) >= floor, out0[-3000:]
    rc1, got, failed1, out1 = _run(paths, install=True)
    assert "reference names rebound" in out1, out1[-3000:]
    new_failures = [f for f in failed1 if f not in failed0]
    if new_failures:   # a timing assertion of the reference's suite on a busy box: one more try for exactly those tests
        work = os.path.dirname(os.path.dirname(paths[0]))
        _, _, still, out2 = _run([os.path.join(work, f) for f in new_failures], install=True)
        out1 += "\n--- retry ---\n" + out2
        new_failures = [f for f in new_failures if any(f.endswith(x) or x.endswith(f) for x in still)]
    parity_note(f"reference {label} ({len(paths)} files): untouched {base}, against install() {got}")
    assert not new_failures, "\n".join(new_failures) + "\n" + out1[-6000:]
    assert got.get("passed", 0) + len(failed1) >= base.get("passed", 0), (base, got)


def test_reference_unit_tests_pass_against_install(native):
    """the 11 files that call rebound functions directly"""
    paths = build_ref.reference_test_paths(only_the_path=True)
    if not paths:
        pytest.skip("the archive was packed without the reference tests")
    _check(paths, "unit tests of the path", 100)


def test_whole_reference_suite_passes_against_install(native):
    """Every test file of backend/tests/ and backend/tests/unit/ that can be collected in this image (all but the four that
    import the FastAPI app, whose `slowapi` dependency is absent, and the hardware-marked integration directory): trunking
    configuration and workers, validation, decoders, FEC, packers, device detection ... — with install() active nothing that
    passed before may fail."""
    paths = build_ref.reference_test_paths()
    if len(paths) < 30:
        pytest.skip("the archive holds only the path's test files")
    _check(paths, "test suite", 500)
