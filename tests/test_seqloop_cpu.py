"""Host check of the straight-line symbol-clock walkers (wavecap-sdr_b200/csrc/seqloop.cuh): compiled with g++ against
float32 host arithmetic, they must consume the same samples, fire at the same sample and leave the same clock bits as
the reference's per-sample `clock += symbol_time; if clock > 1` loop (decoders/p25.py:1255-1261, :520-527) — including
NaN / negative / tiny steps and every tile-end `room`."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_clock_walkers_match_the_per_sample_loop(tmp_path):
    exe = tmp_path / "seqloop_check"
    hdr = os.path.join(ROOT, "wavecap-sdr_b200", "csrc", "seqloop.cuh")
    subprocess.run(["g++", "-O1", "-x", "c++", os.path.join(ROOT, "tools", "seqloop_host_check.cpp"),
                    f'-DSEQLOOP_HEADER="{hdr}"', "-DITERS=400000", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches: 0" in r.stdout
