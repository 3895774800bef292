"""CPU: oracle/spectrum.py pinned bit-exact to the reference's ScipyFFTBackend outputs."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import spectrum as osp


@pytest.fixture(scope="module")
def g():
    return np.load(golden_path("spectrum.npz"))


def test_c3_frames_bit_exact(g):
    fs = 61_440_000
    for i in range(4):
        p, f, b = osp.execute(osp.synth_c3(seed=3, n=65536, t0=i * 65536), fs, 65536)
        assert np.array_equal(p[::16], g["c3_frames_dec16"][i])
        if i == 0:
            assert np.array_equal(p, g["c3_frame0"])
    assert np.array_equal(f[::16], g["c3_freqs_dec16"]) and b == float(g["c3_bin_hz"])
    assert np.array_equal(osp.hann_window(65536)[::16], g["c3_window_dec16"])


def test_small_and_short_inputs(g):
    p, f, _ = osp.execute(osp.synth_c3(seed=4, n=5000, fs=2_400_000), 2_400_000, 2048)
    assert np.array_equal(p, g["s2048_power"]) and np.array_equal(f, g["s2048_freqs"])
    p, f, _ = osp.execute(osp.synth_c3(seed=5, n=100), 48000, 512)
    assert np.array_equal(p, g["short_power"]) and not p.any() and not f.any()


def test_reference_property_tests_hold():
    # tests/unit/test_fft_backends.py:42-75: window ends are 0, a 1 kHz tone peaks within 50 Hz
    w = osp.hann_window(1024)
    assert w[0] == 0 and w[-1] == 0 and abs(w[512] - 1) < 1e-3
    fs, n = 48000, 4096
    x = np.exp(2j * np.pi * 1000 * np.arange(n) / fs).astype(np.complex64)
    p, f, _ = osp.execute(x, fs, n)
    assert abs(f[np.argmax(p)] - 1000) < 50
    assert np.allclose(osp.averaged(np.stack([p, p + 2]), 2)[0], p + 1, atol=1e-5)
