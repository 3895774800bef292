"""GPU parity (C ABI): noise blanker (exact) and spectral noise reduction (<= 1e-4 relative RMS) vs oracle + goldens."""
import numpy as np
import pytest

from conftest import golden_path, rel_rms
from oracle import analog as oa

pytestmark = pytest.mark.gpu


def test_noise_blanker_exact(native):
    from wavecap_sdr_b200.dsp.filters import noise_blanker

    g = np.load(golden_path("audiofx.npz"))
    x = g["x"]
    assert np.array_equal(noise_blanker(x, 10.0, 3), g["nb"])
    assert np.array_equal(noise_blanker(x[:4001], 6.0, 0), g["nb_w0"])
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 1000, 100_001):
        y = (rng.standard_normal(n) * 0.2).astype(np.float32)
        if n > 10:
            y[rng.integers(0, n, 5)] *= 40
        for thr, w in ((10.0, 3), (3.0, 1), (40.0, 5)):
            assert np.array_equal(noise_blanker(y, thr, w), oa.noise_blanker(y, thr, w)), (n, thr, w)
    z = np.zeros(100, np.float32)
    assert np.array_equal(noise_blanker(z), z) and noise_blanker(np.zeros(0, np.float32)).size == 0


def test_spectral_noise_reduction(native):
    from wavecap_sdr_b200.dsp.filters import spectral_noise_reduction

    g = np.load(golden_path("audiofx.npz"))
    x = g["x"]
    for got, exp in ((spectral_noise_reduction(x, 48000, 12.0), g["nr"]), (spectral_noise_reduction(x[:3000], 48000, 18.0), g["nr18"])):
        assert got.shape == exp.shape and got.dtype == np.float32
        assert rel_rms(got, exp) < 1e-4
    short = x[:500]
    assert np.array_equal(spectral_noise_reduction(short, 48000), short)


def test_wbfm_chain_with_both_flags(native):
    from wavecap_sdr_b200.capture import freq_shift
    from wavecap_sdr_b200.dsp.fm import wbfm_demod

    g = np.load(golden_path("audiofx.npz"))
    iq = freq_shift(oa.synth_c1(seed=1, n=120_000), 200000.0, 2_400_000)
    got = wbfm_demod(iq, 2_400_000, 48000, enable_noise_blanker=True, enable_noise_reduction=True)
    exp = g["wbfm_nb_nr"]
    assert got.shape == exp.shape and rel_rms(got, exp) < 1e-4


def test_am_ssb_demod_with_noise_blanker(native):
    """am_demod / ssb_demod(enable_noise_blanker=True) against outputs of the reference itself (dsp/am.py:100-101, 213-215)."""
    from oracle.make_golden import am_blanker_input
    from wavecap_sdr_b200.dsp.am import am_demod, ssb_demod

    g = np.load(golden_path("am_blanker.npz"))
    x = am_blanker_input()
    cases = ((am_demod(x, 48000, 16000, enable_noise_blanker=True, noise_blanker_threshold_db=8.0), g["am_nb"]),
             (am_demod(x, 48000, 48000, enable_agc=False, enable_noise_blanker=True), g["am_nb_noagc"]),
             (ssb_demod(x, 48000, 16000, mode="lsb", enable_noise_blanker=True, noise_blanker_threshold_db=6.0), g["ssb_nb"]))
    for got, exp in cases:
        assert got.shape == exp.shape and got.dtype == np.float32
        assert rel_rms(got, exp) < 1e-4
