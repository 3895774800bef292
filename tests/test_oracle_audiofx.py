"""CPU: oracle restatements of dsp/filters.py noise_blanker / spectral_noise_reduction vs the reference goldens."""
import numpy as np

from conftest import golden_path
from oracle import analog as oa


def test_blanker_and_nr_match_golden():
    g = np.load(golden_path("audiofx.npz"))
    x = g["x"]
    assert np.array_equal(oa.noise_blanker(x, 10.0, 3), g["nb"])
    assert np.array_equal(oa.noise_blanker(x[:4001], 6.0, 0), g["nb_w0"])
    for got, exp in ((oa.spectral_noise_reduction(x, 48000, 12.0), g["nr"]),
                     (oa.spectral_noise_reduction(x[:3000], 48000, 18.0), g["nr18"])):
        assert got.shape == exp.shape and np.max(np.abs(got - exp)) < 1e-6
    assert g["nr"].shape == (19968,) and g["nr18"].shape == (2560,)   # only whole STFT frames are returned


def test_am_ssb_with_noise_blanker_match_golden():
    """dsp/am.py:100-101, 213-215: the blanker runs on the real envelope / real part ahead of the filters."""
    from oracle.make_golden import am_blanker_input

    g = np.load(golden_path("am_blanker.npz"))
    x = am_blanker_input()
    assert np.array_equal(oa.am_demod(x, 48000, 16000, enable_noise_blanker=True, noise_blanker_threshold_db=8.0), g["am_nb"])
    assert np.array_equal(oa.am_demod(x, 48000, 48000, enable_agc=False, enable_noise_blanker=True), g["am_nb_noagc"])
    assert np.array_equal(oa.ssb_demod(x, 48000, 16000, mode="lsb", enable_noise_blanker=True, noise_blanker_threshold_db=6.0),
                          g["ssb_nb"])
