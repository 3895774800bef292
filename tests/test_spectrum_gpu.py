"""GPU parity: csrc/spectrum.cu (FFTBackend "cuda") vs the oracle / reference goldens.
Tolerance: 1e-4 relative RMS on the dB spectrum (north_star)."""
import numpy as np
import pytest

from conftest import rel_rms, golden_path
from oracle import spectrum as osp

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def g():
    return np.load(golden_path("spectrum.npz"))


def test_registry_surface(native):
    from wavecap_sdr_b200.dsp.fft import FFTBackend, available_backends, get_backend

    assert available_backends() == ["cuda"]
    be = get_backend("auto", fft_size=65536)
    assert isinstance(be, FFTBackend) and be.name == "cuda" and be.fft_size == 65536
    with pytest.raises(ValueError):
        get_backend("fftw", fft_size=1024)
    assert np.abs(be.window - osp.hann_window(65536)).max() < 1e-7


def test_c3_65536_against_golden(native, g):
    from wavecap_sdr_b200.dsp.fft import get_backend

    be = get_backend("cuda", fft_size=65536)
    fs = 61_440_000
    r = be.execute(osp.synth_c3(seed=3, n=65536), fs)
    assert r.power_db.dtype == np.float32 and r.power_db.shape == (65536,)
    assert rel_rms(r.power_db, g["c3_frame0"]) < TOL
    assert np.argmax(r.power_db) == np.argmax(g["c3_frame0"])
    assert np.array_equal(r.freqs[::16], g["c3_freqs_dec16"]) and r.bin_hz == float(g["c3_bin_hz"])


def test_c3_frames_and_k4_average(native, g):
    from wavecap_sdr_b200.dsp.fft import get_backend

    be = get_backend("cuda", fft_size=65536)
    chunk = 3_072_000 // 16  # a chunk longer than the frame: only its first 65536 samples are used
    x = np.zeros(4 * chunk, np.complex64)
    for i in range(4):
        x[i * chunk:i * chunk + 65536] = osp.synth_c3(seed=3, n=65536, t0=i * 65536)
        x[i * chunk + 65536:(i + 1) * chunk] = 7.0  # must be ignored
    frames = be.execute_frames(x, n_frames=4, frame_stride=chunk, avg=1)
    for i in range(4):
        assert rel_rms(frames[i][::16], g["c3_frames_dec16"][i]) < TOL
    avg = be.execute_frames(x, n_frames=4, frame_stride=chunk, avg=4)
    exp = osp.averaged(np.stack([osp.execute(x[i * chunk:], 61_440_000, 65536)[0] for i in range(4)]), 4)
    assert avg.shape == (1, 65536) and rel_rms(avg, exp) < TOL


@pytest.mark.parametrize("n", [64, 512, 2048, 4096, 16384, 131072])
def test_other_sizes_vs_oracle(native, n):
    from wavecap_sdr_b200.dsp.fft import get_backend

    be = get_backend("cuda", fft_size=n)
    x = osp.synth_c3(seed=n, n=n + 37, fs=2_400_000)
    r = be.execute(x, 2_400_000)
    p, f, b = osp.execute(x, 2_400_000, n)
    assert rel_rms(r.power_db, p) < TOL and np.array_equal(r.freqs, f) and r.bin_hz == b


def test_short_input_gives_zeros_and_bad_size_raises(native, g):
    from wavecap_sdr_b200._native import NativeError
    from wavecap_sdr_b200.dsp.fft import get_backend

    r = get_backend("cuda", fft_size=512).execute(osp.synth_c3(seed=5, n=100), 48000)
    assert np.array_equal(r.power_db, g["short_power"]) and not r.freqs.any()
    with pytest.raises(NativeError):
        get_backend("cuda", fft_size=1000)


def test_tone_peak_property_and_device_tensors(native):
    import torch
    from wavecap_sdr_b200.dsp.fft import get_backend

    fs, n = 61_440_000, 65536
    be = get_backend("cuda", fft_size=n)
    f0 = 1234 * fs / n
    x = torch.from_numpy(np.exp(2j * np.pi * f0 * np.arange(8 * n) / fs).astype(np.complex64)).cuda()
    out = be.execute_frames(x, n_frames=8, avg=4)
    assert out.is_cuda and out.shape == (2, n)
    k = int(out[0].argmax())
    assert abs(be.freqs(fs)[k] - f0) <= fs / n
