"""Oracle: spectrum producer (wavecapsdr/dsp/fft/base.py:31-77, scipy_backend.py:38-79) and the
frontend's dB averaging (frontend/src/components/primitives/SpectrumAnalyzer.react.tsx:309-327).
Test infrastructure only."""
from __future__ import annotations

import numpy as np
from scipy.fft import fft, fftfreq, fftshift


def hann_window(n: int) -> np.ndarray:
    """fft/base.py:54-59: symmetric np.hanning cast to float32."""
    return np.hanning(n).astype(np.float32)


def execute(iq: np.ndarray, sample_rate: int, fft_size: int):
    """scipy_backend.py:38-79 -> (power_db f32, freqs f32, bin_hz). Too few samples -> zeros."""
    if iq.size < fft_size:
        z = np.zeros(fft_size, dtype=np.float32)
        return z, z.copy(), sample_rate / fft_size
    windowed = iq[:fft_size] * hann_window(fft_size)
    spec = fftshift(fft(windowed))
    freqs = fftshift(fftfreq(fft_size, 1.0 / sample_rate))
    power_db = 20.0 * np.log10(np.abs(spec) + 1e-10)
    return power_db.astype(np.float32), freqs.astype(np.float32), sample_rate / fft_size


def averaged(frames_db: np.ndarray, k: int = 4) -> np.ndarray:
    """Arithmetic mean of consecutive groups of k dB frames (SpectrumAnalyzer.react.tsx:309-327)."""
    n = frames_db.shape[0]
    return np.stack([frames_db[i:i + k].astype(np.float64).mean(axis=0) for i in range(0, n, k)]).astype(np.float32)


def synth_c3(seed: int = 3, n: int = 65536, fs: int = 61_440_000, t0: int = 0) -> np.ndarray:
    """5 tones (-20 .. -80 dBFS) + AWGN, cf32."""
    rng = np.random.default_rng(seed)
    t = (np.arange(n) + t0) / fs
    x = 0.003 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for db, f in ((-20, 5.1e6), (-35, -12.3e6), (-50, 20.02e6), (-65, -25.5e6), (-80, 1.234e6)):
        x = x + 10 ** (db / 20) * np.exp(2j * np.pi * f * t)
    return x.astype(np.complex64)
