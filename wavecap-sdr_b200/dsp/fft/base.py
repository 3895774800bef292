"""Spectrum backend contract (the names and array shapes of wavecapsdr/dsp/fft/base.py:16-81, which
`Capture._calculate_fft`, capture.py:2372-2404, and the registry program against).

A backend turns the first `fft_size` samples of an IQ block into an fftshifted dB spectrum:
Hann (symmetric, `np.hanning`) window -> FFT -> shift -> 20*log10(|X| + 1e-10), all float32.
"""
from __future__ import annotations

import abc
import dataclasses

import numpy as np


def hann_window_f32(n: int) -> np.ndarray:
    """The window every backend applies: numpy's symmetric Hann, rounded to float32 (base.py:58)."""
    return np.hanning(n).astype(np.float32)


@dataclasses.dataclass
class FFTResult:
    power_db: np.ndarray   # float32 [fft_size], 0 Hz in the middle
    freqs: np.ndarray      # float32 [fft_size], Hz, same order
    bin_hz: float          # sample_rate / fft_size

    def peak(self) -> tuple[float, float]:
        """(frequency in Hz, level in dB) of the strongest bin — convenience for tests and tools."""
        k = int(np.argmax(self.power_db))
        return float(self.freqs[k]), float(self.power_db[k])


class FFTBackend(abc.ABC):
    """Subclasses provide `name` and `execute(iq, sample_rate) -> FFTResult`; `window` is shared."""

    def __init__(self, fft_size: int = 2048):
        self.fft_size = fft_size
        self._window = None

    def __repr__(self) -> str:
        return f"{type(self).__name__}(fft_size={self.fft_size})"

    @property
    @abc.abstractmethod
    def name(self) -> str:
        """registry key of the backend ('cuda' here; 'scipy', 'fftw', 'mlx' in the reference)"""

    @abc.abstractmethod
    def execute(self, iq, sample_rate: int) -> FFTResult:
        """complex64[>= fft_size] -> FFTResult; shorter input -> all-zero spectrum (scipy_backend.py:49-56)"""

    @property
    def window(self) -> np.ndarray:
        w = self._window
        if w is None or w.shape[0] != self.fft_size:   # fft_size may be reassigned after construction
            w = self._window = hann_window_f32(self.fft_size)
        return w
