"""FFT backend interface — same contract as wavecapsdr/dsp/fft/base.py:16-81."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np


@dataclass
class FFTResult:
    """power_db: float32 [fft_size] fftshifted dB spectrum; freqs: float32 [fft_size] Hz; bin_hz: Hz/bin."""
    power_db: np.ndarray
    freqs: np.ndarray
    bin_hz: float


class FFTBackend(ABC):
    """`execute(iq, sample_rate) -> FFTResult`, `name`, `window` (base.py:31-77)."""

    def __init__(self, fft_size: int = 2048):
        self.fft_size = fft_size
        self._window = None

    @property
    def window(self) -> np.ndarray:
        if self._window is None or len(self._window) != self.fft_size:
            self._window = np.hanning(self.fft_size).astype(np.float32)
        return self._window

    @abstractmethod
    def execute(self, iq, sample_rate: int) -> FFTResult:
        ...

    @property
    @abstractmethod
    def name(self) -> str:
        ...

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(fft_size={self.fft_size})"
