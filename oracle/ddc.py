"""Oracle: streaming complex FIR and the trunking fan-out (NCO + two-stage decimation). Test infrastructure only.

  fir_filter_complex / fir_decimate   wavecapsdr/dsp/filters.py:558-668 (+ the numba kernels :470-556 they call)
  PhaseContinuousNCO                  trunking/system.py:1434-1466 (closure `phase_continuous_freq_shift`),
                                      :561-586 (VoiceRecorder variant, same arithmetic)
  ControlChannelDDC                   trunking/system.py:1392-1406 (filter design), :1753-1779 (two fir_decimate stages,
                                      zi = lfilter_zi * first sample handed to fir_decimate as its state)
  VoiceDDC                            trunking/system.py:453-519 (design), :627-656 (two scipy lfilter stages, complex128)

The closures/methods are not importable in isolation (they live inside TrunkingSystem / VoiceRecorder), so
the two DDC classes restate them line by line around the importable `fir_decimate`, which IS pinned to the
live reference and to tests/golden/ddc.npz.
"""
from __future__ import annotations

import numpy as np
from scipy import signal


def fir_filter_complex(x, taps, zi=None):
    """filters.py:558-620: y[i] = sum_j taps[j] * state[n_zi + i - j] over state = [zi | x] in complex128,
    output complex64, new zi = last n_zi entries of state (complex128)."""
    taps = np.asarray(taps, dtype=np.float64)
    n_zi = len(taps) - 1
    if len(x) == 0:
        return np.empty(0, np.complex64), (zi if zi is not None else np.zeros(n_zi, np.complex128))
    z = np.zeros(n_zi, np.complex128) if zi is None else np.asarray(zi).astype(np.complex128)
    state = np.concatenate([z, np.asarray(x).astype(np.complex128)])
    y = np.convolve(state, taps)[n_zi:n_zi + len(x)]
    return y.astype(np.complex64), state[len(x):len(x) + n_zi].copy()


def fir_decimate(x, taps, decim_factor, zi=None):
    """filters.py:623-652: filter everything, keep every decim_factor-th sample of THIS call."""
    y, nzi = fir_filter_complex(x, taps, zi)
    return y[::decim_factor], nzi


class PhaseContinuousNCO:
    """system.py:1434-1466."""

    def __init__(self, sample_rate: int):
        self.fs = sample_rate
        self.sample_idx = 0
        self.last_offset = 0.0

    def shift(self, iq, offset_hz: float):
        if offset_hz == 0.0 or iq.size == 0:
            return iq
        if offset_hz != self.last_offset:
            self.sample_idx = 0
            self.last_offset = offset_hz
        n = np.arange(iq.size, dtype=np.float64) + self.sample_idx
        phase = -2.0 * np.pi * offset_hz * n / self.fs
        out = np.asarray(iq.astype(np.complex64, copy=False) * np.exp(1j * phase).astype(np.complex64), dtype=np.complex64)
        self.sample_idx += iq.size
        if self.sample_idx >= self.fs:
            self.sample_idx %= self.fs
        return out


def design(decim1: int, decim2: int):
    """system.py:1392-1406 / :488-505."""
    t1 = signal.firwin(157, 0.8 / decim1, window=("kaiser", 7.857))
    t2 = signal.firwin(73, 0.8 / decim2, window=("kaiser", 7.857)) if decim2 > 1 else None
    return t1, t2


class ControlChannelDDC:
    """on_raw_iq_callback's per-channel chain (system.py:1753-1779): complex64 between and after the stages."""

    def __init__(self, sample_rate: int, decim1: int, decim2: int, offset_hz: float):
        self.nco = PhaseContinuousNCO(sample_rate)
        self.offset = offset_hz
        self.d1, self.d2 = decim1, decim2
        self.t1, self.t2 = design(decim1, decim2)
        self.z1t = signal.lfilter_zi(self.t1, 1.0).astype(np.complex128)
        self.z2t = signal.lfilter_zi(self.t2, 1.0).astype(np.complex128) if self.t2 is not None else None
        self.z1 = self.z2 = None

    def process(self, iq):
        c = self.nco.shift(iq, self.offset)
        if len(c) == 0:
            return c
        if self.z1 is None:
            self.z1 = self.z1t * c[0]
        y1, self.z1 = fir_decimate(c, self.t1, self.d1, zi=self.z1)
        if self.t2 is None or y1.size == 0:
            return y1
        if self.z2 is None:
            self.z2 = self.z2t * y1[0]
        y2, self.z2 = fir_decimate(y1, self.t2, self.d2, zi=self.z2)
        return y2


class VoiceDDC:
    """VoiceRecorder.process_iq's chain (system.py:561-656): scipy lfilter with true DF2T state, complex128 result."""

    def __init__(self, sample_rate: int, decim1: int, decim2: int, offset_hz: float):
        self.nco = PhaseContinuousNCO(sample_rate)
        self.offset = offset_hz
        self.d1, self.d2 = decim1, decim2
        self.t1, self.t2 = design(decim1, decim2)
        self.z1t = signal.lfilter_zi(self.t1, 1.0).astype(np.complex128)
        self.z2t = signal.lfilter_zi(self.t2, 1.0).astype(np.complex128) if self.t2 is not None else None
        self.z1 = self.z2 = None

    def process(self, iq):
        c = self.nco.shift(iq, self.offset)
        if self.z1 is None:
            self.z1 = self.z1t * c[0]
        f1, self.z1 = signal.lfilter(self.t1, 1.0, c, zi=self.z1)
        y1 = f1[:: self.d1]
        if self.t2 is None:
            return y1
        if self.z2 is None:
            self.z2 = self.z2t * y1[0]
        f2, self.z2 = signal.lfilter(self.t2, 1.0, y1, zi=self.z2)
        return f2[:: self.d2]


def synth_wideband(seed: int, n: int, fs: int, offsets, t0: int = 0):
    """narrowband carriers (slow random-phase FM) at the given offsets + AWGN, complex64; deterministic in (seed, t0)."""
    rng = np.random.default_rng(seed * 1000003 + t0)
    t = (np.arange(n, dtype=np.float64) + t0) / fs
    x = 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for i, off in enumerate(offsets):
        x = x + (0.1 + 0.02 * i) * np.exp(1j * (2 * np.pi * off * t + 2.0 * np.sin(2 * np.pi * (700 + 130 * i) * t)))
    return x.astype(np.complex64)
