"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's voice-channel discriminator path.

  fm_discriminator        trunking/system.py:708-717  np.diff(np.unwrap([last_phase | np.angle(iq)])), last phase carried
  DiscriminatorOracle     decoders/p25.py:1105-1345   DiscriminatorDemodulator: auto gain, DC tracker, 65-tap Hamming
                                                      low-pass (np.convolve 'same' per call), MMSE interpolation with
                                                      timing / spread / frequency loops, 4-level slicer

Dtype flow. The reference mixes Python floats (constants, initial state) with np.float32 array elements; under the
NumPy >= 2 promotion rules this image runs (NEP 50: Python scalars are weak) every state variable that meets a
float32 value becomes an np.float32 and stays one. The restatement below performs the same operations on the same
operand kinds, so NumPy itself reproduces that flow; `state_dtypes()` reports what each variable ended up as (the
CUDA kernel hard-codes that result: csrc/discdemod.cu). `portable=True` replaces the float32 np.convolve (OpenBLAS
sdot: summation order depends on the host CPU) by a float64-accumulated dot product rounded once to float32 — what
the CUDA path computes. Pinned to the live reference by tests/golden/p25_discriminator.npz.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import firwin

F32 = np.float32
NTAPS, NSTEPS = 8, 128


def fm_discriminator(iq, last_phase: float):
    """(disc_audio float64 [len(iq)], new last_phase) — trunking/system.py:708-717."""
    phase = np.angle(iq)
    up = np.unwrap(np.concatenate([[last_phase], phase]))
    new_last = up[-1] if len(up) > 1 else last_phase
    return np.diff(up), new_last


def mmse_taps() -> np.ndarray:
    """decoders/p25.py:1165-1186: windowed-sinc fractional-delay table, rows normalised to unit sum, float32."""
    taps = np.zeros((NSTEPS + 1, NTAPS), dtype=F32)
    for step in range(NSTEPS + 1):
        mu = step / NSTEPS
        for tap in range(NTAPS):
            t = tap - 3 - mu
            if abs(t) < 1e-6:
                taps[step, tap] = 1.0
            else:
                s = np.sin(np.pi * t) / (np.pi * t)
                w = 0.5 * (1 + np.cos(np.pi * t / 4)) if abs(t) < 4 else 0
                taps[step, tap] = s * w
        tot = np.sum(taps[step])
        if tot > 0:
            taps[step] /= tot
    return taps


def baseband_taps(sample_rate: int) -> np.ndarray:
    cutoff = min(5200 / (sample_rate / 2), 0.99)
    return np.asarray(firwin(65, cutoff, window="hamming"), dtype=F32)


class DiscriminatorOracle:
    def __init__(self, sample_rate: int = 48000, symbol_rate: int = 4800, portable: bool = False):
        self.sample_rate, self.symbol_rate = sample_rate, symbol_rate
        self.portable = portable
        self.symbol_time = symbol_rate / sample_rate
        self.taps = mmse_taps()
        self.lpf = baseband_taps(sample_rate)
        self.input_gain = 1.0
        self.reset()

    def reset(self):  # decoders/p25.py:1335-1345 (the input gain is not reset)
        self.dc = 0.0
        self.clock = 0.0
        self.spread = 2.0
        self.fine = 0.0
        self.coarse = 0.0
        self.hist = np.zeros(NTAPS, dtype=F32)
        self.hidx = 0

    def state_dtypes(self) -> dict:
        return {k: type(getattr(self, k)).__name__ for k in ("input_gain", "dc", "clock", "spread", "fine", "coarse")}

    def _lowpass(self, x: np.ndarray) -> np.ndarray:
        if len(x) < len(self.lpf):
            return x
        if not self.portable:
            return np.convolve(x, self.lpf, mode="same").astype(F32)
        # 'same' for an odd kernel: out[n] = sum_k h[k] x[n + 32 - k], zero outside; float64 accumulation
        h = self.lpf.astype(np.float64)
        full = np.convolve(x.astype(np.float64), h, mode="full")
        return full[32:32 + len(x)].astype(F32)

    def _interp(self, mu):
        imu = min(round(mu * NSTEPS), NSTEPS)
        acc = 0.0
        for i in range(NTAPS):
            acc += self.taps[imu, i] * self.hist[(self.hidx + i) % NTAPS]
        return acc

    def demodulate(self, audio) -> np.ndarray:
        audio = np.asarray(audio)
        if audio.size == 0:
            return np.array([], dtype=np.uint8)
        x = audio.astype(F32, copy=False)
        if len(x) > 100:
            peak = np.max(np.abs(x))
            if peak > 0.01:
                self.input_gain = self.input_gain * 0.9 + (3.0 / peak) * 0.1
        x = x * self.input_gain
        a = 0.001
        for i in range(len(x)):
            self.dc = self.dc * (1 - a) + x[i] * a
            x[i] = x[i] - self.dc
        x = self._lowpass(x)
        out = []
        for s in x:
            self.hist[self.hidx] = s
            self.hidx = (self.hidx + 1) % NTAPS
            self.clock += self.symbol_time
            if not (self.clock > 1.0):
                continue
            self.clock -= 1.0
            mu = min(self.clock / self.symbol_time, 1.0)
            y = self._interp(mu)
            y1 = self._interp(min(mu + 1.0 / NSTEPS, 1.0))
            y -= self.fine
            y1 -= self.fine
            soft = 2.0 * y / self.spread
            sp = self.spread
            if y < -sp:
                err = y + (1.5 * sp)
            elif y < 0.0:
                err = y + (0.5 * sp)
            elif y < sp:
                err = y - (0.5 * sp)
            else:
                err = y - (1.5 * sp)
            if y < -sp or y >= sp:
                self.spread -= err * 0.5 * 0.0100
            elif y < 0.0:
                self.spread -= err * 0.0100
            else:
                self.spread += err * 0.0100
            self.spread = max(1.6, min(2.4, self.spread))
            if y1 < y:
                self.clock += err * 0.025
            else:
                self.clock -= err * 0.025
            self.coarse += (self.fine - self.coarse) * 0.00125
            self.fine += err * 0.125
            out.append(3 if soft < -2.0 else 2 if soft < 0.0 else 0 if soft < 2.0 else 1)
        return np.array(out, dtype=np.uint8)
