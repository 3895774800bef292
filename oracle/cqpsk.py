"""Oracle: P25 CQPSK / LSM symbol recovery (wavecapsdr/decoders/p25.py:190-669). Test infrastructure only.

Restates `CQPSKDemodulator.demodulate` with the scalar dtype flow the reference has AS EXECUTED here
(NumPy 2.x weak Python scalars; inspected on the live object):

  variable                      precision in the reference                      why
  _agc_gain                     float32                                         Python float met np.mean(float32)
  x after AGC                   complex64
  NCO (from the 2nd chunk on)   complex128 product                              np.exp(-1j * float64 array)
  LPF ('same', per chunk)       float32 dot (1st chunk) / float64 (after NCO)   np.convolve dtype of x.real
  history / curr / prev         complex64, float32 arithmetic
  _symbol_clock/_time/_omega    Python float until the first TED update, float32 afterwards
  _freq_offset                  float64 (first symbol: np.conj(0j) is complex128, so phase is float64)
  _phase_acc                    float64

The per-sample loop of `_cqpsk_timing_recovery` (:481-669) is restated without the 32-entry ring:
"k samples back from the newest" indexes the filtered chunk extended by the last 32 filtered samples
of the previous call, which is what the ring holds.
"""
from __future__ import annotations

import numpy as np
from scipy import signal

F32, F64, C64 = np.float32, np.float64, np.complex64
NTAPS_HIST = 32       # MMSE_NTAPS, decoders/p25.py:221
NSTEPS = 128          # MMSE_NSTEPS, :222


def mmse_table():
    """_generate_mmse_taps (decoders/p25.py:289-323): Hann-windowed sinc, 129 x 8, rows sum to 1, float32."""
    taps = np.zeros((NSTEPS + 1, 8), dtype=F32)
    for step in range(NSTEPS + 1):
        mu = step / NSTEPS
        for tap in range(8):
            t = tap - 3 - mu
            if abs(t) < 1e-6:
                taps[step, tap] = 1.0
            else:
                sinc_val = np.sin(np.pi * t) / (np.pi * t)
                window = 0.5 * (1 + np.cos(np.pi * t / 4)) if abs(t) < 4 else 0
                taps[step, tap] = sinc_val * window
        s = np.sum(taps[step])
        if abs(s) > 1e-6:
            taps[step] /= s
    return taps


def baseband_taps(sample_rate, cutoff_hz=7250, num_taps=63):
    """_design_baseband_filter (decoders/p25.py:375-388)."""
    norm = min(0.99, max(0.01, cutoff_hz / (sample_rate / 2)))
    return np.asarray(signal.firwin(num_taps, norm, window="hamming"), dtype=F32)


def _angle32(z):
    """np.angle for a complex64 scalar with a CORRECTLY ROUNDED float32 result (float64 atan2, then cast).
    numpy's own float32 arctan2 is a SIMD approximation whose last bit depends on the host CPU."""
    return np.float32(np.arctan2(np.float64(z.imag), np.float64(z.real)))


class CQPSKOracle:
    """portable=False: literally the reference's numpy calls (pins the restatement to the reference on THIS host).
    portable=True: the three host-CPU-dependent calls are replaced by their correctly rounded values — float32
    arctan2 (SVML), the first-call float32 np.convolve (OpenBLAS sdot order) and the pairwise float32 np.mean —
    so that the expected output does not depend on which CPU the test runs on. That is also what the CUDA path
    computes; GPU parity tests use this mode, the reference goldens pin both modes."""

    def __init__(self, sample_rate=19200, symbol_rate=4800, portable=False):
        self.portable = portable
        self.sample_rate, self.symbol_rate = sample_rate, symbol_rate
        self.sps = sample_rate / symbol_rate
        self.half_pi, self.quarter_pi, self.three_quarter_pi = np.pi / 2, np.pi / 4, 3 * np.pi / 4
        self.freq_offset = 0.0            # -> float64 after the first symbol
        self.beta, self.fmin, self.fmax = 0.0005, -0.02, 0.02
        self.phase_acc = 0.0
        self.agc_gain, self.agc_alpha, self.agc_target = 1.0, 0.005, 1.0
        self.gain_mu, self.gain_omega = 0.015, 0.0
        self.omega = self.sps
        self.prev_symbol = 0.0 + 0.0j     # Python complex until the first symbol
        self.clock = 0.0
        self.sym_time = 1.0 / self.sps
        self.taps = baseband_taps(sample_rate)
        self.mmse = mmse_table()
        self.tail = np.zeros(NTAPS_HIST, dtype=C64)   # the ring's content, oldest first

    def _interp(self, ext, newest, back, imu):
        """_mmse_interpolate_at_offset (decoders/p25.py:325-359): taps -3..+4 around `back` samples
        before the newest one; history slots outside [0, 32) are skipped."""
        r = 0.0 + 0.0j
        for tap in range(8):
            off = back + (tap - 3)
            if 0 <= off < NTAPS_HIST:
                r += self.mmse[imu, tap] * ext[newest - off]
        return r

    def frontend(self, iq):
        """AGC + NCO + low-pass of one call (decoders/p25.py:413-471) -> complex64 chunk."""
        x = iq.astype(C64, copy=False)
        mags = np.abs(x)
        mean_mag = np.float32(np.mean(mags, dtype=np.float64)) if self.portable else np.mean(mags)
        if mean_mag > 1e-8:
            tg = self.agc_target / mean_mag
            self.agc_gain = self.agc_gain * (1 - self.agc_alpha) + tg * self.agc_alpha
            self.agc_gain = np.clip(self.agc_gain, 0.01, 500.0)
        x = x * self.agc_gain
        if abs(self.freq_offset) > 1e-7:
            n = np.arange(len(x))
            x = x * np.exp(-1j * (self.phase_acc + self.freq_offset * n))
            self.phase_acc += self.freq_offset * len(x)
            self.phase_acc = np.angle(np.exp(1j * self.phase_acc))
        if len(x) >= len(self.taps):
            t = self.taps.astype(np.float64) if self.portable else self.taps
            xi = np.convolve(x.real.astype(np.float64) if self.portable else x.real, t, mode="same")
            xq = np.convolve(x.imag.astype(np.float64) if self.portable else x.imag, t, mode="same")
            x = (xi + 1j * xq).astype(C64)
        return x

    def demodulate(self, iq):
        if iq.size == 0:
            return np.array([], dtype=np.uint8)
        x = self.frontend(np.asarray(iq))
        if x.dtype != C64:          # chunk shorter than the filter after an NCO: stays complex128 in the reference;
            x = x.astype(C64)       # its ring is complex64, so the store rounds exactly like this cast
        ext = np.concatenate([self.tail, x])
        sps = self.sps
        half_sps, full_sps = int(round(sps / 2)), int(round(sps))
        out = []
        self.phases = []          # diagnostics: differential phase of every symbol of this call
        for n in range(len(x)):
            newest = NTAPS_HIST + n
            self.clock += self.sym_time
            if self.clock >= 1.0:
                self.clock -= 1.0
                mu = np.clip(self.clock / self.sym_time, 0.0, 1.0 - 1e-6)
                imu = min(round(mu * NSTEPS), NSTEPS)
                curr = self._interp(ext, newest, 0, imu)
                cm, pm = abs(curr), abs(self.prev_symbol)
                if cm > 1e-6 and pm > 1e-6:
                    diff = (curr / cm) * np.conj(self.prev_symbol / pm)
                else:
                    diff = curr * np.conj(self.prev_symbol)
                phase = _angle32(diff) if (self.portable and diff.dtype == C64) else np.angle(diff)
                self.phases.append(float(phase))
                if phase >= self.half_pi:
                    dibit, expected = 1, self.three_quarter_pi
                elif phase >= 0:
                    dibit, expected = 0, self.quarter_pi
                elif phase >= -self.half_pi:
                    dibit, expected = 2, -self.quarter_pi
                else:
                    dibit, expected = 3, -self.three_quarter_pi
                out.append(dibit)
                pe = phase - expected
                if pe > np.pi:
                    pe -= 2 * np.pi
                elif pe < -np.pi:
                    pe += 2 * np.pi
                self.freq_offset += self.beta * pe * cm
                self.freq_offset = np.clip(self.freq_offset, self.fmin, self.fmax)
                if full_sps + 4 < NTAPS_HIST:
                    mid = self._interp(ext, newest, half_sps, imu)
                    prv = self._interp(ext, newest, full_sps, imu)
                    ted = np.real((curr - prv) * np.conj(mid))
                    self.clock += self.gain_mu * ted
                    self.omega += self.gain_omega * ted
                    self.omega = np.clip(self.omega, sps * 0.95, sps * 1.05)
                    self.sym_time = 1.0 / self.omega
                while self.clock >= 1.0:
                    self.clock -= 1.0
                while self.clock < 0.0:
                    self.clock += 1.0
                self.prev_symbol = curr
        self.tail = ext[len(ext) - NTAPS_HIST:].copy()
        return np.array(out, dtype=np.uint8)


# ---- synthetic pi/4-DQPSK source ----

def modulate_cqpsk(dibits, sample_rate=48000, symbol_rate=4800, snr_db=25.0, cfo_hz=40.0, timing=0.3, seed=0, amp=0.4):
    """pi/4-DQPSK: phase steps {+pi/4, +3pi/4, -pi/4, -3pi/4} for dibits {0, 1, 2, 3}
    (decoders/p25.py:548-560 slicer mapping), RRC(alpha 0.2) shaped impulses, CFO, timing offset, AWGN."""
    rng = np.random.default_rng(seed)
    sps = sample_rate / symbol_rate
    step = {0: np.pi / 4, 1: 3 * np.pi / 4, 2: -np.pi / 4, 3: -3 * np.pi / 4}
    ph = np.cumsum([step[int(d)] for d in dibits])
    n = int(np.ceil((len(dibits) + 8) * sps))
    imp = np.zeros(n, dtype=np.complex128)
    for k, p in enumerate(ph):
        imp[int(round((k + 4 + timing) * sps))] = np.exp(1j * p)
    alpha, span = 0.2, 8
    t = np.arange(-span * sps, span * sps + 1) / sps
    with np.errstate(divide="ignore", invalid="ignore"):
        h = (np.sin(np.pi * t * (1 - alpha)) + 4 * alpha * t * np.cos(np.pi * t * (1 + alpha))) / (
            np.pi * t * (1 - (4 * alpha * t) ** 2))
    h[t == 0] = 1 - alpha + 4 * alpha / np.pi
    h[~np.isfinite(h)] = 0.0
    h /= np.max(h)
    x = amp * np.convolve(imp, h, mode="same")
    x = x * np.exp(2j * np.pi * cfo_hz / sample_rate * np.arange(n))
    sigma = amp * 10 ** (-snr_db / 20) / np.sqrt(2)
    x = x + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)
