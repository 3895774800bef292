"""Oracle: P25 Phase-1 C4FM symbol recovery (wavecapsdr/dsp/p25/c4fm.py). Test infrastructure only.

Restates `C4FMDemodulator.demodulate` (c4fm.py:2528-2807) and its helpers with the dtype behaviour
the reference has AS EXECUTED here (NumPy 2.x weak scalars, numba float32*float32 -> float32
products accumulated in float64):

  stage                         reference                                    precision
  I/Q LPF + RRC                 scipy.signal.lfilter(b_f32, 1.0, x, zi)      float64, streaming state
  symbol-spaced differential    _FMDemodulator.demodulate (:324-395)          f32 products, f64 sum -> f32
  fixed-rate symbol extraction  _symbol_recovery_jit (:649-783)               float64 on f32 buffer
  sync correlation              _SoftSyncDetector (:2268-2321)                f32 products, f64 sum
  lagging detector feed         _Equalizer.get_equalized_symbol (:236-258)    float32 scalar arithmetic
  timing optimiser              _timing_optimize_jit (:543-644)               float64
  message re-slice              _resample_message_jit (:795-869)              float64

Output depends on the call (chunk) sequence exactly like the reference.
"""
from __future__ import annotations

import os

import numpy as np
from scipy import signal

F32, F64 = np.float32, np.float64
_TAPS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wavecap-sdr_b200", "dsp",
                          "p25", "interp_taps_129x8.npy")
TAPS = np.load(_TAPS_PATH)  # (129, 8) float32, c4fm.py:907-2202

LOOP_GAIN, MAX_PLL, MAX_GAIN, INITIAL_GAIN = 0.15, np.pi / 3.0, 1.25, 1.219  # c4fm.py:63-66
TSDU_MESSAGE_DIBITS = 340                                                      # c4fm.py:792
THRESH_DETECT = THRESH_OPT = 100.0                                             # c4fm.py:2408-2409
HALF_PI = 1.5707963267948966
NORM = 1.2732395447351628  # 4/pi


def design_baseband_lpf(sample_rate, passband_hz=5200.0, stopband_hz=6500.0, num_taps=63):
    """c4fm.py:95-132. The reference calls remez(..., Hz=sample_rate) inside try/except and falls back
    to firwin(num_taps, passband_hz, fs=sample_rate, window="hamming") when that raises. scipy >= 1.15
    removed the `Hz` keyword, so with the scipy of this image (1.18.1) the FALLBACK is what runs; the
    same try/except is restated here so the oracle follows the installed scipy exactly like the reference."""
    try:
        h = signal.remez(num_taps, [0, passband_hz, stopband_hz, sample_rate / 2.0], [1, 0], Hz=sample_rate)
    except Exception:
        h = signal.firwin(num_taps, passband_hz, fs=sample_rate, window="hamming")
    return np.asarray(h, dtype=F32)


def design_rrc_filter(samples_per_symbol, num_taps=101, alpha=0.2):
    """c4fm.py:135-183 (sum-normalised, float32)."""
    if num_taps % 2 == 0:
        num_taps += 1
    t = (np.arange(num_taps) - (num_taps - 1) / 2) / samples_per_symbol
    h = np.zeros(num_taps, dtype=F64)
    for i, ti in enumerate(t):
        if ti == 0:
            h[i] = 1 - alpha + 4 * alpha / np.pi
        elif abs(ti) == 1 / (4 * alpha):
            h[i] = (alpha / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha))
                                           + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))) / (
                np.pi * ti * (1 - (4 * alpha * ti) ** 2))
    return (h / np.sum(h)).astype(F32)


def sync_symbols():
    """c4fm.py:2279-2299: +-3 per dibit of 0x5575F5FF77FF (dibit 1 -> +3, everything else -> -3)."""
    pat = 0x5575F5FF77FF
    return np.array([3.0 if ((pat >> ((23 - i) * 2)) & 3) == 1 else -3.0 for i in range(24)], dtype=F32)


SYNC = sync_symbols()


def tap_row(mu: float) -> int:
    """c4fm.py:2229-2232: row = clamp(int((1-mu)*128 + 0.5), 0, 128)."""
    return min(max(int((1.0 - mu) * 128 + 0.5), 0), 128)


def interp8(window_f32: np.ndarray, row: int) -> np.ndarray:
    """_interpolate_8tap_jit (c4fm.py:400-409) vectorised over rows of `window_f32` [n, 8]:
    each product is float32, the running sum float64, taps in order 0..7."""
    p = window_f32 * TAPS[row][None, :]            # float32 products
    acc = p[:, 0].astype(F64)
    for t in range(1, 8):
        acc = acc + p[:, t]                        # float64 + float32 -> float64
    return acc


class DiffDemodOracle:
    """_FMDemodulator (c4fm.py:276-395)."""

    def __init__(self, samples_per_symbol: float, portable: bool = False):
        self.portable = portable
        self.mu = samples_per_symbol % 1.0
        self.interp_offset = max(0, int(np.floor(samples_per_symbol)) - 4)
        self.overlap = int(np.floor(samples_per_symbol)) + 4
        self.row = tap_row(self.mu)
        self.reset()

    def reset(self):
        self.i_buf = np.zeros(20, dtype=F32)
        self.q_buf = np.zeros(20, dtype=F32)

    def demodulate(self, i: np.ndarray, q: np.ndarray) -> np.ndarray:
        n = len(i)
        if n == 0:
            return np.array([], dtype=F32)
        ov = self.overlap
        res_i = self.i_buf[len(self.i_buf) - ov:] if len(self.i_buf) >= ov else self.i_buf
        res_q = self.q_buf[len(self.q_buf) - ov:] if len(self.q_buf) >= ov else self.q_buf
        bi = np.zeros(n + ov, dtype=F32)
        bq = np.zeros(n + ov, dtype=F32)
        k = min(len(res_i), ov)
        bi[:k], bq[:k] = res_i[-k:], res_q[-k:]
        bi[ov:ov + n], bq[ov:ov + n] = i, q
        self.i_buf, self.q_buf = bi, bq
        wi = np.lib.stride_tricks.sliding_window_view(bi, 8)[self.interp_offset:self.interp_offset + n]
        wq = np.lib.stride_tricks.sliding_window_view(bq, 8)[self.interp_offset:self.interp_offset + n]
        i_curr = interp8(wi, self.row).astype(F32)   # python float meets np.float32 operands -> float32
        q_curr = interp8(wq, self.row).astype(F32)
        i_prev = bi[:n]
        q_prev_conj = -bq[:n]
        diff_i = (i_prev * i_curr) - (q_prev_conj * q_curr)
        diff_q = (i_prev * q_curr) + (i_curr * q_prev_conj)
        if self.portable:   # correctly rounded float32 arctan2 (numpy's float32 SIMD arctan2 depends on the host CPU)
            return np.arctan2(diff_q.astype(F64), diff_i.astype(F64)).astype(F32)
        return np.arctan2(diff_q, diff_i).astype(F32)


def timing_score(buf, offset, pll, gain, sps):
    """_timing_score_jit (c4fm.py:416-459)."""
    score = 0.0
    max_offset = len(buf) - 8
    ptr = offset - (23.0 * sps)
    for i in range(24):
        bi = int(ptr)
        io = bi - 3
        if 0 <= io <= max_offset:
            row = int((1.0 - (ptr - bi)) * 128.0 + 0.5)
            row = min(max(row, 0), 128)
            acc = 0.0
            taps = TAPS[row]
            for j in range(8):
                acc += float(buf[io + j] * taps[j])          # float32 product, float64 sum
            score += ((acc + pll) * gain) * float(SYNC[i])
        ptr += sps
    return score


def timing_correction(buf, offset, pll, gain, sps):
    """_timing_correction_jit (c4fm.py:462-540)."""
    max_offset = len(buf) - 8
    bp = bm = ga = 0.0
    pc = mc = 0
    ptr = offset - (23.0 * sps)
    for i in range(24):
        bi = int(ptr)
        io = bi - 3
        if 0 <= io <= max_offset:
            row = int((1.0 - (ptr - bi)) * 128.0 + 0.5)
            row = min(max(row, 0), 128)
            acc = 0.0
            taps = TAPS[row]
            for j in range(8):
                acc += float(buf[io + j] * taps[j])
            soft = (acc + pll) * gain
            ideal = float(SYNC[i])
            if ideal > 0:
                bp += soft - ideal
                pc += 1
            else:
                bm += soft - ideal
                mc += 1
            ga += abs(ideal) - abs(soft)
        ptr += sps
    if pc > 0:
        bp /= -pc
    if mc > 0:
        bm /= -mc
    pllc = min(max((bp + bm) / 2.0, -HALF_PI), HALF_PI)
    return pllc, ga / (24.0 * 2.356194490192345)


def timing_optimize(buf, buffer_offset, pll, gain, sps, fine):
    """_timing_optimize_jit (c4fm.py:543-644): hill climb on the sync score."""
    step = sps / 16.0 if fine else sps / 8.0
    step_min = sps / 200.0
    max_adj = sps if fine else sps / 2.0
    adj = 0.0
    off = buffer_offset
    sc = timing_score(buf, off, pll, gain, sps)
    sl = timing_score(buf, off - step, pll, gain, sps)
    sr = timing_score(buf, off + step, pll, gain, sps)
    while step > step_min and abs(adj) <= max_adj:
        if sl > sr and sl > sc:
            adj -= step
            sr, sc = sc, sl
            sl = timing_score(buf, off + adj - step, pll, gain, sps)
        elif sr > sl and sr > sc:
            adj += step
            sl, sc = sc, sr
            sr = timing_score(buf, off + adj + step, pll, gain, sps)
        else:
            step *= 0.5
            if step > step_min:
                sl = timing_score(buf, off + adj - step, pll, gain, sps)
                sr = timing_score(buf, off + adj + step, pll, gain, sps)
    pa, ga = timing_correction(buf, off + adj, pll, gain, sps)
    return adj, sc, pa, ga


def slice_dibit(soft_rad: float) -> int:
    """pi/2 slicer (c4fm.py:750-757)."""
    if soft_rad >= HALF_PI:
        return 1
    if soft_rad >= 0:
        return 0
    if soft_rad >= -HALF_PI:
        return 2
    return 3


class SoftSyncOracle:
    """_SoftSyncDetector (c4fm.py:2268-2321)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.buf = np.zeros(48, dtype=F32)
        self.ptr = 0

    def process(self, soft) -> float:
        self.buf[self.ptr] = soft
        self.buf[self.ptr + 24] = soft
        self.ptr = (self.ptr + 1) % 24
        score = 0.0
        for i in range(24):
            score += float(SYNC[i] * self.buf[self.ptr + i])
        return score


class C4FMOracle:
    """C4FMDemodulator (c4fm.py:2379-2807)."""

    def __init__(self, sample_rate=19200, symbol_rate=4800, wide_pulse=False, portable=False):
        """portable=True replaces numpy's host-CPU-dependent float32 arctan2 by its correctly rounded value (what
        the CUDA path computes); everything else is identical. The reference goldens pin both modes."""
        self.sample_rate, self.symbol_rate = sample_rate, symbol_rate
        self.sps = sample_rate / symbol_rate
        pb, sb, alpha = (10000.0, 12000.0, 0.5) if wide_pulse else (5200.0, 6500.0, 0.2)
        self.lpf = design_baseband_lpf(sample_rate, pb, sb)
        self.rrc = design_rrc_filter(self.sps, num_taps=int(16 * self.sps) + 1, alpha=alpha)
        self.fm = DiffDemodOracle(self.sps, portable)
        self.det, self.det_lag = SoftSyncOracle(), SoftSyncOracle()
        self.lag_offset = self.sps / 2.0
        self.max_fine_adj = self.sps * 0.2
        self.reset()

    def reset(self):
        self.fm.reset()
        self.pll, self.gain, self.eq_init = 0.0, INITIAL_GAIN, False
        self.det.reset()
        self.det_lag.reset()
        self.zl_i = np.zeros(len(self.lpf) - 1, dtype=F32)
        self.zl_q = np.zeros(len(self.lpf) - 1, dtype=F32)
        self.zr_i = np.zeros(len(self.rrc) - 1, dtype=F32)
        self.zr_q = np.zeros(len(self.rrc) - 1, dtype=F32)
        self.sample_point = self.sps
        self.buf = np.zeros(65536, dtype=F32)
        self.buf_ptr = 0
        self.fine = False
        self.since_sync = 0
        self.sync_count = 0

    def phases(self, iq):
        """c4fm.py:2570-2593."""
        i = iq.real.astype(F32)
        q = iq.imag.astype(F32)
        il, self.zl_i = signal.lfilter(self.lpf, 1.0, i, zi=self.zl_i)
        ql, self.zl_q = signal.lfilter(self.lpf, 1.0, q, zi=self.zl_q)
        ir, self.zr_i = signal.lfilter(self.rrc, 1.0, il, zi=self.zr_i)
        qr, self.zr_q = signal.lfilter(self.rrc, 1.0, ql, zi=self.zr_q)
        return self.fm.demodulate(ir.astype(F32), qr.astype(F32))

    def symbol_recovery(self, phases):
        """_symbol_recovery_jit (c4fm.py:649-783), per-sample."""
        buf, blen = self.buf, len(self.buf)
        half = blen // 2
        dib, soft, idxs = [], [], []
        ptr, sp, sps, pll, gain = self.buf_ptr, self.sample_point, self.sps, self.pll, self.gain
        for ph in phases:
            ptr += 1
            sp -= 1.0
            if ptr >= blen - 1:
                buf[:half] = buf[half:].copy()
                buf[half:] = 0.0
                ptr -= half
                idxs = [(j - half) if (j - half) >= 0 else -1 for j in idxs]
            buf[ptr] = ph
            if sp < 1.0:
                mu = 1.0 - sp
                if ptr - 1 >= 0 and ptr < blen:
                    x1, x2 = buf[ptr - 1], buf[ptr]
                    if mu < 0.0:
                        interp = float(x1)
                    elif mu > 1.0:
                        interp = float(x2)
                    else:
                        interp = float(x1) + float(F32(x2 - x1)) * mu
                    sr = (interp + pll) * gain
                    dib.append(slice_dibit(sr))
                    soft.append(F32(sr * NORM))
                    idxs.append(ptr)
                sp += sps
        self.buf_ptr, self.sample_point = ptr, sp
        return np.array(dib, dtype=np.uint8), np.array(soft, dtype=F32), np.array(idxs, dtype=np.int32)

    def lag_symbol(self, offset: int, mu: float):
        """_Equalizer.get_equalized_symbol (c4fm.py:236-258) — float32 scalar arithmetic."""
        buf = self.buf
        if offset >= 0 and offset + 1 < len(buf):
            x1, x2 = buf[offset], buf[offset + 1]
            if mu < 0:
                v = x1
            elif mu > 1:
                v = x2
            else:
                v = x1 + ((x2 - x1) * F32(mu))
        else:
            v = buf[max(0, min(offset, len(buf) - 1))]
        return (v + F32(self.pll)) * F32(self.gain)

    def resample_message(self, sync_pos, num):
        """_resample_message_jit (c4fm.py:795-869)."""
        buf = self.buf
        d = np.empty(num, dtype=np.uint8)
        s = np.empty(num, dtype=F32)
        start = sync_pos + 24 * self.sps
        for i in range(num):
            pos = start + i * self.sps
            idx = int(pos)
            mu = pos - idx
            if idx >= 0 and idx + 1 < len(buf):
                x1, x2 = buf[idx], buf[idx + 1]
                if mu < 0.0:
                    v = float(x1)
                elif mu > 1.0:
                    v = float(x2)
                else:
                    v = float(x1) + float(F32(x2 - x1)) * mu
            else:
                v = float(buf[max(0, min(idx, len(buf) - 1))])
            sr = (v + self.pll) * self.gain
            d[i] = slice_dibit(sr)
            s[i] = sr * NORM
        return d, s

    def apply_correction(self, pa, ga):
        """_Equalizer.apply_correction (c4fm.py:260-272)."""
        if self.eq_init:
            self.pll += pa * LOOP_GAIN
            self.gain += ga * LOOP_GAIN
        else:
            self.pll += pa
            self.gain += ga
            self.eq_init = True
        self.pll = float(np.clip(self.pll, -MAX_PLL, MAX_PLL))
        self.gain = float(np.clip(self.gain, 1.0, MAX_GAIN))

    def demodulate(self, iq):
        """c4fm.py:2528-2807. Returns (dibits u8, soft f32); also records sync events in
        `self.events` (symbol index, optimised score, timing adjustment) for diagnostics."""
        self.events = []
        if len(iq) == 0:
            return np.array([], dtype=np.uint8), np.array([], dtype=F32)
        ph = self.phases(iq)
        dib, soft, idxs = self.symbol_recovery(ph.astype(F32))
        blen = len(self.buf)
        for k in range(len(soft)):
            self.since_sync += 1
            sp = self.det.process(soft[k])
            use_lag, extra = False, 0.0
            if self.fine or idxs[k] < 0:
                score = sp
            else:
                lag_pos = int(idxs[k]) - int(self.lag_offset)
                sl = 0.0
                if lag_pos >= 4:
                    lag_mu = 1.0 - (self.lag_offset - int(self.lag_offset))
                    lo = lag_pos - 4
                    if lo >= 0 and lag_pos < blen:
                        v = self.lag_symbol(lo, lag_mu)
                        sl = self.det_lag.process(v * F32(4.0 / np.pi))
                if sl > sp and sl >= THRESH_DETECT:
                    score, use_lag, extra = sl, True, -self.lag_offset
                else:
                    score = sp
            if score >= THRESH_DETECT:
                if idxs[k] < 0:
                    continue
                adj, osc, pa, ga = timing_optimize(self.buf, float(idxs[k]) + 0.5 + extra, self.pll, self.gain, self.sps,
                                                   self.fine)
                if osc >= THRESH_OPT:
                    if self.fine:
                        adj = float(np.clip(adj, -self.max_fine_adj, self.max_fine_adj))
                    self.sample_point += adj + extra
                    self.apply_correction(pa, ga)
                    self.sync_count += 1
                    self.fine = True
                    self.since_sync = 0
                    start = float(idxs[k]) - 23 * self.sps + adj + extra
                    nres = min(TSDU_MESSAGE_DIBITS, len(dib) - (k + 1))
                    md, ms = self.resample_message(start, nres)
                    if nres > 0:
                        dib[k + 1:k + 1 + nres] = md
                        soft[k + 1:k + 1 + nres] = ms
                    self.events.append((k, osc, adj, use_lag))
            if self.since_sync > 3600:
                self.fine = False
                self.since_sync = 0
        return dib, soft


def _demodulate_discriminator(self, disc_audio):
    """C4FMDemodulator.demodulate_discriminator (c4fm.py:2817-2992). The RRC state `rrc_state_disc` is created on the
    first call as lfilter_zi(rrc) * audio[0] and is NOT touched by reset() (the reference only creates it via hasattr)."""
    a = np.asarray(disc_audio)
    if len(a) == 0:
        return np.array([], dtype=np.uint8), np.array([], dtype=F32)
    if a.ndim > 1:
        a = a[:, 0]
    if getattr(self, "rrc_state_disc", None) is None:
        self.rrc_state_disc = signal.lfilter_zi(self.rrc, 1.0) * a[0]
    y, self.rrc_state_disc = signal.lfilter(self.rrc, 1.0, a.astype(F32), zi=self.rrc_state_disc)
    ph = y * self.sps
    dib, soft, idxs = self.symbol_recovery(ph.astype(F32))
    blen = len(self.buf)
    for k in range(len(soft)):
        self.since_sync += 1
        sp = self.det.process(soft[k])
        use_lag, extra = False, 0.0
        if self.fine or idxs[k] < 0:
            score = sp
        else:
            lag_pos = int(idxs[k]) - int(self.lag_offset)
            sl = 0.0
            if lag_pos >= 4:
                lag_mu = 1.0 - (self.lag_offset - int(self.lag_offset))
                lo = lag_pos - 4
                if lo >= 0 and lag_pos < blen:
                    v = self.lag_symbol(lo, lag_mu)
                    sl = self.det_lag.process(v * F32(4.0 / np.pi))
            if sl > sp and sl >= THRESH_DETECT:
                score, use_lag, extra = sl, True, -self.lag_offset
            else:
                score = sp
        if score >= THRESH_DETECT:
            if idxs[k] < 0:
                continue
            self.since_sync = 0
            if not self.fine:
                # the optimiser is handed the sample point, not a buffer index (:2947-2949)
                adj, _osc, _pa, _ga = timing_optimize(self.buf, self.sample_point, self.pll, self.gain, self.sps, False)
                total = adj + extra
                if abs(total) >= 0.1:
                    self.sample_point += total
                    if self.sample_point >= self.sps:
                        self.sample_point -= self.sps
                    elif self.sample_point < 0:
                        self.sample_point += self.sps
                    self.fine = True
                    self.gain = 1.0
        if self.since_sync > 3600:
            self.fine = False
            self.since_sync = 0
    return dib, soft


C4FMOracle.demodulate_discriminator = _demodulate_discriminator


def discriminator_audio(iq):
    """What the reference's callers feed demodulate_discriminator: np.diff(np.unwrap(np.angle(iq))) (float64)."""
    return np.diff(np.unwrap(np.angle(np.asarray(iq, dtype=np.complex128))))


# ---- synthetic C4FM source (recipe of scripts/generate_p25_test_signal.py:84-168) ----

def random_frames(rng, n_frames=6, payload=150, gap=40):
    """dibit stream: [sync(24) + payload random dibits] frames separated by random dibits."""
    pat = 0x5575F5FF77FF
    sync = [(pat >> ((23 - i) * 2)) & 3 for i in range(24)]
    out = list(rng.integers(0, 4, gap))
    for _ in range(n_frames):
        out += sync + list(rng.integers(0, 4, payload)) + list(rng.integers(0, 4, gap))
    return np.array(out, dtype=np.uint8)


def modulate_c4fm(dibits, sample_rate=48000, snr_db=25.0, cfo_hz=60.0, timing=0.3, seed=0, amp=0.5):
    """Recipe of the reference's scripts/generate_p25_test_signal.py:84-168 `modulate_dibits(use_rrc=True)`:
    rectangular +-1/+-3 frequency pulses (one symbol long), RRC(alpha 0.2, 8 symbols) shaping of the
    frequency, pi/4 of phase per unit level per symbol — plus a fractional timing offset, a carrier
    frequency offset and complex AWGN."""
    rng = np.random.default_rng(seed)
    sps = sample_rate / 4800.0
    level = {0: 1.0, 1: 3.0, 2: -1.0, 3: -3.0}
    n = int(np.ceil((len(dibits) + 2) * sps))
    freq = np.zeros(n, dtype=np.float64)
    for k, d in enumerate(dibits):
        freq[int((k + timing) * sps):int((k + 1 + timing) * sps)] = level[int(d)]
    alpha, span = 0.2, 8
    n_taps = int(span * sps) | 1
    t = np.arange(-(n_taps - 1) // 2, (n_taps + 1) // 2) / sps
    with np.errstate(divide="ignore", invalid="ignore"):
        h = (np.sin(np.pi * t * (1 - alpha)) + 4 * alpha * t * np.cos(np.pi * t * (1 + alpha))) / (
            np.pi * t * (1 - (4 * alpha * t) ** 2))
    h[t == 0] = 1 + alpha * (4 / np.pi - 1)
    h[~np.isfinite(h)] = 0.0
    h /= h.sum()
    freq = signal.lfilter(h, 1.0, freq)
    phase = np.cumsum(freq * (np.pi / 4) / sps) + 2 * np.pi * cfo_hz / sample_rate * np.arange(n)
    x = amp * np.exp(1j * phase)
    sigma = amp * 10 ** (-snr_db / 20) / np.sqrt(2)
    x = x + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)
