"""TEST INFRASTRUCTURE ONLY — CPU restatement of the control-channel scanner measurement.

Source: /root/reference/backend/wavecapsdr/trunking/cc_scanner.py
  _measure_channel        :166-277  freq_shift (capture.py:166-193), firwin(65, 0.8/D, kaiser 6.0) lfilter from zero state,
                                    [::D], mean / max of |y|^2, noise floor = min over the two band-edge measurements
                                    (+-(fs/2 - 15 kHz - 25 kHz)), dB values, sync check only at SNR >= 8 dB
  _detect_sync_pattern    :279-351  angle(y[1:] conj(y[:-1])) sampled at [5::10], normalised correlation with the +-0.2356
                                    sync waveform at every symbol offset, first maximum of |corr|, detected if > 0.6
Pinned to the live reference by tests/golden/cc_scanner.npz (oracle/make_golden.py:gen_cc_scanner).
"""
from __future__ import annotations

import numpy as np
from scipy import signal

from .analog import freq_shift

SYNC_DIBITS = np.array([1, 1, 1, 1, 1, 3, 1, 1, 3, 3, 1, 1, 3, 3, 3, 3, 1, 3, 1, 3, 3, 3, 3, 3], dtype=np.uint8)
DEV = 0.2356


def scanner_taps(decim: int) -> np.ndarray:
    return signal.firwin(65, 0.8 / decim, window=("kaiser", 6.0))


def shift_decimate(iq, offset_hz, sample_rate, decim, taps):
    s = freq_shift(iq, offset_hz, sample_rate)
    if decim > 1:
        return signal.lfilter(taps, 1.0, s)[::decim]
    return s


def sync_correlation(y) -> float:
    """best (signed) normalised correlation, 0.0 when the block is too short (the reference then reports no sync)."""
    if len(y) < 10 * 24 + 10:
        return 0.0
    fm = np.angle(y[1:] * np.conj(y[:-1]))
    count = len(fm) // 10
    if count < 24:
        return 0.0
    sym = fm[5::10][:count]
    w = np.where(SYNC_DIBITS == 1, DEV, -DEV)
    search = min(len(sym) - 24, count - 24)
    best = 0.0
    norm_w = np.sqrt(np.sum(w ** 2))
    for i in range(max(search, 0)):
        win = sym[i:i + 24]
        c = np.sum(win * w) / (np.sqrt(np.sum(win ** 2) + 1e-10) * norm_w)
        if abs(c) > abs(best):
            best = c
    return float(best)


def measure(iq, sample_rate, center_hz, freq_hz, sync_check=True):
    """dict(power_db, peak_power_db, noise_floor_db, snr_db, sync_detected, sample_count, correlation)."""
    decim = max(1, sample_rate // 48000)
    taps = scanner_taps(decim)
    y = shift_decimate(iq, freq_hz - center_hz, sample_rate, decim, taps)
    p = np.abs(y) ** 2
    mo = sample_rate / 2 - 15000
    noise = min(np.mean(np.abs(shift_decimate(iq, e, sample_rate, decim, taps)) ** 2) for e in (-mo + 25000, mo - 25000))
    eps = 1e-12
    power_db = 10 * np.log10(np.mean(p) + eps)
    floor_db = 10 * np.log10(noise + eps)
    snr = power_db - floor_db
    corr = sync_correlation(y) if (sync_check and len(y) > 0 and snr >= 8.0) else 0.0
    return dict(power_db=float(power_db), peak_power_db=float(10 * np.log10(np.max(p) + eps)), noise_floor_db=float(floor_db),
                snr_db=float(snr), sync_detected=bool(abs(corr) > 0.6), sample_count=len(y), correlation=corr)


def synth_band(sample_rate=1_200_000, seconds=0.1, seed=71):
    """wideband test capture: C4FM control channels (valid sync words) at several offsets and levels, one CW carrier and
    AWGN. Returns (iq complex64, center_hz, channel frequency list)."""
    from . import c4fm as oc

    rng = np.random.default_rng(seed)
    n = int(sample_rate * seconds)
    center = 851_000_000.0
    chans = [(-412_500.0, 0.20), (-150_000.0, 0.03), (87_500.0, 0.08), (300_000.0, 0.004), (512_500.0, 0.0)]
    t = np.arange(n) / sample_rate
    x = 0.003 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    up = sample_rate // 48000
    for k, (off, amp) in enumerate(chans):
        if amp == 0.0:
            continue
        dib = oc.random_frames(np.random.default_rng(seed + k), n_frames=n // up // 10 // 214 + 2, payload=150, gap=40)
        bb = oc.modulate_c4fm(dib, 48000, snr_db=60.0, cfo_hz=0.0, timing=0.0, seed=seed + k, amp=1.0)
        bb = signal.resample_poly(bb, up, 1)[:n]
        if len(bb) < n:
            bb = np.concatenate([bb, np.zeros(n - len(bb))])
        x = x + amp * bb * np.exp(2j * np.pi * off * t)
    x = x + 0.05 * np.exp(2j * np.pi * 231_250.0 * t)   # an unmodulated carrier: power without sync
    return x.astype(np.complex64), center, [center + off for off, _ in chans] + [center + 231_250.0, center + 590_000.0]
