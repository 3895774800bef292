"""GPU AM / SSB demodulation with the call surface of `wavecapsdr.dsp.am` (dsp/am.py)."""
from __future__ import annotations

import numpy as np

from . import _stages as S
from . import agc as A
from . import filters as F


def _n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def freq_shift(iq, offset_hz: float, sample_rate: int):
    """iq * complex64(exp(+2j pi f t)), t = n/fs in float64 (dsp/am.py:23-42)."""
    if _n(iq) == 0:
        return iq
    import torch

    x = S.to_device(iq, np.complex64).reshape(-1)
    n = torch.arange(x.numel(), device=x.device, dtype=torch.float64)
    ph = (n / float(sample_rate)) * float(offset_hz)
    ph = 2.0 * np.pi * (ph - torch.round(ph))
    shift = torch.complex(torch.cos(ph), torch.sin(ph)).to(torch.complex64)
    return S.like_input(x * shift, iq)


def am_post_chain(sample_rate: int, enable_highpass=True, highpass_hz=100, enable_lowpass=True, lowpass_hz=5000,
                  notch_frequencies=None) -> list:
    """IIR stages of am_demod in order (dsp/am.py:105-119)."""
    st = []
    if enable_highpass and highpass_hz > 0:
        st.append(F.highpass_coeffs(sample_rate, highpass_hz))
    if enable_lowpass and lowpass_hz > 0:
        st.append(F.lowpass_coeffs(sample_rate, lowpass_hz))
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            st.append(F.notch_coeffs(sample_rate, f, 30.0))
    return [s for s in st if s is not None]


def ssb_post_chain(sample_rate: int, enable_bandpass=True, bandpass_low=300, bandpass_high=3000,
                   notch_frequencies=None) -> list:
    """IIR stages of ssb_demod in order (dsp/am.py:231-239)."""
    st = []
    if enable_bandpass:
        st.append(F.bandpass_coeffs(sample_rate, bandpass_low, bandpass_high))
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            st.append(F.notch_coeffs(sample_rate, f, 30.0))
    return [s for s in st if s is not None]


def am_tail(rows, sample_rate: int, audio_rate: int, stages, enable_agc: bool, agc_target_db: float,
            want_stats: bool = False, blanker_db=None):
    """[noise blanker] -> IIRs -> [AGC] -> resample -> [agc soft clip when AGC is off] (dsp/am.py:100-141, 213-247)."""
    y = rows
    if blanker_db is not None:
        y = F.noise_blanker_rows(y, blanker_db, 3)   # on the real envelope / real part, ahead of the filters (:100-101, :213-215)
    for b, a in stages:
        y = S.lfilter(b, a, y)
    if enable_agc:
        y = A.agc_rows(y, sample_rate, target_db=agc_target_db, attack_ms=5.0, release_ms=50.0)
    epi = S.EPI_NONE if enable_agc else S.EPI_CLIP_AGC
    if sample_rate == audio_rate:
        out = y if enable_agc else S.elementwise(y, S.OP_SOFT_CLIP_AGC)
        if want_stats:
            import torch

            return out, (out.double() ** 2).sum(dim=1), (~torch.isfinite(out).all(dim=1)).int()
        return out
    up, down = S.rate_ratio(sample_rate, audio_rate)
    return S.resample(y, up, down, epi, want_stats=want_stats)


def am_demod(iq, sample_rate: int, audio_rate: int = 48_000, enable_agc: bool = True, enable_highpass: bool = True,
             highpass_hz: float = 100, enable_lowpass: bool = True, lowpass_hz: float = 5000,
             enable_noise_blanker: bool = False, noise_blanker_threshold_db: float = 10.0,
             agc_target_db: float = -20.0, notch_frequencies=None):
    """Envelope AM demodulation (dsp/am.py:45-141)."""
    if _n(iq) == 0:
        return np.empty(0, dtype=np.float32)
    x = S.to_device(iq, np.complex64).reshape(-1)
    env, _, _, _ = S.front(x, S.FMT_CF32, x.numel(), 1, [S.MODE_AM], [0.0], None, int(sample_rate))
    stages = am_post_chain(sample_rate, enable_highpass, highpass_hz, enable_lowpass, lowpass_hz, notch_frequencies)
    out = am_tail(env.reshape(1, -1), int(sample_rate), int(audio_rate), stages, enable_agc, agc_target_db,
                  blanker_db=noise_blanker_threshold_db if enable_noise_blanker else None)
    return S.like_input(out.reshape(-1), iq)


def ssb_demod(iq, sample_rate: int, audio_rate: int = 48_000, mode: str = "usb", enable_agc: bool = True,
              enable_bandpass: bool = True, bandpass_low: float = 300, bandpass_high: float = 3000,
              enable_noise_blanker: bool = False, noise_blanker_threshold_db: float = 10.0,
              agc_target_db: float = -20.0, notch_frequencies=None, bfo_offset_hz: float = 1500.0):
    """SSB product detection with BFO (dsp/am.py:144-247)."""
    if _n(iq) == 0:
        return np.empty(0, dtype=np.float32)
    x = S.to_device(iq, np.complex64).reshape(-1)
    bfo = bfo_offset_hz if mode.lower() == "usb" else -bfo_offset_hz
    re, _, _, _ = S.front(x, S.FMT_CF32, x.numel(), 1, [S.MODE_SSB], [0.0], [bfo], int(sample_rate))
    stages = ssb_post_chain(sample_rate, enable_bandpass, bandpass_low, bandpass_high, notch_frequencies)
    out = am_tail(re.reshape(1, -1), int(sample_rate), int(audio_rate), stages, enable_agc, agc_target_db,
                  blanker_db=noise_blanker_threshold_db if enable_noise_blanker else None)
    return S.like_input(out.reshape(-1), iq)
