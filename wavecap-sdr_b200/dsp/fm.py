"""GPU FM demodulation with the call surface of `wavecapsdr.dsp.fm` (dsp/fm.py).

Every function accepts numpy arrays (returns numpy) or CUDA tensors (returns CUDA tensors, stays on
the device). `wbfm_demod` / `nbfm_demod` run the whole chain on the device with ONE upload and ONE
download: discriminator -> [de-emphasis] -> [Butterworth / notch IIRs] -> RMS -> polyphase resampler
with the RMS scale and tanh soft clip fused into its epilogue.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from . import _stages as S
from . import filters as F


def _n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def soft_clip(x):
    """tanh(1.5 x)/tanh(1.5) * 0.95 (dsp/fm.py:26-39)."""
    if _n(x) == 0:
        return F._f32_passthrough(x)
    t = S.to_device(x, np.float32)
    return S.like_input(S.elementwise(t, S.OP_SOFT_CLIP_FM), x)


def rms_normalize(x, target_rms: float = 0.18, min_rms: float = 1e-4):
    """Scale to the target RMS unless the RMS is below min_rms (dsp/fm.py:42-62)."""
    if _n(x) == 0:
        return x
    t = S.to_device(x, np.float32)
    rows = t.reshape(1, -1)
    rms = float(np.float32(np.sqrt(float(S.sumsq(rows)[0].item()) / rows.shape[1])))
    if rms > min_rms:
        return S.like_input(S.elementwise(t, S.OP_SCALE, float(np.float32(target_rms / rms))), x)
    return x


def quadrature_demod(iq, sample_rate: int):
    """out[0]=0, out[n]=angle(x[n] conj(x[n-1])) * float32(fs/(2 pi 75000)) (dsp/fm.py:65-97)."""
    if _n(iq) == 0:
        return np.empty(0, dtype=np.float32)
    x = S.to_device(iq, np.complex64).reshape(-1)
    out, _, _, _ = S.front(x, S.FMT_CF32, x.numel(), 1, [S.MODE_NBFM], [0.0], None, int(sample_rate))
    return S.like_input(out.reshape(-1), iq)


@lru_cache(maxsize=32)
def _deemph_coeffs(sample_rate: int, tau_us: int):
    # dsp/fm.py:101-109: float32 coefficients b=[alpha], a=[1, -(1-alpha)]
    tau = tau_us * 1e-6
    alpha = 1.0 / (1.0 + (1.0 / (2.0 * np.pi * tau * sample_rate)))
    b = np.array([alpha], dtype=np.float32)
    a = np.array([1.0, -(1.0 - alpha)], dtype=np.float32)
    return tuple(float(v) for v in b), tuple(float(v) for v in a)


def deemphasis_coeffs(sample_rate: int, tau: float = 75e-6):
    return _deemph_coeffs(int(sample_rate), int(tau * 1e6))


def deemphasis_filter(x, sample_rate: int, tau: float = 75e-6):
    """One-pole de-emphasis (dsp/fm.py:112-126); the float32 recursion is evaluated in float64
    (3.6e-8 relative RMS from the reference)."""
    return F._run_iir(x, deemphasis_coeffs(sample_rate, tau))


def mpx_coeffs(sample_rate: int, cutoff: float = 15_000, order: int = 5):
    wn = int(cutoff) / (sample_rate / 2.0)  # integer cut-off key (dsp/fm.py:171)
    return None if wn >= 1.0 else F._butter("low", (wn,), order)


def lpf_audio(x, sample_rate: int, cutoff: float = 15_000):
    """Order-5 Butterworth MPX low-pass (dsp/fm.py:130-181)."""
    return F._run_iir(x, mpx_coeffs(sample_rate, cutoff))


def resample_poly(x, in_rate: int, out_rate: int):
    """scipy.signal.resample_poly(x.astype(f64), up, down).astype(f32) (dsp/fm.py:184-221)."""
    if _n(x) == 0 or in_rate == out_rate:
        return F._f32_passthrough(x)
    up, down = S.rate_ratio(in_rate, out_rate)
    t = S.to_device(x, np.float32).reshape(1, -1)
    return S.like_input(S.resample(t, up, down).reshape(-1), x)


resample_linear = resample_poly  # dsp/fm.py:225


def fm_post_chain(sample_rate: int, *, wide: bool, enable_deemphasis: bool, deemphasis_tau: float,
                  enable_mpx_filter: bool = False, mpx_cutoff_hz: float = 15_000, enable_highpass: bool = False,
                  highpass_hz: float = 100, enable_lowpass: bool = False, lowpass_hz: float = 3_000,
                  notch_frequencies=None) -> list:
    """The ordered list of (b, a) IIR stages between discriminator and RMS for wbfm (dsp/fm.py:279-301)
    and nbfm (:372-393)."""
    stages = []
    if enable_deemphasis:
        stages.append(deemphasis_coeffs(sample_rate, deemphasis_tau))
    if wide:
        if enable_mpx_filter:
            stages.append(mpx_coeffs(sample_rate, mpx_cutoff_hz))
        if enable_highpass and highpass_hz > 0:
            stages.append(F.highpass_coeffs(sample_rate, highpass_hz))
    else:
        if enable_highpass and highpass_hz > 0:
            stages.append(F.highpass_coeffs(sample_rate, highpass_hz))
        if enable_lowpass and lowpass_hz > 0:
            stages.append(F.lowpass_coeffs(sample_rate, lowpass_hz))
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            stages.append(F.notch_coeffs(sample_rate, f, 30.0))
    return [s for s in stages if s is not None]


def fm_tail(fm_rows, sample_rate: int, audio_rate: int, stages, want_stats: bool = False, blanker_db=None, nr_db=None,
            sumsq_rows=None):
    """[noise blanker] -> IIR stages -> [spectral noise reduction] -> RMS -> resample with fused scale + soft clip
    (dsp/fm.py:277-314, :368-406). fm_rows: CUDA [n_seq, n] float32. sumsq_rows: sum(fm_rows**2) per row when the front
    end already produced it (only used if nothing between discriminator and RMS changes the signal)."""
    y = fm_rows
    if blanker_db is not None:
        y = F.noise_blanker_rows(y, blanker_db, 3)
    for b, a in stages:
        y = S.lfilter(b, a, y)
    if nr_db is not None:
        y = F.spectral_nr_rows(y, nr_db)
    ss = sumsq_rows if (sumsq_rows is not None and y is fm_rows) else S.sumsq(y)
    if sample_rate == audio_rate:
        # resample_linear returns its input; scale and clip elementwise
        import torch

        rms = torch.sqrt(ss / y.shape[1]).to(torch.float32)
        scale = torch.where(rms > 1e-4, (0.18 / rms.double()).float(), torch.ones_like(rms))
        out = S.elementwise(y * scale[:, None], S.OP_SOFT_CLIP_FM)
        if want_stats:
            return out, (out.double() ** 2).sum(dim=1), (~torch.isfinite(out).all(dim=1)).int()
        return out
    up, down = S.rate_ratio(sample_rate, audio_rate)
    return S.resample(y, up, down, S.EPI_RMS_CLIP, ss, 0.18, 1e-4, want_stats=want_stats)


def _fm_demod(iq, sample_rate, audio_rate, stages, blanker_db, nr_db):
    if _n(iq) == 0:
        return np.empty(0, dtype=np.float32)
    x = S.to_device(iq, np.complex64).reshape(-1)
    fm, _, _, _ = S.front(x, S.FMT_CF32, x.numel(), 1, [S.MODE_NBFM], [0.0], None, int(sample_rate))
    audio = fm_tail(fm.reshape(1, -1), int(sample_rate), int(audio_rate), stages, blanker_db=blanker_db, nr_db=nr_db)
    return S.like_input(audio.reshape(-1), iq)


def wbfm_demod(iq, sample_rate: int, audio_rate: int = 48_000, enable_deemphasis: bool = True,
               deemphasis_tau: float = 75e-6, enable_mpx_filter: bool = True, mpx_cutoff_hz: float = 15_000,
               enable_highpass: bool = False, highpass_hz: float = 100, enable_noise_blanker: bool = False,
               noise_blanker_threshold_db: float = 10.0, notch_frequencies=None,
               enable_noise_reduction: bool = False, noise_reduction_db: float = 12.0):
    """Wideband FM (dsp/fm.py:228-314)."""
    stages = fm_post_chain(sample_rate, wide=True, enable_deemphasis=enable_deemphasis,
                           deemphasis_tau=deemphasis_tau, enable_mpx_filter=enable_mpx_filter,
                           mpx_cutoff_hz=mpx_cutoff_hz, enable_highpass=enable_highpass, highpass_hz=highpass_hz,
                           notch_frequencies=notch_frequencies)
    return _fm_demod(iq, sample_rate, audio_rate, stages, noise_blanker_threshold_db if enable_noise_blanker else None,
                     noise_reduction_db if enable_noise_reduction else None)


def nbfm_demod(iq, sample_rate: int, audio_rate: int = 48_000, enable_deemphasis: bool = False,
               deemphasis_tau: float = 75e-6, enable_highpass: bool = False, highpass_hz: float = 300,
               enable_lowpass: bool = False, lowpass_hz: float = 3_000, enable_noise_blanker: bool = False,
               noise_blanker_threshold_db: float = 10.0, notch_frequencies=None,
               enable_noise_reduction: bool = False, noise_reduction_db: float = 12.0):
    """Narrowband FM (dsp/fm.py:317-406)."""
    stages = fm_post_chain(sample_rate, wide=False, enable_deemphasis=enable_deemphasis,
                           deemphasis_tau=deemphasis_tau, enable_highpass=enable_highpass, highpass_hz=highpass_hz,
                           enable_lowpass=enable_lowpass, lowpass_hz=lowpass_hz,
                           notch_frequencies=notch_frequencies)
    return _fm_demod(iq, sample_rate, audio_rate, stages, noise_blanker_threshold_db if enable_noise_blanker else None,
                     noise_reduction_db if enable_noise_reduction else None)
