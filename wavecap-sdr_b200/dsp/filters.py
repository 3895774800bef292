"""GPU filters with the call surface of `wavecapsdr.dsp.filters` (dsp/filters.py).

IIR section (:41-264): Butterworth order-5 high/low/band-pass and iirnotch, executed by the float64
block-scan `lfilter` in csrc/analog.cu. Coefficients are designed on the host with scipy exactly as the
reference does (:59-61, :82) and cached. Invalid cut-offs return the input unchanged as float32.
Streaming FIR section (:471-668): `fir_filter_complex`, `fir_decimate` — csrc/firdec.cu.
`noise_blanker` (:267-343) and `spectral_noise_reduction` (:346-459) — csrc/audiofx.cu (real float32 input; the
default fft_size=1024 / overlap=0.5 of the reference are the only STFT geometry built).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .. import _native as N
from . import _stages as S

NUMBA_AVAILABLE = False  # the reference's flag; nothing here uses numba


@lru_cache(maxsize=128)
def _butter(btype: str, cutoff: tuple, order: int):
    from scipy import signal

    wn = cutoff[0] if len(cutoff) == 1 else list(cutoff)
    b, a = signal.butter(order, wn, btype=btype)
    return tuple(b), tuple(a)


@lru_cache(maxsize=64)
def _notch(w0: float, q: float):
    from scipy import signal

    b, a = signal.iirnotch(w0, q)
    return tuple(b), tuple(a)


def _as_rows(x):
    t = S.to_device(x, np.float32)
    return t.reshape(1, -1) if t.dim() == 1 else t.reshape(-1, t.shape[-1])


def _run_iir(x, coeffs):
    if coeffs is None or _size(x) == 0:
        return _f32_passthrough(x)
    b, a = coeffs
    rows = _as_rows(x)
    y = S.lfilter(b, a, rows).reshape(_shape(x))
    return S.like_input(y, x)


def _size(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def _shape(x):
    return tuple(x.shape)


def _f32_passthrough(x):
    if hasattr(x, "is_cuda"):
        import torch

        return x.to(torch.float32)
    return np.asarray(x).astype(np.float32, copy=False)


def highpass_coeffs(sample_rate: int, cutoff: float, order: int = 5):
    wn = cutoff / (sample_rate / 2.0)
    return None if (wn <= 0 or wn >= 1.0) else _butter("high", (wn,), order)


def lowpass_coeffs(sample_rate: int, cutoff: float, order: int = 5):
    wn = cutoff / (sample_rate / 2.0)
    return None if (wn <= 0 or wn >= 1.0) else _butter("low", (wn,), order)


def bandpass_coeffs(sample_rate: int, low: float, high: float, order: int = 5):
    lo, hi = low / (sample_rate / 2.0), high / (sample_rate / 2.0)
    return None if (lo <= 0 or hi >= 1.0 or lo >= hi) else _butter("band", (lo, hi), order)


def notch_coeffs(sample_rate: int, freq: float, q: float = 30.0):
    w0 = freq / (sample_rate / 2.0)
    return None if (w0 <= 0 or w0 >= 1.0) else _notch(w0, q)


def highpass_filter(x, sample_rate: int, cutoff: float, order: int = 5):
    """dsp/filters.py:85-126."""
    return _run_iir(x, highpass_coeffs(sample_rate, cutoff, order))


def lowpass_filter(x, sample_rate: int, cutoff: float, order: int = 5):
    """dsp/filters.py:129-172."""
    return _run_iir(x, lowpass_coeffs(sample_rate, cutoff, order))


def bandpass_filter(x, sample_rate: int, low: float, high: float, order: int = 5):
    """dsp/filters.py:175-221."""
    return _run_iir(x, bandpass_coeffs(sample_rate, low, high, order))


def notch_filter(x, sample_rate: int, freq: float, q: float = 30.0):
    """dsp/filters.py:224-264."""
    return _run_iir(x, notch_coeffs(sample_rate, freq, q))


def noise_blanker_rows(rows, threshold_db: float = 10.0, blanking_width: int = 3):
    """rows: CUDA float32 [n_seq, n] -> blanked rows (one median + one masked copy per sequence)."""
    import torch

    rows = rows.contiguous()
    out = torch.empty_like(rows)
    N.check(N.lib().wc_noise_blanker(S.ptr(rows), S.ptr(out), int(rows.shape[1]), int(rows.shape[1]), int(rows.shape[0]),
                                     float(threshold_db), int(blanking_width), S.stream()))
    return out


def noise_blanker(x, threshold_db: float = 10.0, blanking_width: int = 3):
    """dsp/filters.py:267-343: zero every sample within `blanking_width` of a sample whose magnitude exceeds the
    median magnitude by `threshold_db`."""
    if _size(x) == 0:
        return _f32_passthrough(x)
    if (hasattr(x, "is_complex") and x.is_complex()) or (isinstance(x, np.ndarray) and np.iscomplexobj(x)):
        raise NotImplementedError("noise_blanker: only the real (post-discriminator) path of the FM chains is built")
    y = noise_blanker_rows(_as_rows(x), threshold_db, blanking_width).reshape(_shape(x))
    return S.like_input(y, x)


def spectral_nr_rows(rows, reduction_db: float = 12.0):
    """rows: CUDA float32 [n_seq, n] -> [n_seq, out_len] (out_len = samples covered by whole STFT frames)."""
    import torch

    rows = rows.contiguous()
    n = int(rows.shape[1])
    m = int(N.lib().wc_spectral_nr_out_len(n))
    out = torch.empty((rows.shape[0], m), dtype=torch.float32, device=rows.device)
    N.check(N.lib().wc_spectral_nr(S.ptr(rows), n, n, int(rows.shape[0]), float(reduction_db), S.ptr(out), m, S.stream()))
    return out


def spectral_noise_reduction(x, sample_rate: int, reduction_db: float = 12.0, fft_size: int = 1024, overlap: float = 0.5):
    """dsp/filters.py:346-459 (Wiener-style spectral gain against the per-bin 10th-percentile noise floor)."""
    if fft_size != 1024 or overlap != 0.5:
        raise NotImplementedError("spectral_noise_reduction: only fft_size=1024, overlap=0.5 (the reference defaults) are built")
    if _size(x) == 0 or _size(x) < fft_size:
        return _f32_passthrough(x)
    y = spectral_nr_rows(_as_rows(x), reduction_db)
    return S.like_input(y.reshape(-1) if len(_shape(x)) == 1 else y, x)


def __getattr__(name):
    # streaming FIR entry points live in _firdec (csrc/firdec.cu) and are re-exported lazily
    if name in ("fir_filter_complex", "fir_decimate", "warmup_numba_filters"):
        from . import _firdec

        return getattr(_firdec, name)
    raise AttributeError(name)
