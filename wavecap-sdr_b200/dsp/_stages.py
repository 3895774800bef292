"""Device-side stage operators (thin Python over the C ABI) used by dsp/fm.py, dsp/am.py, dsp/agc.py,
dsp/filters.py and capture.py. torch is only the device allocator / stream provider here.

Filter DESIGN (butter / iirnotch / firwin) stays on the host with scipy, exactly where the reference
does it (dsp/fm.py:143, dsp/filters.py:59-61,82; scipy.signal.resample_poly's internal firwin);
filter EXECUTION is on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache

import numpy as np

from .. import _native as N

MODE_NONE, MODE_WBFM, MODE_NBFM, MODE_AM, MODE_SSB, MODE_RAW = 0, 1, 2, 3, 4, 5
FMT_CF32, FMT_CS16 = 0, 1
EPI_NONE, EPI_RMS_CLIP, EPI_CLIP, EPI_RMS, EPI_CLIP_AGC = 0, 1, 2, 3, 4
OP_SOFT_CLIP_FM, OP_SOFT_CLIP_AGC, OP_SCALE = 0, 1, 2


def _torch():
    import torch

    return torch


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream() -> C.c_void_p:
    return N.torch_stream_ptr()


def to_device(x, np_dtype):
    """numpy array or CUDA tensor -> contiguous CUDA tensor of the given numpy dtype."""
    torch = _torch()
    N.ensure_init()
    if N.is_torch_cuda(x):
        want = {np.dtype(np.float32): torch.float32, np.dtype(np.complex64): torch.complex64,
                np.dtype(np.int16): torch.int16}[np.dtype(np_dtype)]
        return x.to(want).contiguous()
    a = np.ascontiguousarray(x, dtype=np_dtype)
    return torch.from_numpy(a).cuda()


def like_input(result, reference_input):
    """numpy in -> numpy out; CUDA tensor in -> CUDA tensor out."""
    if N.is_torch_cuda(reference_input):
        return result
    return result.cpu().numpy()


class _Handle:
    def __init__(self, h: C.c_void_p, destroy):
        self.h, self._destroy = h, destroy

    def __del__(self):
        try:
            if self.h and self.h.value:
                self._destroy(self.h)
        except Exception:
            pass


@lru_cache(maxsize=256)
def iir_handle(b: tuple, a: tuple) -> _Handle:
    N.ensure_init()
    bb = np.asarray(b, dtype=np.float64)
    aa = np.asarray(a, dtype=np.float64)
    h = C.c_void_p()
    N.check(N.lib().wc_iir_create(N.np_ptr(bb), bb.size, N.np_ptr(aa), aa.size, C.byref(h)))
    return _Handle(h, N.lib().wc_iir_destroy)


def resample_taps(up: int, down: int) -> np.ndarray:
    """The FIR scipy.signal.resample_poly designs internally (scipy 1.18 `resample_poly`):
    half_len = 10*max(up,down); firwin(2*half_len+1, 1/max(up,down), window=("kaiser", 5.0)) * up."""
    from scipy.signal import firwin

    mx = max(up, down)
    half_len = 10 * mx
    return firwin(2 * half_len + 1, 1.0 / mx, window=("kaiser", 5.0)) * up


@lru_cache(maxsize=64)
def resampler_handle(up: int, down: int) -> _Handle:
    N.ensure_init()
    taps = np.ascontiguousarray(resample_taps(up, down), dtype=np.float64)
    h = C.c_void_p()
    N.check(N.lib().wc_resampler_create(int(up), int(down), N.np_ptr(taps), taps.size, C.byref(h)))
    return _Handle(h, N.lib().wc_resampler_destroy)


def rate_ratio(in_rate: int, out_rate: int) -> tuple[int, int]:
    g = math.gcd(int(in_rate), int(out_rate))  # dsp/fm.py:207-209
    return int(out_rate) // g, int(in_rate) // g


# ---- stage calls (2-D tensors [n_seq, n], contiguous) ----------------------------------------------

def lfilter(b, a, x, abs_input: bool = False, out=None):
    torch = _torch()
    assert x.dim() == 2 and x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty_like(x) if out is None else out
    if x.numel() == 0:
        return y
    h = iir_handle(tuple(float(v) for v in np.atleast_1d(b)), tuple(float(v) for v in np.atleast_1d(a)))
    N.check(N.lib().wc_iir_lfilter(h.h, ptr(x), ptr(y), x.shape[1], x.shape[1], x.shape[0], int(abs_input), stream()))
    return y


def sumsq(x):
    torch = _torch()
    out = torch.empty((x.shape[0],), dtype=torch.float64, device=x.device)
    N.check(N.lib().wc_sumsq(ptr(x), x.shape[1], x.shape[1], x.shape[0], ptr(out), stream()))
    return out


def elementwise(x, op: int, p0: float = 0.0):
    torch = _torch()
    y = torch.empty_like(x)
    N.check(N.lib().wc_elementwise(ptr(x), ptr(y), x.numel(), op, float(p0), stream()))
    return y


def agc_apply(x, env_a, env_r, target_linear: float, max_gain_linear: float):
    torch = _torch()
    y = torch.empty_like(x)
    N.check(N.lib().wc_agc_apply(ptr(x), ptr(env_a), ptr(env_r), ptr(y), x.numel(), float(target_linear),
                                 float(max_gain_linear), stream()))
    return y


def resample(x, up: int, down: int, epilogue: int = EPI_NONE, sumsq_dev=None, target_rms: float = 0.18,
             min_rms: float = 1e-4, want_stats: bool = False, max_abs: float = 1.2):
    """x [n_seq, n_in] float32 -> [n_seq, ceil(n_in*up/down)] float32 (+ optional power / invalid)."""
    torch = _torch()
    h = resampler_handle(int(up), int(down))
    n_seq, n_in = x.shape
    n_out = int(N.lib().wc_resampler_out_len(h.h, n_in))
    out = torch.empty((n_seq, n_out), dtype=torch.float32, device=x.device)
    power = torch.zeros((n_seq,), dtype=torch.float64, device=x.device) if want_stats else None
    invalid = torch.zeros((n_seq,), dtype=torch.int32, device=x.device) if want_stats else None
    if n_seq and n_out:
        N.check(N.lib().wc_resampler_run(h.h, ptr(x), n_in, n_in, n_seq, ptr(out), epilogue, ptr(sumsq_dev),
                                         float(target_rms), float(min_rms), ptr(power), ptr(invalid),
                                         float(max_abs), stream()))
    return (out, power, invalid) if want_stats else out


def front(iq, fmt: int, n: int, n_chunks: int, modes, offsets_hz, bfo_hz, sample_rate: int,
          want_out: bool = True, want_base: bool = False, want_sumsq: bool = False):
    """Frequency shift + RSSI power + demod front end for all channels of all chunks.
    Returns (out [C,B,n] f32 | None, base [C,B,n] c64 | None, power [C,B] f64, nonfinite [B] i32)
    (+ sum(out**2) [C,B] f64 as a fifth item when want_sumsq)."""
    torch = _torch()
    n_ch = len(modes)
    dev = iq.device
    out = torch.empty((n_ch, n_chunks, n), dtype=torch.float32, device=dev) if want_out else None
    base = torch.empty((n_ch, n_chunks, n), dtype=torch.complex64, device=dev) if want_base else None
    power = torch.empty((n_ch, n_chunks), dtype=torch.float64, device=dev)
    nonfinite = torch.empty((n_chunks,), dtype=torch.int32, device=dev)
    scratch = torch.empty((int(N.lib().wc_front_chan_scratch_bytes(n_ch)),), dtype=torch.uint8, device=dev)
    m = np.ascontiguousarray(modes, dtype=np.int32)
    o = np.ascontiguousarray(offsets_hz, dtype=np.float64)
    b = np.ascontiguousarray(bfo_hz if bfo_hz is not None else np.zeros(n_ch), dtype=np.float64)
    ss = torch.empty((n_ch, n_chunks), dtype=torch.float64, device=dev) if want_sumsq else None
    N.check(N.lib().wc_front_run_ex(ptr(iq), fmt, n, n_chunks, n, n_ch, N.np_ptr(m), N.np_ptr(o), N.np_ptr(b),
                                    int(sample_rate), ptr(out), ptr(base), ptr(power), ptr(ss), ptr(nonfinite),
                                    ptr(scratch), stream()))
    if want_sumsq:
        return out, base, power, nonfinite, ss
    return out, base, power, nonfinite
