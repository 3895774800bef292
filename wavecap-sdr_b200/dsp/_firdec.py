"""Streaming complex FIR entry points of `wavecapsdr.dsp.filters` on the GPU.

Mirrors wavecapsdr/dsp/filters.py:558-668: `fir_filter_complex(x, taps, zi, use_parallel)`,
`fir_decimate(x, taps, decim_factor, zi)`, `warmup_numba_filters()`. The state `zi` is the reference's:
the last len(taps)-1 inputs as complex128. Arithmetic: csrc/firdec.cu (`wc_fir_complex`), float64
accumulation, only kept outputs computed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _native as N
from . import _stages as S

NUMBA_AVAILABLE = False


def _run(x, taps, decim: int, zi):
    taps64 = np.ascontiguousarray(np.asarray(taps), dtype=np.float64)
    n_zi = len(taps64) - 1
    n = S_n(x)
    if n == 0:
        empty_zi = zi if zi is not None else np.zeros(n_zi, dtype=np.complex128)
        return np.empty(0, dtype=np.complex64), empty_zi
    N.ensure_init()
    import torch

    xd = S.to_device(x, np.complex64).reshape(-1)
    n_out = (n + decim - 1) // decim
    yd = torch.empty((n_out,), dtype=torch.complex64, device=xd.device)
    zin = None if zi is None else np.ascontiguousarray(np.asarray(zi), dtype=np.complex128)
    zout = np.zeros(max(n_zi, 1), dtype=np.complex128)
    N.check(N.lib().wc_fir_complex(S.ptr(xd), n, N.np_ptr(taps64), len(taps64), int(decim),
                                   None if zin is None else N.np_ptr(zin), S.ptr(yd), N.np_ptr(zout), S.stream()))
    return S.like_input(yd, x), zout[:n_zi]


def S_n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def fir_filter_complex(x, taps, zi=None, use_parallel: bool = True):
    """filters.py:558-620 -> (complex64 output, complex128 state)."""
    return _run(x, taps, 1, zi)


def fir_decimate(x, taps, decim_factor: int, zi=None):
    """filters.py:623-652: filter, then keep every decim_factor-th sample of this call."""
    return _run(x, taps, int(decim_factor), zi)


def warmup_numba_filters() -> None:
    """filters.py:659-668 pre-compiles numba kernels; here it just makes sure the CUDA library is loaded."""
    N.ensure_init()
