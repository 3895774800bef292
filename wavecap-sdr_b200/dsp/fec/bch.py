"""BCH(63,16,23) decoding of the P25 Network ID on the GPU.

Call surface of `wavecapsdr.dsp.fec.bch` (dsp/fec/bch.py:225-658): `bch_decode(codeword, tracked_nac=None)` and
`BCH_63_16_23().decode(...)` return `(data16, errors)` with `errors == -1` for an uncorrectable word. The work is
done by `wc_bch_decode` (csrc/p25frame.cu: syndromes, Berlekamp-Massey, Chien search, re-check, second attempt with
the tracked NAC), one thread per codeword; `bch_decode_batch` is the form the GPU is meant for — all NID candidates
of all channels of a chunk in one launch. There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _native as N


def bch_decode_batch(codewords, tracked_nac=None) -> tuple[np.ndarray, np.ndarray]:
    """codewords: array-like [B][>=63] of 0/1 (first 63 used); tracked_nac: None, int or int array [B].
    Returns (data int32 [B], errors int32 [B])."""
    cw = np.asarray(codewords)
    if cw.ndim != 2 or cw.shape[1] < 63:
        raise ValueError("bch_decode_batch: expected [B][>=63] bits")
    bits = np.ascontiguousarray(cw[:, :63].astype(np.uint8) & 1)
    b = bits.shape[0]
    data = np.zeros(b, dtype=np.int32)
    errs = np.full(b, -1, dtype=np.int32)
    if b == 0:
        return data, errs
    tr = None
    if tracked_nac is not None:
        tr = np.ascontiguousarray(np.broadcast_to(np.asarray(tracked_nac, dtype=np.int32), (b,)))
    N.ensure_init()
    N.check(N.lib().wc_bch_decode_host(N.np_ptr(bits), N.np_ptr(tr) if tr is not None else None, b,
                                       N.np_ptr(data), N.np_ptr(errs)))
    return data, errs


class BCH_63_16_23:
    """Parameters and `decode` of the reference class (dsp/fec/bch.py:225-641)."""

    M = 6
    N = 63
    K = 16
    T = 11
    PRIMITIVE_POLYNOMIAL = 0x43
    MESSAGE_NOT_CORRECTED = -1

    def decode(self, codeword, tracked_nac: int | None = None) -> tuple[int, int]:
        cw = np.asarray(codeword)
        if cw.size < self.N:  # bch.py:585-587
            return 0, self.MESSAGE_NOT_CORRECTED
        t = tracked_nac if (tracked_nac is not None and tracked_nac > 0) else None
        d, e = bch_decode_batch(cw.reshape(1, -1), None if t is None else [t])
        return int(d[0]), int(e[0])


_decoder: BCH_63_16_23 | None = None


def bch_decode(codeword, tracked_nac: int | None = None) -> tuple[int, int]:
    """`wavecapsdr.dsp.fec.bch.bch_decode` (bch.py:644-658)."""
    global _decoder
    if _decoder is None:
        _decoder = BCH_63_16_23()
    return _decoder.decode(codeword, tracked_nac)
