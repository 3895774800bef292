"""P25 1/2-rate trellis (Viterbi) decoding on the GPU.

Call surface of `wavecapsdr.dsp.fec.trellis` (dsp/fec/trellis.py:89-329): `trellis_decode(dibits, soft_values=None)` and
`TrellisDecoder().decode(...)` return `(decoded_dibits uint8, error_metric int)`. The work is `wc_trellis12_decode`
(csrc/p25frame.cu), one thread per block; `trellis_decode_batch` decodes many blocks per launch and `tsbk_decode_batch`
is the TSBK block decode of `decoders/p25.py:2037-2109` (196 message bits -> deinterleave -> decode -> 96 bits and the
header fields) for all TSBK messages a framer bank produced. There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _native as N

MAX_DIBITS = 2048


def trellis_decode_batch(dibits, soft_values=None, lengths=None):
    """dibits [B][n] (0-3); soft_values None or float [B][n]; lengths None or int [B] (<= n).
    Returns (decoded uint8 [B][n//2], n_out int32 [B], metric int32 [B])."""
    import torch

    N.ensure_init()
    d = np.ascontiguousarray(np.asarray(dibits).astype(np.uint8))
    if d.ndim != 2:
        raise ValueError("trellis_decode_batch: expected [B][n] dibits")
    b, n = d.shape
    if n > MAX_DIBITS:
        raise ValueError(f"trellis_decode_batch: blocks longer than {MAX_DIBITS} dibits are not supported")
    out = torch.zeros((b, max(1, n // 2)), dtype=torch.uint8, device="cuda")
    n_out = torch.zeros((b,), dtype=torch.int32, device="cuda")
    met = torch.zeros((b,), dtype=torch.int32, device="cuda")
    if b == 0 or n == 0:
        return out.cpu().numpy()[:, :0], n_out.cpu().numpy(), met.cpu().numpy()
    dd = torch.from_numpy(d).cuda()
    sv = None if soft_values is None else torch.from_numpy(np.ascontiguousarray(soft_values, dtype=np.float64)).cuda()
    ln = None if lengths is None else torch.from_numpy(np.ascontiguousarray(lengths, dtype=np.int32)).cuda()
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    N.check(N.lib().wc_trellis12_decode(vp(dd), n, vp(ln), n, vp(sv), b, vp(out), out.shape[1], vp(n_out), vp(met),
                                        N.torch_stream_ptr()))
    return out.cpu().numpy(), n_out.cpu().numpy(), met.cpu().numpy()


def tsbk_decode_batch(bits196):
    """bits196 [B][>=196] of 0/1 (P25P1Message.bits of TSBK blocks). Returns (bits96 uint8 [B][96], metric int32 [B],
    fields int32 [B][4] = last_block, protected, opcode, mfid, data uint8 [B][8])."""
    a = np.asarray(bits196)
    if a.ndim != 2 or a.shape[1] < 196:
        raise ValueError("tsbk_decode_batch: expected [B][>=196] bits")
    bits = np.ascontiguousarray(a[:, :196].astype(np.uint8) & 1)
    b = bits.shape[0]
    out = np.zeros((b, 96), dtype=np.uint8)
    met = np.zeros(b, dtype=np.int32)
    fields = np.zeros((b, 4), dtype=np.int32)
    data = np.zeros((b, 8), dtype=np.uint8)
    if b:
        N.ensure_init()
        N.check(N.lib().wc_tsbk_decode_host(N.np_ptr(bits), b, N.np_ptr(out), N.np_ptr(met), N.np_ptr(fields), N.np_ptr(data)))
    return out, met, fields, data


class TrellisDecoder:
    """`decode` of the reference class (dsp/fec/trellis.py:214-272); the decoder is reset per call there too."""

    NUM_STATES = 4
    TRACEBACK_DEPTH = 12

    def reset(self) -> None:
        pass

    def decode(self, dibits, soft_values=None, debug: bool = False):
        d = np.asarray(dibits)
        if d.size == 0:
            return np.array([], dtype=np.uint8), 0
        sv = None if soft_values is None else np.asarray(soft_values, dtype=np.float64).reshape(1, -1)
        out, n_out, met = trellis_decode_batch(d.reshape(1, -1), sv)
        return out[0, : int(n_out[0])].copy(), int(met[0])


def trellis_decode(dibits, soft_values=None, debug: bool = False):
    """`wavecapsdr.dsp.fec.trellis.trellis_decode` (trellis.py:312-329)."""
    return TrellisDecoder().decode(dibits, soft_values, debug=debug)
