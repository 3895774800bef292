"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's P25 1/2-rate trellis (Viterbi) decoder and TSBK block
decode. Source: /root/reference/backend/wavecapsdr/dsp/fec/trellis.py
  encoder table        :31-66    output nibbles {2,12,1,15},{14,0,13,3},{9,7,10,4},{5,11,6,8}; next state = input dibit
  branch metric        :122-150  hard: bit errors of the two dibits; soft: squared distance to the +-1/+-3 levels
  decode_step          :152-212  survivors per next state (strict <, previous states in order), streaming output of the
                                 CURRENT best path's decision 12 steps back once 12 steps have been taken
  decode / _flush      :214-289  reset, step over dibit pairs, then the last 11 decisions of the FINAL best path
and decoders/p25.py:2037-2087,2552-2775 (TSBK: 196 bits -> 98 dibits -> deinterleave -> decode -> 48 dibits -> 96 bits).
The survivor bookkeeping is array-based here (metric[4], path[4]); the observable behaviour — including the mix of
mid-stream and final-path decisions and the first-minimum tie-breaks — is the reference's. Pinned to the live reference by
tests/golden/p25_trellis.npz (oracle/make_golden.py:gen_p25_trellis).
"""
from __future__ import annotations

import numpy as np

OUT_NIBBLE = ((2, 12, 1, 15), (14, 0, 13, 3), (9, 7, 10, 4), (5, 11, 6, 8))   # [state][input]
LEVEL = (1.0, 3.0, -1.0, -3.0)
DEPTH = 12
# decoders/p25.py:2552-2660: 12 groups of four dibit pairs taken from the quarters of the block, then the odd pair out
DEINTERLEAVE = np.array([b + 2 * i + k for i in range(12) for b in (0, 26, 50, 74) for k in (0, 1)] + [24, 25], dtype=np.int64)


def encode(dibits) -> np.ndarray:
    state, out = 0, []
    for d in dibits:
        d = int(d) & 3
        nib = OUT_NIBBLE[state][d]
        out += [nib >> 2, nib & 3]
        state = d
    return np.array(out, dtype=np.uint8)


def interleave(block98) -> np.ndarray:
    """inverse of the reference's gather: interleaved[DEINTERLEAVE[i]] = block[i]."""
    out = np.zeros(98, dtype=np.uint8)
    out[DEINTERLEAVE] = np.asarray(block98, dtype=np.uint8)
    return out


def _branch(rx, nib, soft):
    e0, e1 = nib >> 2, nib & 3
    if soft is not None:
        return (soft[0] - LEVEL[e0]) ** 2 + (soft[1] - LEVEL[e1]) ** 2
    return float(bin(rx[0] ^ e0).count("1") + bin(rx[1] ^ e1).count("1"))


def decode(dibits, soft_values=None):
    """TrellisDecoder.decode: returns (decoded dibits uint8, int error metric)."""
    dibits = np.asarray(dibits)
    if len(dibits) % 2:
        dibits = dibits[:-1]
        if soft_values is not None:
            soft_values = soft_values[:-1]
    metric = [0.0, float("inf"), float("inf"), float("inf")]
    path = [[], [], [], []]
    out = []
    for i in range(0, len(dibits), 2):
        rx = (int(dibits[i]), int(dibits[i + 1]))
        soft = None if soft_values is None else (float(soft_values[i]), float(soft_values[i + 1]))
        nm, npth = [], []
        for ns in range(4):
            best, bp = float("inf"), 0
            for p in range(4):
                m = metric[p] + _branch(rx, OUT_NIBBLE[p][ns], soft)
                if m < best:
                    best, bp = m, p
            # when every candidate is +inf the reference keeps input 0 from previous state 0
            nm.append(best)
            npth.append(path[bp] + [ns if best < float("inf") else 0])
        metric, path = nm, npth
        steps = i // 2 + 1
        if steps >= DEPTH:
            b = min(range(4), key=lambda s: metric[s])
            out.append(path[b][steps - DEPTH])
    b = min(range(4), key=lambda s: metric[s])
    out += path[b][max(0, len(path[b]) - DEPTH + 1):]
    m = metric[b]
    return np.array(out, dtype=np.uint8), int(m)


def tsbk_decode_bits(bits196):
    """decoders/p25.py:2037-2087: (96 decoded bits uint8, error metric) or (None, -1)."""
    bits = np.asarray(bits196).astype(np.int64)
    if len(bits) < 196:
        return None, -1
    d = ((bits[0:196:2] << 1) | bits[1:196:2]).astype(np.uint8)
    dec, err = decode(d[DEINTERLEAVE])
    if len(dec) == 0:
        return None, -1
    dec = dec[:48]
    if err < 0 or len(dec) < 48:
        return None, err
    out = np.zeros(96, dtype=np.uint8)
    out[0::2] = (dec >> 1) & 1
    out[1::2] = dec & 1
    return out, err


def tsbk_fields(bits96):
    """decoders/p25.py:2089-2109: (last_block, protected, opcode, mfid, 8 data bytes)."""
    b = [int(v) for v in bits96]
    val = lambda s: int("".join(map(str, s)), 2)
    return b[0], b[1], val(b[2:8]), val(b[8:16]), bytes(val(b[16 + 8 * k:24 + 8 * k]) for k in range(8))
