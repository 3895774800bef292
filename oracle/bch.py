"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's BCH(63,16,23) NID decoder.

Follows /root/reference/backend/wavecapsdr/dsp/fec/bch.py step by step (plain Python integers, small enough):
  gf tables            bch.py:255-271
  syndromes            bch.py:195-222   S_s = XOR over set bits of alpha^((s+1)*(62-bit_pos))
  Berlekamp-Massey     bch.py:52-127    locator truncated to T+1 coefficients
  Chien search         bch.py:131-191   root alpha^i -> error_pos (63-i)%63, stop after `degree` roots
  decode_internal      bch.py:575-641   position inversion (62-pos)%63, re-check of the syndromes
  decode (two passes)  bch.py:533-573   second attempt with the NAC field overwritten by the tracked NAC
Pinned to the live reference by tests/golden/p25_framer.npz (oracle/make_golden.py:gen_p25_framer).
Also holds a systematic encoder (generator polynomial 6331141367235453 octal = lcm of the minimal polynomials of
alpha^1..alpha^22) used to synthesise valid NIDs; the reference has no encoder.
"""
from __future__ import annotations

import numpy as np

N, K, T = 63, 16, 11
GENERATOR = 0o6331141367235453

POW = [0] * 64
LOG = [0] * 64
_x = 1
for _i in range(N):
    POW[_i] = _x
    LOG[_x] = _i
    _x <<= 1
    if _x & 64:
        _x ^= 0x43
POW[N] = 1


def _mul(a: int, b: int) -> int:
    return 0 if (a == 0 or b == 0) else POW[(LOG[a] + LOG[b]) % N]


def syndromes(bits) -> list[int]:
    s = [0] * (2 * T)
    for pos in range(N):
        if bits[pos]:
            for j in range(2 * T):
                s[j] ^= POW[((j + 1) * (N - 1 - pos)) % N]
    return s


def berlekamp_massey(s) -> tuple[list[int], int]:
    c = [0] * (T + 1)
    b = [0] * (T + 1)
    c[0] = b[0] = 1
    L, m, log_b = 0, 1, 0
    for n in range(2 * T):
        d = s[n]
        for i in range(1, min(L + 1, T + 1)):
            if n >= i:
                d ^= _mul(c[i], s[n - i])
        if d == 0:
            m += 1
            continue
        keep = list(c)
        log_d = LOG[d]
        log_db = (log_d + N - log_b) % N
        for i in range(T + 1 - m):
            if b[i]:
                c[i + m] ^= POW[(LOG[b[i]] + log_db) % N]
        if n >= 2 * L:
            L, b, log_b, m = n + 1 - L, keep, log_d, 1
        else:
            m += 1
    return c, L


def chien(c, degree) -> list[int]:
    roots = []
    for i in range(N):
        v = 0
        for k in range(degree + 1):
            if c[k]:
                v ^= POW[(LOG[c[k]] + i * k) % N]
        if v == 0:
            roots.append((N - i) % N)
            if len(roots) >= degree:
                break
    return roots


def decode_once(bits) -> tuple[int, int]:
    bits = [int(v) & 1 for v in bits[:N]]
    s = syndromes(bits)
    if not any(s):
        return int("".join(map(str, bits[:K])), 2), 0
    c, L = berlekamp_massey(s)
    if L == 0 or L > T:
        return 0, -1
    roots = chien(c, L)
    if len(roots) != L:
        return 0, -1
    fixed = list(bits)
    for p in roots:
        fixed[(N - 1 - p) % N] ^= 1
    if any(syndromes(fixed)):
        return 0, -1
    return int("".join(map(str, fixed[:K])), 2), L


def bch_decode(bits, tracked_nac=None) -> tuple[int, int]:
    data, errs = decode_once(bits)
    if errs != -1:
        return data, errs
    if tracked_nac is not None and tracked_nac > 0:
        cur = int("".join(str(int(v) & 1) for v in bits[:12]), 2)
        if cur != tracked_nac:
            patched = [int(v) & 1 for v in bits[:N]]
            for i in range(12):
                patched[i] = (tracked_nac >> (11 - i)) & 1
            return decode_once(patched)
    return data, errs


def bch_encode(data16: int) -> np.ndarray:
    """Systematic codeword (63 bits, data first) for a 16-bit NAC|DUID value."""
    m = (data16 & 0xFFFF) << 47
    rem = m
    for bit in range(62, 46, -1):
        if (rem >> bit) & 1:
            rem ^= GENERATOR << (bit - 47)
    cw = m | rem
    return np.array([(cw >> (62 - i)) & 1 for i in range(N)], dtype=np.uint8)
