#!/usr/bin/env python
"""Minimax (relative error) odd polynomials atan(t) ~= t * P(t^2) on [0, 1] for the channelizer's fused
FM discriminator (csrc/common.cuh fast_atan2f*). Solved as a linear program on a dense grid, coefficients
rounded to float32 and the achieved error re-measured in float32 Horner arithmetic.

    python tools/fit_atan.py [n_coeffs ...]      # default 5 6 7
"""
import sys

import numpy as np
from scipy.optimize import linprog


def fit(nc: int):
    t = np.concatenate([np.linspace(1e-6, 1.0, 20001), 1 - np.logspace(-6, -1, 200)])
    s = t * t
    f = np.arctan(t) / t
    A = np.stack([s ** k for k in range(nc)], axis=1)
    # minimise E subject to |A c - f| <= E f
    c_obj = np.zeros(nc + 1)
    c_obj[-1] = 1.0
    Aub = np.vstack([np.hstack([A, -f[:, None]]), np.hstack([-A, -f[:, None]])])
    bub = np.concatenate([f, -f])
    r = linprog(c_obj, A_ub=Aub, b_ub=bub, bounds=[(None, None)] * nc + [(0, None)], method="highs")
    c = r.x[:nc].astype(np.float32)
    tt = np.linspace(0, 1, 400001).astype(np.float32)
    ss = (tt * tt).astype(np.float32)
    p = np.full_like(tt, c[-1])
    for k in range(nc - 2, -1, -1):
        p = (p * ss + c[k]).astype(np.float32)
    a = (p * tt).astype(np.float32)
    ref = np.arctan(tt.astype(np.float64))
    rel = np.abs(a - ref)[1:] / ref[1:]
    return c, float(r.x[-1]), float(rel.max()), float(np.abs(a - ref).max())


if __name__ == "__main__":
    for nc in [int(v) for v in sys.argv[1:]] or [5, 6, 7]:
        c, e_lp, e_rel, e_abs = fit(nc)
        print(f"degree {2 * nc - 1}: LP rel err {e_lp:.3e}, float32 Horner max rel {e_rel:.3e}, max abs {e_abs:.3e}")
        print("   coefficients (s^0 .. s^%d): " % (nc - 1) + ", ".join("%.9gf" % v for v in c))
