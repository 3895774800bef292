#!/usr/bin/env python
"""CQPSK bank, 64 channels x 72 000 samples: per-kernel times (CUDA events around each launch via the launch list under ncu)
or just a few calls for profiling: python tools/dev_cqpsk.py [C] [n] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavecap_sdr_b200._native as N
N.init(0)
from oracle.cqpsk import modulate_cqpsk
from wavecap_sdr_b200.decoders.p25 import CQPSKBank
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 72000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(1)
cb = modulate_cqpsk(rng.integers(0, 4, n // 10 + 8), 48000, 4800, seed=2)[:n]
xq = torch.from_numpy(np.ascontiguousarray(np.tile(cb, (C, 1)))).cuda()
qb = CQPSKBank(C, 48000)
import hashlib
hsh = hashlib.sha256()
for _ in range(2):
    r = qb.demodulate(xq)
    for t in (r if isinstance(r, (tuple, list)) else (r,)):
        if torch.is_tensor(t):
            hsh.update(t.cpu().numpy().tobytes())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    qb.demodulate(xq)
e1.record(); torch.cuda.synchronize()
print("cqpsk ms per demodulate:", e0.elapsed_time(e1) / iters, "digest", hsh.hexdigest()[:16])
