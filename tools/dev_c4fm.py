#!/usr/bin/env python
"""C4FM bank, 64 channels x 72 000 samples: a few demodulate() calls for timing / the ncu launch list:
python tools/dev_c4fm.py [C] [n] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavecap_sdr_b200._native as N
N.init(0)
from oracle.c4fm import modulate_c4fm, random_frames
from wavecap_sdr_b200.dsp.p25.c4fm import C4FMBank
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 72000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(1); torch.manual_seed(0)
base = modulate_c4fm(random_frames(rng, n_frames=(n // 2140) + 2, payload=150, gap=40), 48000, seed=1)[:n]
x = torch.from_numpy(np.ascontiguousarray(np.tile(base, (C, 1)))).cuda()
x = x * torch.exp(1j * torch.rand((C, 1), device="cuda") * 6.28).to(torch.complex64)
bank = C4FMBank(C, 48000)
import hashlib
hsh = hashlib.sha256()
for _ in range(2):
    for t in bank.demodulate(x):
        if torch.is_tensor(t):
            hsh.update(t.cpu().numpy().tobytes())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    bank.demodulate(x)
e1.record(); torch.cuda.synchronize()
print("c4fm ms per demodulate:", e0.elapsed_time(e1) / iters, "digest", hsh.hexdigest()[:16])
