#!/usr/bin/env python
"""Voice-channel DiscriminatorBank, C channels x n audio samples with per-channel symbol phases (the channels of a site
are not symbol-aligned): a few demodulate() calls for timing / the ncu launch list, and a digest of the dibits so two
library builds can be compared on the same box:
python tools/dev_discdemod.py [C] [n] [iters] [aligned]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.decoders.p25 import DiscriminatorBank
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 72000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
aligned = len(sys.argv) > 4 and sys.argv[4] == "aligned"
rng = np.random.default_rng(7)
sps = 10
nsym = n // sps + 8
h = np.hanning(2 * sps + 1); h /= h.sum() / sps
rows = []
for c in range(C):
    lev = rng.choice(np.array([-3.0, -1.0, 1.0, 3.0]), size=nsym) * 0.2
    up = np.zeros(nsym * sps); up[::sps] = lev
    a = np.convolve(up, h, "same")
    d = 0 if aligned else int(rng.integers(0, sps))
    rows.append((a[d:d + n] + 0.01 * rng.standard_normal(n) + 0.05).astype(np.float32))
x = torch.from_numpy(np.stack(rows)).cuda()
bank = DiscriminatorBank(C, 48000)
outs = []
for _ in range(2):
    outs.append(bank.demodulate(x))
torch.cuda.synchronize()
hsh = hashlib.sha256()
for dib, soft, cnt in outs:
    hsh.update(dib.cpu().numpy().tobytes()); hsh.update(soft.cpu().numpy().tobytes()); hsh.update(cnt.cpu().numpy().tobytes())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    bank.demodulate(x)
e1.record(); torch.cuda.synchronize()
print(f"discdemod C={C} n={n} {'aligned' if aligned else 'staggered'}: ms per demodulate {e0.elapsed_time(e1) / iters:.3f}  "
      f"symbols/ch {int(outs[0][2][0])}  digest {hsh.hexdigest()[:16]}")
