import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults, process_channels_batch
fs, n, B = 10_000_000, 500_000, 8
q = torch.randint(-2000, 2000, (B, n, 2), device="cuda", dtype=torch.int16)
cfgs = []
for i in range(16):
    c = apply_mode_defaults("nbfm", ChannelConfig(id=str(i), capture_id="c", mode="nbfm", offset_hz=-3.75e6 + 5e5 * i))
    c.squelch_db = -45.0
    cfgs.append(c)
for _ in range(3):
    process_channels_batch(q, fs, cfgs, n_chunks=B, in_fmt="cs16", apply_squelch=True, return_device=True)
torch.cuda.synchronize()
