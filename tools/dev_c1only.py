import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults, process_channels_batch
fs, n, B = 2_400_000, 120_000, 64
x = torch.view_as_complex(torch.randn((B, n, 2), device="cuda") * 0.5)
cfg = apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=200000.0))
for _ in range(3):
    process_channels_batch(x, fs, [cfg], n_chunks=B, return_device=True)
torch.cuda.synchronize()
