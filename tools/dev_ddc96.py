import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.trunking import DDCBank
fs, n, K = 6_000_000, 300_000, 96
x = torch.view_as_complex(torch.randn((n, 2), device="cuda") * 0.1)
b = DDCBank(K, fs, 30, 4)
b.set_offsets(np.linspace(-2.9e6, 2.9e6, K))
for _ in range(4):
    b.process(x)
torch.cuda.synchronize()
