"""dev: time the exact-replay IIR (iir_seq_kernel) on 64 x 120000 and 1 x 120000 rows (100 Hz high-pass at 48 kS/s)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavecap_sdr_b200.dsp import _stages as S, filters as F

b, a = F.highpass_coeffs(48000, 100.0)
x = torch.randn(64, 120000, device="cuda")
for n_seq in (1, 8, 64):
    S.lfilter(b, a, x[:n_seq].contiguous())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        S.lfilter(b, a, x[:n_seq].contiguous())
    torch.cuda.synchronize()
    print(f"iir_seq K=5: {n_seq} rows x 120000: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
