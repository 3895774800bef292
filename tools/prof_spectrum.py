"""Small driver for profiling the spectrum kernels under ncu (368 frames = 8 chunks of config C3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavecap_sdr_b200._native as N
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend

N.init(0)
frames = int(os.environ.get("FRAMES", "368"))
be = CudaFFTBackend(65536)
x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
for _ in range(int(os.environ.get("ITERS", "3"))):
    out = be.execute_frames(x, frames, 65536, 4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = be.execute_frames(x, frames, 65536, 4)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print({"frames": frames, "ms": ms, "GS/s": frames * 65536 / ms / 1e6, "checksum": float(out.double().sum())})
