"""Dev timing of the spectrum passes (not the bench contract): WC_SPECTRUM_PASS_A variants at 4096 frames."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavecap_sdr_b200._native as N
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend
N.init(0)
be = CudaFFTBackend(65536)
frames = 4096
x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
ref = None
for v in os.environ.get("VALS", "3,5,6").split(","):
    os.environ["WC_SPECTRUM_PASS_A"] = v
    out = be.execute_frames(x, frames, 65536, 4)
    for _ in range(2): be.execute_frames(x, frames, 65536, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): be.execute_frames(x, frames, 65536, 4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    o = out if torch.is_tensor(out) else torch.as_tensor(out)
    if ref is None: ref = o.clone()
    err = float((o - ref).abs().max())
    print(json.dumps({"pass_a": v, "ms": round(ms, 4), "GS/s": round(frames * 65536 / ms / 1e6, 1), "max_abs_diff_vs_first_dB": err}), flush=True)
