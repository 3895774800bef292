"""Dev: spectrum variant 5 (two-stream slab pipeline) vs the default, 4096 frames of 65536, K = 4."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavecap_sdr_b200._native as N
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend
N.init(0)
frames = 4096
x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
ref = None
for var, mb in [(3, 512)] + [(5, int(v)) for v in os.environ.get("MBS", "64,96,128,192,256,384,512").split(",")]:
    os.environ["WC_SPECTRUM_VARIANT"] = str(var)
    os.environ["WC_SPECTRUM_PIPE_SLAB_MB" if var == 5 else "WC_SPECTRUM_SLAB_MB"] = str(mb)
    be = CudaFFTBackend(65536)
    out = be.execute_frames(x, frames, 65536, 4)
    for _ in range(2): be.execute_frames(x, frames, 65536, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): be.execute_frames(x, frames, 65536, 4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    o = out if torch.is_tensor(out) else torch.as_tensor(out)
    if ref is None: ref = o.clone()
    print(json.dumps({"variant": var, "slab_mb": mb, "ms": round(ms, 4), "GS/s": round(frames * 65536 / ms / 1e6, 1),
                      "max_abs_diff_vs_first_dB": float((o - ref).abs().max())}), flush=True)
    del be
