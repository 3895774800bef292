"""dev: the single-kernel group spectrum (WC_SPECTRUM_VARIANT=6, WC_DEV build) against the two-pass pipeline: equality and time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend

be = CudaFFTBackend(65536)
for frames in (4, 64, 368, 1024, 4096):
    x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
    res = {}
    for var in ("5", "3", "6"):
        os.environ["WC_SPECTRUM_VARIANT"] = var
        y = be.execute_frames(x, frames, 65536, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if frames >= 1024 else 30
        e0.record()
        for _ in range(reps):
            y = be.execute_frames(x, frames, 65536, 4)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[var] = (y.clone(), ms)
    d = float((res["6"][0] - res["3"][0]).abs().max())
    print(f"frames {frames}: two-stream {res['5'][1]:.3f} ms ({frames*65536/res['5'][1]/1e6:.1f} GS/s)  single-stream {res['3'][1]:.3f} ms ({frames*65536/res['3'][1]/1e6:.1f})  "
          f"group kernel {res['6'][1]:.3f} ms ({frames*65536/res['6'][1]/1e6:.1f} GS/s)  max |diff| {d:.2e} dB", flush=True)
