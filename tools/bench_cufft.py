#!/usr/bin/env python
"""C3 against the bar SURVEY §2.2 names: cuFFT C2C + separate elementwise passes on the same B200 (what the reference's
CuPy backend would do, dsp/fft/cupy_backend.py:74-120: window multiply, cufft, fftshift, abs, log10 — here through
torch.fft, which calls cuFFT), batched over the same frames, plus the K-frame dB mean. Prints JSON lines:
cuFFT alone (the transform, nothing else), the cuFFT pipeline, and this repo's fused spectrum kernels, with the maximum
dB difference between the two results."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import wavecap_sdr_b200._native as N

N.init(0)
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n, K = 65536, 4
    peak = 6550.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    be = CudaFFTBackend(n)
    w = torch.from_numpy(np.hanning(n).astype(np.float32)).cuda()
    for frames in (368, 1024, 4096):
        x = torch.view_as_complex(torch.randn((frames * n, 2), device="cuda") * 0.2).reshape(frames, n)

        def cufft_only():
            return torch.fft.fft(x, dim=1)

        def cufft_pipeline():
            X = torch.fft.fftshift(torch.fft.fft(x * w, dim=1), dim=1)
            p = 20.0 * torch.log10(torch.abs(X) + 1e-10)
            return p.reshape(frames // K, K, n).mean(dim=1)

        def ours():
            return be.execute_frames(x.reshape(-1), frames, n, K)

        ref = cufft_pipeline()
        got = ours().reshape(frames // K, n)
        diff = float((got - ref).abs().max().item())
        for name, fn, alg in (("cufft_c2c_only", cufft_only, 16.0), ("cufft_pipeline (window, cufft, fftshift, abs, log10, mean)", cufft_pipeline, 9.0),
                              ("wcsdr_b200 fused spectrum", ours, 9.0)):
            ms = timeit(fn)
            gs = frames * n / ms / 1e6
            print(json.dumps({"frames": frames, "what": name, "ms": round(ms, 4), "GSps": round(gs, 1),
                              "alg_bytes_per_sample": alg, "hbm_frac_at_alg_bytes": round(alg * gs / peak, 4),
                              "max_abs_db_diff_vs_cufft_pipeline": round(diff, 6) if "wcsdr" in name else None}), flush=True)
        del x, ref, got
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
