"""dev: time the resampler alone (CUDA events) for the C2 (3/625) and C1 (1/50) shapes."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.dsp import _stages as S

def run(label, in_rate, out_rate, n, n_seq):
    x = torch.randn((n_seq, n), device="cuda")
    f = lambda: S.resample_dev(x, in_rate, out_rate) if hasattr(S, "resample_dev") else None
    return f

if __name__ == "__main__":
    import ctypes as C
    from math import gcd
    from scipy import signal
    import numpy as np
    for label, in_rate, out_rate, n, n_seq in (("C2 3/625", 10_000_000, 48_000, 500_000, 128), ("C1 1/50", 2_400_000, 48_000, 120_000, 64)):
        g = gcd(in_rate, out_rate); up, down = out_rate // g, in_rate // g
        mx = max(up, down)
        taps = np.ascontiguousarray(signal.firwin(2 * 10 * mx + 1, 1.0 / mx, window=("kaiser", 5.0)) * up, dtype=np.float64)
        h = C.c_void_p()
        N.check(N.lib().wc_resampler_create(up, down, N.np_ptr(taps), taps.size, C.byref(h)))
        x = torch.randn((n_seq, n), device="cuda")
        n_out = int(N.lib().wc_resampler_out_len(h, n))
        out = torch.empty((n_seq, n_out), device="cuda")
        st = N.torch_stream_ptr()
        def call():
            N.check(N.lib().wc_resampler_run(h, C.c_void_p(x.data_ptr()), n, n, n_seq, C.c_void_p(out.data_ptr()), 0, None, 0.0, 0.0, None, None, 1e30, st))
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        ref = signal.resample_poly(x[0].cpu().numpy().astype(np.float64), up, down)
        err = float(np.sqrt(np.mean((out[0].cpu().numpy() - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)))
        print(json.dumps({"case": label, "env": {k: v for k, v in os.environ.items() if k.startswith("WC_RS") or k.startswith("WC_RESAMPLE")}, "us": round(e0.elapsed_time(e1) * 100, 1), "rel_rms_vs_scipy": err}), flush=True)
