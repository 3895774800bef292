"""Dev timing (not the bench contract): fused-FM channelizer with one phase skipped (WC_DEV_ABLATE builds)
and/or experimental variants (WC_CHAN_VAR). Results of ablated kernels are wrong by design; timing only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctypes as C
import wavecap_sdr_b200._native as N
from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer, fm_scale

N.init(0)
n = 6_250_000
B = int(os.environ.get("B", "16"))
x = torch.view_as_complex(torch.randn((B * n, 2), device="cuda") * 0.5)
ch = PolyphaseChannelizer(125_000_000, 488281)
F = ch.frames_for(n)
outf = torch.empty((B * F, 256), dtype=torch.float32, device="cuda")
sc = fm_scale(976562)
def run():
    N.check(N.lib().wc_chan_process(ch._h, C.c_void_p(x.data_ptr()), n, B, n, 1, sc, C.c_void_p(outf.data_ptr()), N.torch_stream_ptr()))
ref = None
for key in os.environ.get("KEYS", "WC_CHAN_ABL").split(","):
    for v in os.environ.get("VALS", "0,1,2,3,4").split(","):
        os.environ[key] = v
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gs = B * n / ms / 1e6
        chk = float(outf[:F].double().abs().sum().item())
        print(json.dumps({key: v, "ms": round(ms, 4), "GS/s": round(gs, 2), "frac_hbm": round(gs * 16 / 6550.1, 3), "checksum": chk}), flush=True)
        os.environ.pop(key, None)
