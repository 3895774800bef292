"""Dev timing of the channelizer kernel (not the bench contract): sweeps occupancy (WC_CHAN_OCC), run length R and batch."""
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
outc = torch.empty((B * F, 256), dtype=torch.complex64, device="cuda")
outf = torch.empty((B * F, 256), dtype=torch.float32, device="cuda")
sc = fm_scale(976562)
def run(mode, out, nb):
    N.check(N.lib().wc_chan_process(ch._h, C.c_void_p(x.data_ptr()), n, nb, n, mode, sc, C.c_void_p(out.data_ptr()), N.torch_stream_ptr()))
modes = [int(m) for m in os.environ.get("MODES", "0,1").split(",")]
for occ in [int(v) for v in os.environ.get("OCCS", "5").split(",")]:
    os.environ["WC_CHAN_OCC"] = str(occ)
    for R in [int(r) for r in os.environ.get("RS", "0").split(",")]:
        if R: os.environ["WC_CHAN_R"] = str(R)
        else: os.environ.pop("WC_CHAN_R", None)
        for mode, out, bps in ((0, outc, 24), (1, outf, 16)):
            if mode not in modes: continue
            for nb in (1, B):
                for _ in range(3): run(mode, out, nb)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                it = 20 if nb == 1 else 8
                e0.record()
                for _ in range(it): run(mode, out, nb)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / it
                gs = nb * n / ms / 1e6
                print(json.dumps({"occ": occ, "R": R or "auto", "mode": mode, "chunks": nb, "ms": round(ms, 4), "GS/s": round(gs, 2), "GB/s": round(gs * bps, 1), "frac_hbm": round(gs * bps / 6550.1, 3)}), flush=True)
