#!/usr/bin/env python
"""device-resident audio mode timing: python tools/dev_audio.py [chunks]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer, OUT_AUDIO, OUT_FM, IN_CF32, fm_scale
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
CH = 6_250_000
ch = PolyphaseChannelizer(125_000_000, 488281)
lib = N.lib()
N.check(lib.wc_chan_audio_config(ch._h, 976560, 48828))
na = int(lib.wc_chan_audio_len(ch._h, CH)); F = ch.frames_for(CH)
x = torch.view_as_complex(torch.randn((nb * CH, 2), device="cuda") * 0.5)
out = torch.empty((nb * na, 256), dtype=torch.float32, device="cuda")
fm = torch.empty((nb * F, 256), dtype=torch.float32, device="cuda")
st = N.torch_stream_ptr()
def t(fn, it=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
a = t(lambda: N.check(lib.wc_chan_process_ex(ch._h, C.c_void_p(x.data_ptr()), IN_CF32, CH, nb, CH, OUT_AUDIO, 0.0, C.c_void_p(out.data_ptr()), st)))
f = t(lambda: N.check(lib.wc_chan_process_ex(ch._h, C.c_void_p(x.data_ptr()), IN_CF32, CH, nb, CH, OUT_FM, fm_scale(976560), C.c_void_p(fm.data_ptr()), st)))
print(f"audio mode {a:.3f} ms ({nb*CH/a/1e6:.1f} GS/s) | fm mode {f:.3f} ms | decimator+finish ~{a-f:.3f} ms")
