#!/usr/bin/env python
"""Where a C1 / C2 step goes with the one-call plan: stage path vs plan (eager) vs plan (graph replay), and the bare
wc_analog_run call without result assembly. CUDA events + host wall clock per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavecap_sdr_b200._native as N
N.init(0)
from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults, process_channels_batch
from wavecap_sdr_b200 import analog_plan as AP
from wavecap_sdr_b200 import capture as CAP
from wavecap_sdr_b200.dsp import _stages as S

def timeit(fn, warm=4, iters=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); host = (time.perf_counter() - t0) / iters * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, host

def case(name, x, fs, cfgs, B, fmt):
    kw = dict(n_chunks=B, in_fmt=fmt, apply_squelch=True, return_device=True)
    a = timeit(lambda: process_channels_batch(x, fs, cfgs, use_plan=False, **kw))
    b = timeit(lambda: process_channels_batch(x, fs, cfgs, **kw))
    sigs = [CAP._chain_signature(c, fs) for c in cfgs]
    chains = [s[:5] if s[0] in ("fm", "am") else (s[0],) for s in sigs]
    modes = [CAP._MODE_CODE.get(c.mode, 0) for c in cfgs]
    n = (x.numel() // 2 if fmt == "cs16" else x.numel()) // B
    plan = AP.get_plan(fs, n, S.FMT_CS16 if fmt == "cs16" else S.FMT_CF32, modes, [float(c.offset_hz) for c in cfgs], [0.0] * len(cfgs),
                       [c.squelch_db for c in cfgs], chains)
    xr = x.reshape(B, n, 2) if fmt == "cs16" else x.reshape(B, n)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):        # graphs need a capturable (non-legacy) stream
        c = timeit(lambda: plan.run(xr, B))
        N.check(N.lib().wc_analog_plan_use_graph(plan._h, 0))
        d = timeit(lambda: plan.run(xr, B))
        N.check(N.lib().wc_analog_plan_use_graph(plan._h, 1))
        if B > 1:                        # one chunk per call: where launch overhead matters
            x1 = xr[:1].contiguous()
            g1 = timeit(lambda: plan.run(x1, 1), iters=100)
            N.check(N.lib().wc_analog_plan_use_graph(plan._h, 0))
            e1 = timeit(lambda: plan.run(x1, 1), iters=100)
            N.check(N.lib().wc_analog_plan_use_graph(plan._h, 1))
            print(f"{name} single chunk: graph {g1[0]:.4f} ms (host {g1[1]:.4f}) | eager {e1[0]:.4f} ms (host {e1[1]:.4f})", flush=True)
    print(f"{name}: stage path {a[0]:.3f} ms (host {a[1]:.3f}) | process_channels_batch via plan {b[0]:.3f} (host {b[1]:.3f}) | "
          f"bare wc_analog_run graph {c[0]:.3f} (host {c[1]:.3f}) | bare eager {d[0]:.3f} (host {d[1]:.3f})", flush=True)

fs, n, B = 2_400_000, 120_000, 64
x = torch.view_as_complex(torch.randn((B * n, 2), device="cuda") * 0.3)
cfg = apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=200000.0))
case("C1", x, fs, [cfg], B, "cf32")
fs, n, B = 10_000_000, 500_000, 8
q = torch.randint(-2000, 2000, (B, n, 2), device="cuda", dtype=torch.int16)
cfgs = []
for i in range(16):
    c = apply_mode_defaults("nbfm", ChannelConfig(id=str(i), capture_id="c", mode="nbfm", offset_hz=-3.75e6 + 5e5 * i))
    c.squelch_db = -45.0
    cfgs.append(c)
case("C2", q, fs, cfgs, B, "cs16")
