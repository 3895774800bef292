#!/usr/bin/env python
"""A few bare wc_analog_run calls of C1 or C2 (eager, no graph) for the ncu launch list: python tools/dev_plan_only.py c1|c2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import wavecap_sdr_b200._native as N
N.init(0)
import bench_configs as BC
from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    fs, n, B, fmt = 2_400_000, 120_000, 64, "cf32"
    x = torch.view_as_complex(torch.randn((B * n, 2), device="cuda") * 0.3).reshape(B, n)
    cfgs = [apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=200000.0))]
else:
    fs, n, B, fmt = 10_000_000, 500_000, 8, "cs16"
    x = torch.randint(-2000, 2000, (B, n, 2), device="cuda", dtype=torch.int16)
    cfgs = []
    for i in range(16):
        c = apply_mode_defaults("nbfm", ChannelConfig(id=str(i), capture_id="c", mode="nbfm", offset_hz=-3.75e6 + 5e5 * i))
        c.squelch_db = -45.0
        cfgs.append(c)
plan = BC._plan_for(cfgs, fs, n, fmt)
N.check(N.lib().wc_analog_plan_use_graph(plan._h, 0))
for _ in range(4):
    plan.run(x, B)
torch.cuda.synchronize()
print("ok", which)
