"""dev: SAM stateless branch vs the golden, with intermediate checks and a throughput probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import analog as oa
from oracle.make_golden import sam_stateless_input, sam_stateless_cases
import wavecap_sdr_b200.capture as cap
from wavecap_sdr_b200.dsp import sam as gs, am as AM, _stages as S

g = np.load("tests/golden/sam.npz")
xs = sam_stateless_input()
for name, kw in sam_stateless_cases():
    cfg = cap.ChannelConfig(id="s", capture_id="c", mode="sam", offset_hz=30000.0)
    for k, v in kw.items():
        setattr(cfg, k, v)
    sig = cap._chain_signature(cfg, 240000)
    print(name, sig)
    a, m = cap._process_channel_dsp_stateless(xs, 240000, cfg)
    print("  ->", None if a is None else (a.shape, float(np.abs(a).max())), m, g[f"st_{name}_m"], float(np.abs(g[f"st_{name}"]).max()))
    base = cap.freq_shift(xs, 30000.0, 240000)
    ob = oa.freq_shift(xs, 30000.0, 240000)
    print("  base err", np.abs(base - ob).max())
    al, be = gs.pll_coefficients(240000.0, sig[6], 0.707)
    rows, st, _ = gs.pll_rows(torch.from_numpy(base).cuda().reshape(1, -1), al, be, sig[5])
    pll = oa.CarrierRecoveryPLLOracle(240000.0, sig[6])
    ci, cq, f = pll.process(ob)
    ref = ci + cq if sig[5] == 1 else ci - cq if sig[5] == 2 else ci
    print("  pll rows err", np.abs(rows.cpu().numpy()[0] - ref).max(), "finite", bool(torch.isfinite(rows).all()))
    out, p, inv = AM.am_tail(rows, 240000, sig[4], sig[1], sig[2], sig[3], want_stats=True)
    from conftest import rel_rms
    print("  final vs golden", rel_rms(a, g[f"st_{name}"]))
    r = rows.cpu().numpy()[0]
    y = r
    for bb, aa in sig[1]:
        from scipy import signal
        y64 = signal.lfilter(np.asarray(bb), np.asarray(aa), y)
        yg = S.lfilter(bb, aa, torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda().reshape(1, -1)).cpu().numpy()[0]
        print("   stage rel", rel_rms(yg, y64.astype(np.float32)), "rms out", float(np.sqrt(np.mean(y64 ** 2))), "rms in", float(np.sqrt(np.mean(np.asarray(y, dtype=np.float64) ** 2))))
        y = y64.astype(np.float32)
    print("  tail", out.shape, float(out.abs().max()), p, inv)

# throughput: 64 sequences x 120000 samples
x = (torch.randn(64, 120000, dtype=torch.complex64, device="cuda") * 0.01 + 0.3)
for n_seq in (1, 64):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gs.pll_rows(x[:n_seq], 1e-3, 1e-6, 0)
    torch.cuda.synchronize(); print(n_seq, "seq x 120000:", (time.perf_counter() - t0) * 1e3, "ms")
