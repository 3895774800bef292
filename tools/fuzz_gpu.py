"""Randomised GPU-vs-oracle sweep (dev tool; the fixed cases live in tests/): ragged call sequences, awkward sizes and random
channel configurations, every result compared with the oracle restatement.

    python tools/fuzz_gpu.py [--seed S] [--rounds R] [--only chan,fm,analog,spectrum,c4fm,cqpsk,fir,ddc,disc]

Prints one line per family and exits non-zero on the first mismatch (with the parameters that reproduce it)."""
from __future__ import annotations

import argparse
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from conftest import rel_rms, wrap_rel_rms  # noqa: E402

TOL = 1e-4


def iq(rng, n, scale=0.5):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def odd_len(rng, m, hop, big):
    """lengths around the frame boundaries, tiny ones, and ordinary ones"""
    kind = rng.integers(0, 6)
    if kind == 0:
        return int(rng.integers(0, m + 2))
    if kind == 1:
        return int(m + hop * rng.integers(0, 12) + rng.integers(-1, 2))
    if kind == 2:
        return int(m + hop * rng.integers(0, 3))
    return int(rng.integers(m, big))


def fuzz_chan(rng, rounds):
    from oracle.channelizer import ChannelizerOracle
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer

    worst, calls = 0.0, 0
    geoms = [(125_000_000, 488281, 9), (8_000_000, 25000, 9), (2_400_000, 200000, 5), (1_000_000, 12500, 9), (6_000_000, 12500, 3)]
    for r in range(rounds):
        fs, bw, t = geoms[r % len(geoms)]
        ch, o = PolyphaseChannelizer(fs, bw, t), ChannelizerOracle(fs, bw, t)
        m = o.channel_count
        for c in range(int(rng.integers(2, 7))):
            n = odd_len(rng, m, m // 2, 40 * m)
            x = iq(rng, n)
            got, exp = ch.process_array(x), o.process_vectorized(x)
            assert got.shape == exp.shape, ("chan shape", fs, bw, t, r, c, n, got.shape, exp.shape)
            if exp.size:
                e = rel_rms(got, exp)
                assert e < TOL, ("chan", fs, bw, t, r, c, n, e)
                worst = max(worst, e)
            assert np.array_equal(ch.arm_history, o.arm_history), ("chan history", fs, bw, t, r, c, n)
            calls += 1
            if rng.integers(0, 9) == 0:
                ch.reset(); o.reset()
    return f"channelizer: {calls} calls over {rounds} objects (5 geometries, ragged lengths, carried history), worst rel-RMS {worst:.1e}"


def fuzz_fm(rng, rounds):
    import torch
    from oracle.channelizer import ChannelizerOracle, channelize_fm
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer

    worst, calls = 0.0, 0
    for r in range(rounds):
        ch, o = PolyphaseChannelizer(125_000_000, 488281), ChannelizerOracle(125_000_000, 488281)
        rate = int(ch.channel_sample_rate)
        period = 2 * np.pi * float(np.float32(rate / (2.0 * np.pi * 75000.0)))
        cs16 = bool(rng.integers(0, 2))
        for c in range(int(rng.integers(1, 5))):
            n = odd_len(rng, 256, 128, 90000)
            if cs16:
                q = rng.integers(-20000, 20000, size=(n, 2), dtype=np.int16)
                x = (q[:, 0].astype(np.float32) / np.float32(32768.0) + 1j * (q[:, 1].astype(np.float32) / np.float32(32768.0))).astype(np.complex64)
                arg = torch.from_numpy(q).cuda() if rng.integers(0, 2) else q
            else:
                x = iq(rng, n)
                arg = torch.from_numpy(x).cuda() if rng.integers(0, 2) else x
            got = ch.process_fm(arg, rate)
            got = got.cpu().numpy() if hasattr(got, "is_cuda") else got
            exp = channelize_fm(o.process_vectorized(x), rate)
            assert got.shape == exp.shape, ("fm shape", r, c, n, cs16, got.shape, exp.shape)
            if exp.size:
                e = wrap_rel_rms(got, exp, period)
                assert e < TOL, ("fm", r, c, n, cs16, e)
                worst = max(worst, e)
            calls += 1
    return f"fused FM: {calls} calls over {rounds} objects (cf32 / cs16, host / device input, ragged lengths), worst rel-RMS {worst:.1e}"


def fuzz_analog(rng, rounds):
    from oracle import analog as oa
    from wavecap_sdr_b200.capture import ChannelConfig, process_channels_batch

    worst, pairs, floors = 0.0, 0, []
    for r in range(rounds):
        hi = bool(rng.integers(0, 2))
        fs = int(rng.choice([240_000, 1_000_000, 2_400_000])) if hi else int(rng.choice([48_000, 96_000]))
        n = int(rng.integers(2000, 30000)) if hi else int(rng.integers(1500, 9000))
        n_chunks = int(rng.integers(1, 4))
        modes = ["wbfm", "nbfm", "raw", "p25"] if hi else ["nbfm", "am", "ssb", "sam", "raw"]
        n_ch = int(rng.integers(1, 6))
        t = np.arange(n * n_chunks) / float(fs)
        x = 0.01 * iq(rng, n * n_chunks)
        kws, offs = [], []
        for c in range(n_ch):
            mode = str(rng.choice(modes))
            off = float(int(rng.integers(-fs // 3, fs // 3)))
            kw = dict(mode=mode, audio_rate=int(rng.choice([48000, 16000, 24000, fs])) if fs <= 96_000 else int(rng.choice([48000, 16000])))
            if mode in ("wbfm", "nbfm"):
                kw.update(enable_deemphasis=bool(rng.integers(0, 2)), enable_mpx_filter=bool(rng.integers(0, 2)))
                if not hi:
                    kw.update(enable_fm_highpass=bool(rng.integers(0, 2)), enable_fm_lowpass=bool(rng.integers(0, 2)))
                dev = 75000.0 if mode == "wbfm" else 5000.0
                sig = 0.3 * np.exp(1j * (dev / 1000.0) * np.sin(2 * np.pi * (400.0 + 150 * c) * t))
            else:
                kw.update(enable_agc=bool(rng.integers(0, 2)))
                if mode == "ssb":
                    kw.update(ssb_mode=str(rng.choice(["usb", "lsb"])))
                if mode == "sam":
                    kw.update(sam_sideband=str(rng.choice(["dsb", "usb", "lsb"])), sam_pll_bandwidth_hz=float(rng.choice([30.0, 50.0, 100.0])))
                if rng.integers(0, 3) == 0 and mode in ("am", "ssb"):
                    kw.update(notch_frequencies=[1000.0])
                sig = 0.3 * (1 + 0.5 * np.sin(2 * np.pi * (500.0 + 130 * c) * t)) * np.exp(1j * 0.7)
            x = x + (sig * np.exp(-2j * np.pi * off * t) / max(n_ch, 1)).astype(np.complex64)
            kws.append(kw)
            offs.append(off)
        x = x.astype(np.complex64)
        cfgs = []
        for off, kw in zip(offs, kws):
            cfg = ChannelConfig(id="f", capture_id="c", mode=kw["mode"], offset_hz=off)
            for k, v in kw.items():
                setattr(cfg, k, v)
            cfgs.append(cfg)
        res = process_channels_batch(x, fs, cfgs, n_chunks=n_chunks, use_plan=bool(rng.integers(0, 2)))
        for b in range(n_chunks):
            for ci, (off, kw) in enumerate(zip(offs, kws)):
                ea, em = oa.process_channel_dsp_stateless(x[b * n:(b + 1) * n], fs, oa.OracleChannelConfig(offset_hz=off, **kw))
                a, m = res[b][ci]
                ctx = ("analog", r, fs, n, n_chunks, b, ci, off, kw)
                assert (a is None) == (ea is None), ctx + ("audio presence", a is None, ea is None)
                assert abs(m["rssi_db"] - em["rssi_db"]) < 2e-3, ctx + (m, em)
                if ea is not None:
                    assert a.shape == ea.shape, ctx + (a.shape, ea.shape)
                    e = rel_rms(a, ea)
                    tol = TOL
                    if kw["mode"] in ("am", "ssb", "sam"):
                        # The reference's tf-form Butterworth stages (order-5 100 Hz high-pass, order-10 SSB band-pass) amplify
                        # their own float64 rounding as the sample rate grows: measure how far the ORACLE's
                        # output moves when eight input samples change by one float32 ulp and do not ask for more than that.
                        xp = x[b * n:(b + 1) * n].copy()
                        for k in range(5, 13):          # a handful of early samples, so that at least one survives the
                            #                             float32 roundings of the shift / BFO ahead of the filters
                            xp[k] = np.complex64(complex(np.nextafter(xp[k].real, np.float32(4.0)), np.nextafter(xp[k].imag, np.float32(4.0))))
                        fa, _ = oa.process_channel_dsp_stateless(xp, fs, oa.OracleChannelConfig(offset_hz=off, **kw))
                        floor = rel_rms(fa, ea) if fa is not None and fa.shape == ea.shape else 0.0
                        floors.append(floor)
                        # several carriers in one band through the SAM loop: tests/test_sam_gpu.py (detector singular at the origin)
                        tol = max(5e-4 if kw["mode"] == "sam" else TOL, 4.0 * floor)
                    assert e < tol, ctx + (e, tol)
                    if tol == TOL:
                        worst = max(worst, e)
                    assert abs(m["signal_power_db"] - em["signal_power_db"]) < 5e-3, ctx + (m, em)
                pairs += 1
    return f"analog chain: {pairs} (chunk, channel) pairs over {rounds} random captures (wbfm/nbfm/am/ssb/sam/raw/p25, plan and stage paths), worst rel-RMS {worst:.1e} on the pairs held to 1e-4; {sum(f > 2.5e-5 for f in floors)} am/ssb/sam pairs held to 4 x the reference's own 1-ulp floor (largest floor {max(floors, default=0.0):.1e})"


def fuzz_spectrum(rng, rounds):
    from oracle import spectrum as osp
    from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend

    worst, calls = 0.0, 0
    for r in range(rounds):
        n = int(2 ** rng.integers(8, 17))
        be = CudaFFTBackend(n)
        x = iq(rng, n + int(rng.integers(0, 100)), 0.05)
        x[: n] += (0.3 * np.exp(2j * np.pi * float(rng.uniform(-0.4, 0.4)) * np.arange(n))).astype(np.complex64)
        fs = int(rng.choice([2_400_000, 61_440_000]))
        got = be.execute(x, fs)
        exp_db, exp_f, _ = osp.execute(x, fs, n)
        e = rel_rms(got.power_db, exp_db)
        assert e < TOL and np.allclose(got.freqs, exp_f), ("spectrum", r, n, fs, e)
        worst = max(worst, e)
        calls += 1
    return f"spectrum: {calls} frames of 256..65536 points, worst rel-RMS (dB) {worst:.1e}"


def fuzz_c4fm(rng, rounds):
    from oracle.c4fm import C4FMOracle, modulate_c4fm, random_frames
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator

    total = 0
    for r in range(rounds):
        fs = int(rng.choice([48000, 50000]))
        x = modulate_c4fm(random_frames(rng, n_frames=4, payload=150, gap=40), fs, snr_db=float(rng.uniform(18, 30)),
                          cfo_hz=float(rng.uniform(-60, 60)), timing=float(rng.uniform(0, 1)), seed=int(rng.integers(0, 1 << 30)))
        d, o = C4FMDemodulator(sample_rate=fs), C4FMOracle(sample_rate=fs, portable=True)
        s = 0
        while s < len(x):
            n = int(rng.choice([1, 37, 240, 2400, 2500, 5000, 9999]))
            gd, gs = d.demodulate(x[s:s + n])
            ed, es = o.demodulate(x[s:s + n])
            assert np.array_equal(gd, ed), ("c4fm dibits", r, fs, s, n, int((gd != ed).sum()) if len(gd) == len(ed) else (len(gd), len(ed)))
            assert len(gs) == len(es) and (len(es) == 0 or float(np.max(np.abs(gs - es))) < 1e-4), ("c4fm soft", r, fs, s, n)
            total += len(ed)
            s += n
    return f"c4fm: {total} dibits identical over {rounds} signals cut into random chunk lengths (1 .. 9999 samples)"


def fuzz_cqpsk(rng, rounds):
    from oracle.cqpsk import CQPSKOracle, modulate_cqpsk
    from wavecap_sdr_b200.decoders.p25 import CQPSKDemodulator

    total, knife = 0, 0
    bounds = np.array([-np.pi, -np.pi / 2, 0.0, np.pi / 2, np.pi])
    for r in range(rounds):
        fs = int(rng.choice([48000, 50000]))
        x = modulate_cqpsk(rng.integers(0, 4, 2500), fs, 4800, snr_db=float(rng.uniform(18, 30)), cfo_hz=float(rng.uniform(-60, 60)),
                           timing=float(rng.uniform(0, 1)), seed=int(rng.integers(0, 1 << 30)), amp=float(rng.uniform(0.1, 0.8)))
        d, o = CQPSKDemodulator(sample_rate=fs, symbol_rate=4800), CQPSKOracle(sample_rate=fs, symbol_rate=4800, portable=True)
        s = 0
        while s < len(x):
            n = int(rng.choice([64, 240, 2400, 2500, 7200]))
            g, e = d.demodulate(x[s:s + n]), o.demodulate(x[s:s + n])
            assert len(g) == len(e), ("cqpsk count", r, fs, s, n, len(g), len(e))
            bad = np.nonzero(g != e)[0]
            if bad.size:
                ph = np.array(o.phases)
                margin = np.min(np.abs(ph[bad, None] - bounds[None, :]), axis=1)
                assert np.all(margin < 1e-4), ("cqpsk dibits", r, fs, s, n, margin)
                knife += int(bad.size)
            total += len(e)
            s += n
    return f"cqpsk: {total - knife} of {total} dibits identical over {rounds} signals in random chunk lengths, {knife} knife-edge"


def fuzz_fir(rng, rounds):
    """streaming complex FIR / FIR-decimate (dsp/filters.py:558-668): random tap counts, decimation factors and call lengths,
    state carried through zi exactly as the trunking code does"""
    from scipy import signal
    from oracle import ddc
    from wavecap_sdr_b200.dsp.filters import fir_decimate, fir_filter_complex

    worst, calls = 0.0, 0
    for r in range(rounds):
        nt = int(rng.choice([9, 33, 73, 157, 255]))
        taps = signal.firwin(nt, float(rng.uniform(0.02, 0.4)), window=("kaiser", 7.857))
        d = int(rng.choice([1, 2, 5, 30]))
        x = iq(rng, int(rng.integers(nt, 60000)))
        zi_g = zi_o = signal.lfilter_zi(taps, 1.0).astype(np.complex128) * x[0] if rng.integers(0, 2) else None
        s = 0
        while s < len(x):
            n = int(rng.choice([1, 7, nt - 1, nt, nt + 1, 1000, 12077, 30000]))
            part = x[s:s + n]
            if d == 1:
                (g, zi_g), (e, zi_o) = fir_filter_complex(part, taps, zi_g), ddc.fir_filter_complex(part, taps, zi_o)
            else:
                (g, zi_g), (e, zi_o) = fir_decimate(part, taps, d, zi=zi_g), ddc.fir_decimate(part, taps, d, zi=zi_o)
            assert g.shape == e.shape and g.dtype == e.dtype, ("fir shape", r, nt, d, s, n, g.shape, e.shape, g.dtype, e.dtype)
            assert np.array_equal(zi_g, zi_o), ("fir zi", r, nt, d, s, n)
            if e.size:
                err = rel_rms(g, e) if float(np.abs(e).max()) > 0 else float(np.abs(g).max())
                assert err < TOL, ("fir", r, nt, d, s, n, err)
                worst = max(worst, err)
            calls += 1
            s += n
    return f"streaming FIR: {calls} calls over {rounds} filters (9..255 taps, decimation 1..30, call lengths 1..30000), worst rel-RMS {worst:.1e}"


def fuzz_ddc(rng, rounds):
    from oracle import ddc
    from wavecap_sdr_b200.trunking import DDCBank

    worst, calls = 0.0, 0
    for r in range(rounds):
        fs, d1, d2 = [(6_000_000, 30, 4), (2_400_000, 25, 2), (2_400_000, 25, 1), (10_000_000, 50, 4)][r % 4]
        flavor = "control" if rng.integers(0, 2) else "voice"
        K = int(rng.integers(1, 9))
        offs = [float(int(rng.integers(-fs // 2 + 50_000, fs // 2 - 50_000) // 12_500) * 12_500) for _ in range(K)]
        x = ddc.synth_wideband(int(rng.integers(0, 1 << 30)), int(rng.integers(20_000, 200_000)), fs, offs[:3])
        bank = DDCBank(K, fs, d1, d2, flavor=flavor)
        bank.set_offsets(offs)
        cls = ddc.ControlChannelDDC if flavor == "control" else ddc.VoiceDDC
        orc = [cls(fs, d1, d2, o) for o in offs]
        s = 0
        while s < len(x):
            n = int(rng.choice([d1 * d2 * 3, 5000, 30_000, 123_457]))
            got = bank.process(x[s:s + n])
            for k in range(K):
                exp = orc[k].process(x[s:s + n])
                assert got.shape[1] == len(exp), ("ddc count", r, fs, d1, d2, flavor, k, s, n, got.shape, len(exp))
                if len(exp):
                    err = rel_rms(got[k], exp)
                    assert err < TOL, ("ddc", r, fs, d1, d2, flavor, k, s, n, err)
                    worst = max(worst, err)
            calls += 1
            s += n
    return f"trunking DDC bank: {calls} calls over {rounds} banks (1..8 channels, control / voice, 4 rate plans), worst rel-RMS {worst:.1e}"


def fuzz_disc(rng, rounds):
    """voice-channel path: discriminator audio -> DiscriminatorDemodulator (decoders/p25.py:1105-1345) with random gains
    (auto-gain and spread clamps), offsets and call lengths; dibits and loop state against the oracle"""
    from oracle import discriminator as od
    from oracle.c4fm import modulate_c4fm, random_frames
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorBank

    total = 0
    for r in range(rounds):
        C = int(rng.integers(1, 7))
        aus = []
        for c in range(C):
            x = modulate_c4fm(random_frames(rng, n_frames=3, payload=150, gap=40), 48000, snr_db=float(rng.uniform(15, 30)),
                              cfo_hz=float(rng.uniform(-400, 400)), timing=float(rng.uniform(0, 1)), seed=int(rng.integers(0, 1 << 30)))
            au, _ = od.fm_discriminator(x, 0.0)
            aus.append((au * float(rng.choice([1.0, 0.3, 4.0, 0.02]))).astype(np.float32))
        n = min(len(a) for a in aus)
        A = np.array([a[:n] for a in aus])
        bank = DiscriminatorBank(C, 48000)
        orc = [od.DiscriminatorOracle(48000, portable=True) for _ in range(C)]
        s0 = 0
        while s0 < n:
            ln = int(rng.choice([1, 10, 97, 480, 3000, 4801]))
            dib, soft, cnt = bank.demodulate(A[:, s0:s0 + ln])
            for c in range(C):
                ref = orc[c].demodulate(A[c, s0:s0 + ln].copy())
                k = int(cnt[c])
                assert k == len(ref) and np.array_equal(dib[c, :k], ref), ("disc", r, c, s0, ln, k, len(ref))
                total += k
            s0 += ln
        for c in range(C):
            st = bank.state(c)
            assert abs(st["symbol_spread"] - float(orc[c].spread)) <= 1e-6 and abs(st["symbol_clock"] - float(orc[c].clock)) <= 1e-6, ("disc state", r, c)
    return f"discriminator demodulator: {total} dibits identical over {rounds} banks (1..6 channels, call lengths 1..4801), loop state within 1e-6"


FAMILIES = {"chan": fuzz_chan, "fm": fuzz_fm, "analog": fuzz_analog, "spectrum": fuzz_spectrum, "c4fm": fuzz_c4fm, "cqpsk": fuzz_cqpsk, "fir": fuzz_fir, "ddc": fuzz_ddc, "disc": fuzz_disc}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--only", default=",".join(FAMILIES))
    a = ap.parse_args()
    import wavecap_sdr_b200._native as N

    N.init(0)
    failed = 0
    for name in a.only.split(","):
        rng = np.random.default_rng([a.seed, sorted(FAMILIES).index(name)])
        t0 = time.time()
        try:
            print(FAMILIES[name](rng, a.rounds) + f"  ({time.time() - t0:.1f} s)", flush=True)
        except AssertionError as e:
            failed += 1
            print(f"FAIL {name} seed {a.seed}: {e.args[0] if e.args else e}", flush=True)
    sys.exit(1 if failed else 0)
