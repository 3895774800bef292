#!/usr/bin/env python
"""Throughput of the non-headline configs (BASELINE.json configs[0..3] + the trunking fan-out) on one B200.

`headline_configs()` is what bench.py folds into its JSON line as `configs`: C1, C2, C3 and C4 (C4FM + CQPSK banks, 64
channels), each with ms per step (CUDA events after warm-up, inputs resident in HBM), the config's natural metric, its
roofline (algorithmic HBM bytes of SURVEY §8d; for the compute-bound C2 also the share of the FP32 peak) and the SM
clocks sampled during its own timed region. The CLI prints the same records plus the §8f rows as JSON lines
(kept under profiles/).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

import wavecap_sdr_b200._native as N  # noqa: E402

PEAK = 6550.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # SMs x lanes x FMA x max SM clock


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


def timed_with_clocks(fn, sampler_cls=None, warm=3, min_seconds=0.25, max_iters=2000):
    """ms per call (CUDA events around a region of >= min_seconds) + the SM clocks NVML reported inside that region."""
    ms0 = timeit(fn, warm=warm, iters=3)
    iters = int(min(max_iters, max(5, min_seconds * 1e3 / max(ms0, 1e-3))))
    sampler = sampler_cls(torch.cuda.current_device()) if sampler_cls else None
    if sampler:
        sampler.start()
    ms = timeit(fn, warm=0, iters=iters)
    clocks = sampler.stop() if sampler else None
    return ms, iters, clocks


def roofline(alg_bytes, ms, kernel, note=None, flops=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": round(gbs, 1), "peak": PEAK, "unit": "GB/s", "frac": round(gbs / PEAK, 4),
         "alg_bytes_per_step": int(alg_bytes), "kernel": kernel, "traffic": None}
    if flops:
        tf = flops / (ms * 1e-3) / 1e12
        r["fp32"] = {"achieved_tflops": round(tf, 2), "peak_tflops": round(FP32_PEAK_TFLOPS, 1),
                     "frac": round(tf / FP32_PEAK_TFLOPS, 4), "flops_per_step": int(flops)}
    if note:
        r["note"] = note
    return r


def emit(**kw):
    if "alg_bytes" in kw and kw.get("ms"):
        kw["alg_gbs"] = round(kw["alg_bytes"] / (kw["ms"] * 1e-3) / 1e9, 1)
        kw["hbm_frac"] = round(kw["alg_gbs"] / PEAK, 4)
    print(json.dumps(kw), flush=True)


def _plan_for(cfgs, fs, n, fmt):
    from wavecap_sdr_b200 import analog_plan as AP
    from wavecap_sdr_b200 import capture as CAP
    from wavecap_sdr_b200.dsp import _stages as S

    sigs = [CAP._chain_signature(c, fs) for c in cfgs]
    chains = [s[:5] if s[0] in ("fm", "am") else (s[0],) for s in sigs]
    modes = [CAP._MODE_CODE.get(c.mode, 0) for c in cfgs]
    return AP.get_plan(fs, n, S.FMT_CS16 if fmt == "cs16" else S.FMT_CF32, modes, [float(c.offset_hz) for c in cfgs],
                       [0.0] * len(cfgs), [c.squelch_db for c in cfgs], chains)


def _analog_cfg(sampler_cls, name, workload, x, xr, fs, n, B, cfgs, fmt, alg_bytes, flops, kernels, note):
    """step = ONE wc_analog_run call for B chunks x all channels (what capture.py:2489-2597 would issue per batch) + the
    device->host read of its metrics; `python_batch_api_ms` = the same batch through capture.process_channels_batch
    (adds the per-(chunk, channel) Python result assembly the reference's callers expect)."""
    from wavecap_sdr_b200.capture import process_channels_batch

    plan = _plan_for(cfgs, fs, n, fmt)

    def step():
        _, metrics = plan.run(xr, B)
        return metrics.cpu()
    side = torch.cuda.Stream()            # a capturable stream: the plan replays its CUDA graph from the third call on
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ms, iters, clocks = timed_with_clocks(step, sampler_cls)
    torch.cuda.current_stream().wait_stream(side)
    api_ms = timeit(lambda: process_channels_batch(x, fs, cfgs, n_chunks=B, in_fmt=fmt, apply_squelch=True, return_device=True), iters=10)
    stage_ms = timeit(lambda: process_channels_batch(x, fs, cfgs, n_chunks=B, in_fmt=fmt, apply_squelch=True, return_device=True,
                                                     use_plan=False), iters=10)
    return {"config": name, "workload": workload, "step": f"{B} chunks x {n} samples" + (f" x {len(cfgs)} channels" if len(cfgs) > 1 else ""),
            "ms": round(ms, 4), "steps": iters, "metric": "input MS/s", "value": round(B * n / ms / 1e3, 1), "unit": "MS/s",
            "channel_msps": round(len(cfgs) * B * n / ms / 1e3, 1), "realtime_x": round(B * n / fs / (ms * 1e-3), 1),
            "api": "wc_analog_run (one call per batch, CUDA-graph replay) + metrics read on the host",
            "python_batch_api_ms": round(api_ms, 4), "stage_by_stage_path_ms": round(stage_ms, 4),
            "roofline": roofline(alg_bytes, ms, kernels, note=note, flops=flops), "clocks": clocks}


def cfg_c1(sampler_cls=None, B=64):
    """C1: one WBFM channel, 2.4 MS/s cf32, B chunks of 120 000 per call (capture._process_channel_dsp_stateless batch)."""
    from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults

    fs, n = 2_400_000, 120_000
    x = torch.view_as_complex(torch.randn((B * n, 2), device="cuda") * 0.3)
    cfg = apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=200000.0))
    return _analog_cfg(sampler_cls, "C1", "WBFM 1 channel from 2.4 MS/s cf32 (freq_shift, RSSI, discriminator, de-emphasis, 15 kHz MPX "
                       "low-pass, RMS, 1/50 resampler, soft clip)", x, x.reshape(B, n), fs, n, B, [cfg], "cf32",
                       B * n * 8 + B * 2400 * 4, 60 * B * n, "iir_kernel (block scan) + front_kernel + resample_residue_kernel",
                       "8.08 B/sample algorithmic; the chain is latency/FP64-bound (SURVEY §8d), the HBM fraction is stated for completeness")


def cfg_c2(sampler_cls=None, B=8):
    """C2: 16 NBFM channels + squelch from one 10 MS/s int16 capture, B chunks of 500 000."""
    from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults

    fs, n = 10_000_000, 500_000
    q = torch.randint(-2000, 2000, (B, n, 2), device="cuda", dtype=torch.int16)
    cfgs = []
    for i in range(16):
        c = apply_mode_defaults("nbfm", ChannelConfig(id=str(i), capture_id="c", mode="nbfm", offset_hz=-3.75e6 + 5e5 * i))
        c.squelch_db = -45.0
        cfgs.append(c)
    return _analog_cfg(sampler_cls, "C2", "16 NBFM channels + squelch from one 10 MS/s cs16 capture (per channel: NCO, discriminator, RMS, "
                       "3/625 resampler, soft clip)", q, q, fs, n, B, cfgs, "cs16", B * n * 4 + 16 * B * 2400 * 4, 1900 * B * n,
                       "front_kernel + resample_residue_kernel",
                       "4.31 B/sample algorithmic; FP32/SFU compute-bound by construction (16 sincos + 16 atan2 per input sample, "
                       "SURVEY §8d): the fp32 block is the relevant fraction")


def cfg_c3(sampler_cls=None, frames=4096):
    """C3: 65536-point Hann spectrum, dB, fftshift, K=4 dB mean on contiguous frames of a 61.44 MS/s stream."""
    from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend

    be = CudaFFTBackend(65536)
    x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
    be.execute_frames(x, frames, 65536, 4)
    ms, iters, clocks = timed_with_clocks(lambda: be.execute_frames(x, frames, 65536, 4), sampler_cls)
    return {"config": "C3", "workload": "65536-pt windowed spectrum (Hann, |X| dB, fftshift) with K=4 dB averaging, 61.44 MS/s cf32",
            "step": f"{frames} contiguous frames", "ms": round(ms, 4), "steps": iters, "metric": "input MS/s",
            "value": round(frames * 65536 / ms / 1e3, 1), "unit": "MS/s", "frames_per_s": round(frames / (ms * 1e-3)),
            "realtime_x": round(frames * 65536 / 61.44e6 / (ms * 1e-3), 1),
            "roofline": roofline(frames * 65536 * 8 + frames // 4 * 65536 * 4, ms, "spectrum passes (csrc/spectrum.cu)",
                                 note="9 B/sample algorithmic (8 in + 4/K out)"),
            "clocks": clocks}


def cfg_c4(sampler_cls=None, C=64, n=72000):
    """C4: P25 C4FM and CQPSK banks, C channels at 48 kS/s, one demodulate() of n samples per channel."""
    from oracle.c4fm import modulate_c4fm, random_frames
    from oracle.cqpsk import modulate_cqpsk
    from wavecap_sdr_b200.decoders.p25 import CQPSKBank
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMBank

    fs = 48000
    rng = np.random.default_rng(1)
    base = modulate_c4fm(random_frames(rng, n_frames=(n // 2140) + 2, payload=150, gap=40), fs, seed=1)[:n]
    x = torch.from_numpy(np.ascontiguousarray(np.tile(base, (C, 1)))).cuda()
    x = x * torch.exp(1j * torch.rand((C, 1), device="cuda") * 6.28).to(torch.complex64)
    bank = C4FMBank(C, fs)
    out = []
    ms, iters, clocks = timed_with_clocks(lambda: bank.demodulate(x), sampler_cls, warm=2)
    out.append({"config": "C4-c4fm", "workload": f"P25 C4FM symbol recovery bank, {C} channels at 48 kS/s", "step": f"one demodulate() of {n} samples/channel "
                f"({n / fs * 1e3:.0f} ms of signal)", "ms": round(ms, 4), "steps": iters, "metric": "channel MS/s", "value": round(C * n / ms / 1e3, 2),
                "unit": "MS/s", "channels_x_realtime": round(C * n / fs / (ms * 1e-3)),
                "roofline": roofline(C * n * 8 + C * (n // 10) * 5, ms, "p25_fir_kernel + c4fm_phase_kernel + c4fm_sync_kernel",
                                     note="8.5 B per channel-sample algorithmic; sequential per channel, latency-bound by construction (SURVEY §8d)"),
                "clocks": clocks})
    cb = modulate_cqpsk(rng.integers(0, 4, n // 10 + 8), fs, 4800, seed=2)[:n]
    xq = torch.from_numpy(np.ascontiguousarray(np.tile(cb, (C, 1)))).cuda()
    qb = CQPSKBank(C, fs)
    ms, iters, clocks = timed_with_clocks(lambda: qb.demodulate(xq), sampler_cls, warm=2)
    out.append({"config": "C4-cqpsk", "workload": f"P25 CQPSK symbol recovery bank, {C} channels at 48 kS/s", "step": f"one demodulate() of {n} samples/channel",
                "ms": round(ms, 4), "steps": iters, "metric": "channel MS/s", "value": round(C * n / ms / 1e3, 2), "unit": "MS/s",
                "channels_x_realtime": round(C * n / fs / (ms * 1e-3)),
                "roofline": roofline(C * n * 8 + C * (n // 10), ms, "cqpsk_* kernels (csrc/cqpsk.cu)",
                                     note="sequential MMSE/Gardner/frequency loop per channel, latency-bound by construction (SURVEY §8d)"),
                "clocks": clocks})
    return out


def headline_configs(sampler_cls=None):
    """C1..C4 records for bench.py's `configs` array (inputs generated on the device, resident in HBM when timed)."""
    recs = []
    for fn in (cfg_c1, cfg_c2, cfg_c3):
        try:
            recs.append(fn(sampler_cls))
        except Exception as e:  # a config must never take the headline down with it
            recs.append({"config": fn.__name__[4:].upper(), "error": f"{type(e).__name__}: {e}"})
        torch.cuda.empty_cache()
    try:
        recs.extend(cfg_c4(sampler_cls))
    except Exception as e:
        recs.append({"config": "C4", "error": f"{type(e).__name__}: {e}"})
    torch.cuda.empty_cache()
    return recs


def c1_c2():
    for r in (cfg_c1(), cfg_c2()):
        print(json.dumps(r), flush=True)


def c3():
    for frames in (46 * 8, 4096):
        print(json.dumps(cfg_c3(frames=frames)), flush=True)


def c4():
    for C in (64, 1024):
        for n in (2400, 72000):
            for r in cfg_c4(C=C, n=n):
                print(json.dumps(r), flush=True)


def framer():
    """NID BCH decode + message framer, batched over channels (device-resident symbols)."""
    from oracle import p25_framer as of
    from wavecap_sdr_b200.decoders.p25_framer import P25FramerBank
    from wavecap_sdr_b200.dsp.fec.bch import bch_decode_batch
    import ctypes as C

    rng = np.random.default_rng(3)
    s = list(rng.integers(0, 4, 50))
    while len(s) < 7300:
        s += of.frame_dibits(rng, 0x293, 0x7, 588, nid_errors=int(rng.integers(0, 6)))
    dib = np.array(s[:7200], dtype=np.uint8)
    soft = (of.dibits_to_soft(dib) + rng.normal(0, 0.3, len(dib))).astype(np.float32)
    for Cn in (64, 1024):
        D = torch.from_numpy(np.tile(dib, (Cn, 1))).cuda()
        S = torch.from_numpy(np.tile(soft, (Cn, 1))).cuda()
        for mode in (0, 1):
            bank = P25FramerBank(Cn)
            l = N.lib()
            mm, pc = l.wc_p25framer_max_msgs(7200), l.wc_p25framer_pool_bytes(7200)
            hdr = torch.zeros((Cn, mm, 6), dtype=torch.int32, device="cuda")
            hs = torch.zeros((Cn, mm), dtype=torch.int64, device="cuda")
            pool = torch.zeros((Cn, pc), dtype=torch.uint8, device="cuda")
            summ = torch.zeros((Cn, 8), dtype=torch.int32, device="cuda")

            def run():
                N.check(l.wc_p25framer_process(bank._h, C.c_void_p(S.data_ptr()), C.c_void_p(D.data_ptr()), 7200, None, 7200,
                                               mode, 1, None, C.c_void_p(hdr.data_ptr()), C.c_void_p(hs.data_ptr()),
                                               C.c_void_p(pool.data_ptr()), C.c_void_p(summ.data_ptr()), N.torch_stream_ptr()))
            ms = timeit(run, warm=2, iters=5)
            sm = summ.cpu().numpy()
            emit(config=f"P25 framer bank, {Cn} ch ({'process_batch' if mode == 0 else 'process_with_soft_sync'} order)",
                 step="one call of 7200 symbols/channel (1.5 s of signal): sync scores + NID BCH + message assembly",
                 ms=round(ms, 3), channel_ksym_per_s=round(Cn * 7200 / ms, 1), channels_x_realtime=round(Cn * 1.5 / (ms * 1e-3)),
                 msgs_per_channel=int(sm[0, 0]), nids_per_channel=int(sm[0, 1]), alg_bytes=Cn * 7200 * 5)
    cw = rng.integers(0, 2, (1 << 16, 63)).astype(np.uint8)
    bits = torch.from_numpy(cw).cuda()
    data = torch.zeros(1 << 16, dtype=torch.int32, device="cuda")
    errs = torch.zeros(1 << 16, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: N.check(N.lib().wc_bch_decode(C.c_void_p(bits.data_ptr()), None, 1 << 16, C.c_void_p(data.data_ptr()),
                                                      C.c_void_p(errs.data_ptr()), N.torch_stream_ptr())))
    emit(config="BCH(63,16,23) batch decode, 65536 random words (worst case: all run BM + Chien)", step="one launch",
         ms=round(ms, 3), mwords_per_s=round(65536 / ms / 1e3, 2))


def voice():
    """voice-channel FM discriminator + DiscriminatorDemodulator bank on device-resident IQ."""
    from oracle.c4fm import modulate_c4fm, random_frames
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorBank
    from wavecap_sdr_b200.trunking import VoiceDiscriminator

    fs, n = 48000, 72000
    rng = np.random.default_rng(2)
    base = modulate_c4fm(random_frames(rng, n_frames=(n // 2140) + 2, payload=150, gap=40), fs, seed=2)[:n]
    for Cn in (64, 1024):
        x = torch.from_numpy(np.ascontiguousarray(np.tile(base, (Cn, 1)))).cuda()
        disc, bank = VoiceDiscriminator(Cn), DiscriminatorBank(Cn, fs)
        au = disc.process(x).to(torch.float32)
        ms_a = timeit(lambda: disc.process(x), warm=2, iters=5)
        ms_b = timeit(lambda: bank.demodulate(au), warm=2, iters=5)
        emit(config=f"voice path, {Cn} ch, 48 kS/s: FM discriminator + DiscriminatorDemodulator", step=f"one call of {n} samples/channel (1.5 s)",
             ms=round(ms_a + ms_b, 3), ms_discriminator=round(ms_a, 3), ms_demodulator=round(ms_b, 3),
             channel_msps=round(Cn * n / (ms_a + ms_b) / 1e3, 2), channels_x_realtime=round(Cn * 1.5 / ((ms_a + ms_b) * 1e-3)),
             alg_bytes=Cn * n * 8 + Cn * (n // 10))


def scanner():
    """control-channel scan of one 100 ms wideband block (device resident)."""
    from wavecap_sdr_b200.cc_scanner import ControlChannelScanner

    fs, n = 6_000_000, 600_000
    x = torch.view_as_complex(torch.randn((n, 2), device="cuda") * 0.1)
    for K in (8, 24, 96):
        sc = ControlChannelScanner(center_hz=851e6, sample_rate=fs, control_channels=list(851e6 + np.linspace(-2.8e6, 2.8e6, K)))
        ms = timeit(lambda: sc.scan_all(x), warm=2, iters=5)
        emit(config=f"control-channel scan, {K} candidates + 2 noise probes from 6 MS/s", step=f"one {n}-sample block (100 ms), incl. host result assembly",
             ms=round(ms, 3), candidates_per_s=round(K / (ms * 1e-3)), realtime_x=round(0.1 / (ms * 1e-3), 1), alg_bytes=n * 8)


def ddc():
    from wavecap_sdr_b200.trunking import DDCBank

    fs, n = 6_000_000, 300_000
    x = torch.view_as_complex(torch.randn((n, 2), device="cuda") * 0.1)
    for K in (1, 24, 96):
        b = DDCBank(K, fs, 30, 4)
        b.set_offsets(np.linspace(-2.9e6, 2.9e6, K))
        ms = timeit(lambda: b.process(x))
        emit(config=f"trunking fan-out, {K} ch from 6 MS/s (NCO + 157-tap /30 + 73-tap /4)", step=f"one {n}-sample chunk (50 ms)",
             ms=round(ms, 3), input_msps=round(n / ms / 1e3, 1), channel_msps=round(K * n / ms / 1e3, 1),
             realtime_x=round(n / fs / (ms * 1e-3), 1), alg_bytes=n * 8 + K * (n // 120) * 8)


if __name__ == "__main__":
    N.init(0)
    which = sys.argv[1:] or ["c1c2", "c3", "c4", "ddc", "framer", "voice", "scanner"]
    t0 = time.time()
    if "c1c2" in which:
        c1_c2()
    if "c3" in which:
        c3()
    if "c4" in which:
        c4()
    if "ddc" in which:
        ddc()
    if "framer" in which:
        framer()
    if "voice" in which:
        voice()
    if "scanner" in which:
        scanner()
    print(json.dumps({"wall_s": round(time.time() - t0, 1), "hbm_peak_gbs": PEAK}), flush=True)
