#!/usr/bin/env python
"""Throughput of the non-headline configs (BASELINE.json configs[0..3] + the trunking fan-out) on one B200.
Not the bench.py contract — a measurement aid whose JSON lines are kept under profiles/.

Each line: config, what one "step" is, ms per step (CUDA events, after warm-up, inputs resident in HBM), the
metric in the unit natural to the config, algorithmic HBM bytes per step and the HBM fraction they imply
(SURVEY §8d says which configs are HBM-bound and which are compute/latency-bound by construction).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import wavecap_sdr_b200._native as N  # noqa: E402

N.init(0)
PEAK = 6550.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


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


def emit(**kw):
    if "alg_bytes" in kw and kw.get("ms"):
        kw["alg_gbs"] = round(kw["alg_bytes"] / (kw["ms"] * 1e-3) / 1e9, 1)
        kw["hbm_frac"] = round(kw["alg_gbs"] / PEAK, 4)
    print(json.dumps(kw), flush=True)


def c1_c2():
    from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults, process_channels_batch

    # C1: one WBFM channel, 2.4 MS/s cf32, B chunks of 120 000 per call
    fs, n, B = 2_400_000, 120_000, 64
    x = torch.view_as_complex(torch.randn((B * n, 2), device="cuda") * 0.3)
    cfg = apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=200000.0))
    ms = timeit(lambda: process_channels_batch(x, fs, [cfg], n_chunks=B, return_device=True), iters=5)
    emit(config="C1 WBFM 1 ch, 2.4 MS/s cf32", step=f"{B} chunks x {n} samples, full wbfm chain incl. host result assembly",
         ms=round(ms, 3), msps=round(B * n / ms / 1e3, 1), realtime_x=round(B * n / fs / (ms * 1e-3), 1),
         alg_bytes=B * n * 8 + B * 2400 * 4)
    # C2: 16 NBFM channels, 10 MS/s int16, B chunks of 500 000
    fs, n, B = 10_000_000, 500_000, 8
    q = torch.randint(-2000, 2000, (B, n, 2), device="cuda", dtype=torch.int16)
    cfgs = []
    for i in range(16):
        c = apply_mode_defaults("nbfm", ChannelConfig(id=str(i), capture_id="c", mode="nbfm", offset_hz=-3.75e6 + 5e5 * i))
        c.squelch_db = -45.0
        cfgs.append(c)
    ms = timeit(lambda: process_channels_batch(q, fs, cfgs, n_chunks=B, in_fmt="cs16", apply_squelch=True,
                                               return_device=True), iters=5)
    emit(config="C2 16 NBFM ch + squelch, 10 MS/s cs16", step=f"{B} chunks x {n} samples x 16 channels",
         ms=round(ms, 3), input_msps=round(B * n / ms / 1e3, 1), channel_msps=round(16 * B * n / ms / 1e3, 1),
         realtime_x=round(B * n / fs / (ms * 1e-3), 1), alg_bytes=B * n * 4 + 16 * B * 2400 * 4)


def c3():
    from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend

    be = CudaFFTBackend(65536)
    for frames in (46 * 8, 4096):
        x = torch.view_as_complex(torch.randn((frames * 65536, 2), device="cuda") * 0.2)
        out = be.execute_frames(x, frames, 65536, 4)
        ms = timeit(lambda: be.execute_frames(x, frames, 65536, 4))
        emit(config="C3 65536-pt spectrum, Hann, dB, fftshift, K=4 mean, 61.44 MS/s", step=f"{frames} contiguous frames",
             ms=round(ms, 3), msps=round(frames * 65536 / ms / 1e3, 1), frames_per_s=round(frames / (ms * 1e-3)),
             realtime_x=round(frames * 65536 / 61.44e6 / (ms * 1e-3), 1), alg_bytes=frames * 65536 * 8 + frames // 4 * 65536 * 4)
        del x, out


def c4():
    from wavecap_sdr_b200.decoders.p25 import CQPSKBank
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMBank
    from oracle.c4fm import modulate_c4fm, random_frames
    from oracle.cqpsk import modulate_cqpsk

    fs = 48000
    for C in (64, 1024):
        for n in (2400, 72000):
            rng = np.random.default_rng(1)
            base = modulate_c4fm(random_frames(rng, n_frames=(n // 2140) + 2, payload=150, gap=40), fs, seed=1)[:n]
            x = torch.from_numpy(np.ascontiguousarray(np.tile(base, (C, 1)))).cuda()
            x = x * torch.exp(1j * torch.rand((C, 1), device="cuda") * 6.28).to(torch.complex64)
            bank = C4FMBank(C, fs)
            ms = timeit(lambda: bank.demodulate(x), warm=2, iters=5)
            emit(config=f"C4 C4FM bank, {C} ch, 48 kS/s", step=f"one demodulate() of {n} samples/channel ({n / fs * 1e3:.0f} ms of signal)",
                 ms=round(ms, 3), channel_msps=round(C * n / ms / 1e3, 2), realtime_x_per_channel=round(n / fs / (ms * 1e-3), 1),
                 channels_x_realtime=round(C * n / fs / (ms * 1e-3)), alg_bytes=C * n * 8 + C * (n // 10) * 5)
            cb = modulate_cqpsk(rng.integers(0, 4, n // 10 + 8), fs, 4800, seed=2)[:n]
            xq = torch.from_numpy(np.ascontiguousarray(np.tile(cb, (C, 1)))).cuda()
            qb = CQPSKBank(C, fs)
            ms = timeit(lambda: qb.demodulate(xq), warm=2, iters=5)
            emit(config=f"C4 CQPSK bank, {C} ch, 48 kS/s", step=f"one demodulate() of {n} samples/channel",
                 ms=round(ms, 3), channel_msps=round(C * n / ms / 1e3, 2), realtime_x_per_channel=round(n / fs / (ms * 1e-3), 1),
                 channels_x_realtime=round(C * n / fs / (ms * 1e-3)), alg_bytes=C * n * 8 + C * (n // 10))


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
