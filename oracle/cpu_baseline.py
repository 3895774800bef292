"""CPU baseline legs for bench.py (test infrastructure; never on the product path).

Times the oracle restatement of the reference's CPU algorithm for BASELINE config 5
(PolyphaseChannelizer.process + quadrature_demod on every extracted channel) on the host cores
of whatever box this runs on. `faithful=True` is the reference's own per-frame loop
(channelizer.py:114-135); `faithful=False` is the vectorised closed form (a faster port)."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _worker(args):
    seed, n, faithful, reps = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle.channelizer import ChannelizerOracle, channelize_fm

    rng = np.random.default_rng(seed)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
    o = ChannelizerOracle(125_000_000, 488281)
    rate = int(o.channel_sample_rate)
    # warm-up (numpy FFT plan caches, page faults)
    (o.process if faithful else o.process_vectorized)(x[: 256 + 128 * 64])
    o.reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        o.reset()
        frames = (o.process if faithful else o.process_vectorized)(x)
        channelize_fm(frames, rate)
    return time.perf_counter() - t0


def channelizer_fm_cpu(n_per_worker: int, workers: int, faithful: bool = True, reps: int = 1) -> dict:
    """All `workers` processes run the same-size job concurrently; throughput = total samples /
    slowest worker's time."""
    ctx = mp.get_context("spawn")
    jobs = [(1000 + i, n_per_worker, faithful, reps) for i in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        times = [_worker(jobs[0])]
    else:
        with ctx.Pool(workers) as pool:
            times = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    total = n_per_worker * workers * reps
    return {"msps": total / max(times) / 1e6, "seconds": max(times), "wall": wall, "samples": total,
            "workers": workers, "faithful": faithful}


# ---- per-config CPU arms (BASELINE.json configs[0..4]) -------------------------------------------------
# One job = one unit of the config's work (a chunk, a (chunk, channel) pair, a frame, a channel-chunk), repeated by
# every worker process until `seconds` have elapsed. kind "reference" runs the UNMODIFIED reference staged under
# oracle/_ref (oracle/build_ref.py); kind "port" runs the oracle restatement (pinned to reference outputs by
# tests/test_oracle_*.py). Samples are counted the way bench.py counts them for the GPU arm of the same config.

def _make_unit(config: str, use_ref: bool, seed: int):
    """-> (callable running one unit, samples counted per unit, description)."""
    from oracle import analog as oa

    if use_ref:
        from oracle import build_ref

        build_ref.load()
    if config == "C5":
        n = 256 + 128 * 4000
        rng = np.random.default_rng(seed)
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
        if use_ref:
            from wavecapsdr.dsp.channelizer import PolyphaseChannelizer
            from wavecapsdr.dsp.fm import quadrature_demod

            ch = PolyphaseChannelizer(125_000_000, channel_bandwidth=488281)
            rate = int(ch.channel_sample_rate)

            def unit():
                res = ch.process(x)
                for k in range(ch.channel_count):
                    quadrature_demod(ch.extract_channel(res, k), rate)
            return unit, n, "PolyphaseChannelizer.process + quadrature_demod(extract_channel(k)) for the 256 channels"
        from oracle.channelizer import ChannelizerOracle, channelize_fm

        o = ChannelizerOracle(125_000_000, 488281)
        rate = int(o.channel_sample_rate)

        def unit():
            channelize_fm(o.process(x), rate)
        return unit, n, "oracle restatement of PolyphaseChannelizer.process (per-frame loop) + quadrature_demod of 256 channels"
    if config in ("C1", "C2"):
        if config == "C1":
            fs, n = 2_400_000, 120_000
            x = oa.synth_c1(seed=1, n=n)
            kw = dict(mode="wbfm", offset_hz=200000.0)
            counted = n                      # one channel: input samples
        else:
            fs, n = 10_000_000, 500_000
            q, offs = oa.synth_c2(seed=2, n=n)
            x = oa.cs16_to_cf32(q)
            kw = dict(mode="nbfm", offset_hz=float(offs[seed % 16]), enable_deemphasis=False, enable_mpx_filter=False)
            counted = n / 16.0               # one of the 16 channels of a chunk: 1/16 of the chunk's input samples
        if use_ref:
            from wavecapsdr.capture import ChannelConfig, _process_channel_dsp_stateless

            cfg = ChannelConfig(id="b", capture_id="c", mode=kw["mode"], offset_hz=kw["offset_hz"])
            for k, v in kw.items():
                setattr(cfg, k, v)
            if config == "C1":               # Capture._apply_mode_defaults("wbfm") (capture.py:3434-3442)
                cfg.enable_deemphasis, cfg.deemphasis_tau_us, cfg.enable_mpx_filter, cfg.mpx_cutoff_hz = True, 75.0, True, 15_000
                cfg.enable_fm_highpass = cfg.enable_fm_lowpass = cfg.enable_agc = False
            else:                            # "nbfm" (capture.py:3444-3452)
                cfg.enable_fm_highpass = cfg.enable_fm_lowpass = cfg.enable_agc = False

            def unit():
                _process_channel_dsp_stateless(x, fs, cfg)
            return unit, counted, "capture._process_channel_dsp_stateless per (chunk, channel)"
        cfg = oa.OracleChannelConfig(**kw)

        def unit():
            oa.process_channel_dsp_stateless(x, fs, cfg)
        return unit, counted, "oracle restatement of capture._process_channel_dsp_stateless per (chunk, channel)"
    if config == "C3":
        from oracle import spectrum as osp

        x = osp.synth_c3(seed=3, n=65536 * 4)
        if use_ref:
            from wavecapsdr.dsp.fft.scipy_backend import ScipyFFTBackend

            be = ScipyFFTBackend(65536)

            def unit():
                acc = None
                for f in range(4):
                    r = be.execute(x[f * 65536:(f + 1) * 65536], 61_440_000)
                    acc = r.power_db.astype(np.float64) if acc is None else acc + r.power_db
                return (acc / 4).astype(np.float32)
            return unit, 4 * 65536, "ScipyFFTBackend(65536).execute x 4 frames + dB mean"

        def unit():
            return osp.averaged(np.stack([osp.execute(x[f * 65536:(f + 1) * 65536], 61_440_000, 65536)[0] for f in range(4)]), 4)
        return unit, 4 * 65536, "oracle restatement of ScipyFFTBackend.execute x 4 frames + dB mean"
    if config in ("C4-c4fm", "C4-cqpsk"):
        n = 72_000
        rng = np.random.default_rng(seed)
        if config == "C4-c4fm":
            from oracle.c4fm import C4FMOracle, modulate_c4fm, random_frames

            x = modulate_c4fm(random_frames(rng, n_frames=36, payload=150, gap=40), 48000, seed=seed)[:n]
            if use_ref:
                from wavecapsdr.dsp.p25.c4fm import C4FMDemodulator

                d = C4FMDemodulator(sample_rate=48000)
                return (lambda: d.demodulate(x)), len(x), "C4FMDemodulator.demodulate, 1.5 s chunk per channel"
            o = C4FMOracle(sample_rate=48000, portable=True)
            return (lambda: o.demodulate(x)), len(x), "oracle restatement of C4FMDemodulator.demodulate, 1.5 s chunk per channel"
        from oracle.cqpsk import CQPSKOracle, modulate_cqpsk

        x = modulate_cqpsk(rng.integers(0, 4, n // 10 + 8), 48000, 4800, seed=seed)[:n]
        if use_ref:
            from wavecapsdr.decoders.p25 import CQPSKDemodulator

            d = CQPSKDemodulator(sample_rate=48000, symbol_rate=4800)
            return (lambda: d.demodulate(x)), len(x), "CQPSKDemodulator.demodulate, 1.5 s chunk per channel"
        o = CQPSKOracle(sample_rate=48000, symbol_rate=4800, portable=True)
        return (lambda: o.demodulate(x)), len(x), "oracle restatement of CQPSKDemodulator.demodulate, 1.5 s chunk per channel"
    raise ValueError(config)


def _config_worker(args):
    config, use_ref, seed, seconds = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import warnings

    warnings.filterwarnings("ignore")
    unit, counted, what = _make_unit(config, use_ref, seed)
    unit()                                   # warm-up: numba JIT, FFT plans, filter design caches
    t0 = time.perf_counter()
    done = 0
    while True:
        unit()
        done += 1
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done * counted, el, what


def config_cpu(config: str, workers: int, seconds: float = 3.0, prefer_reference: bool = True) -> dict:
    """Aggregate CPU rate of `workers` processes each looping over units of `config` for ~`seconds`."""
    use_ref = False
    if prefer_reference:
        try:
            from oracle import build_ref

            use_ref = build_ref.staged()
        except Exception:
            use_ref = False
    ctx = mp.get_context("spawn")
    jobs = [(config, use_ref, 1000 + i, seconds) for i in range(workers)]
    if workers == 1:
        res = [_config_worker(jobs[0])]
    else:
        with ctx.Pool(workers) as pool:
            res = pool.map(_config_worker, jobs)
    rate = sum(s / t for s, t, _ in res)
    return {"value": round(rate / 1e6, 4), "unit": "MS/s", "cores": workers, "kind": "reference" if use_ref else "port",
            "sample": f"{workers} processes x ~{seconds:.0f} s of: {res[0][2]}", "seconds": round(max(t for _, t, _ in res), 2),
            "samples": int(sum(s for s, _, _ in res))}
