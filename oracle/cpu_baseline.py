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
