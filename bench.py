#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native WaveCap-SDR hot path.

Workload (BASELINE.json configs[4], the config the metric is quoted on): 256-channel polyphase
channelizer + FM discriminator on synthetic 125 MS/s cf32 wideband IQ. One "step" = one launch over
a batch of `--chunks` consecutive 50 ms capture chunks (6.25 M samples each, exactly the reference's
per-call semantics: frames never straddle chunks, filter history carries across them).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0). `value` = aggregate channelized MS/s with inputs resident in HBM;
`e2e` = the same metric through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed
region); `roofline` = algorithmic HBM bytes of the dominant kernel / its measured duration;
`cpu_baseline` = the oracle port of the reference algorithm timed on this box's host cores.
Multi-GPU: one process per GPU, each an independent capture (SURVEY §8e mode i, weak scaling, no
data-path collective); timing is max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 125_000_000
BW = 488281
CHUNK = FS // 20              # capture.py:3035 chunk = max(8192, sample_rate // 20) = 6 250 000
ALG_BYTES_PER_SAMPLE = 16     # 8 B cf32 in + 256 f32 out per 128 in
METRIC = "aggregate channelized MS/s (256-ch polyphase channelizer + FM demod, 125 MS/s cf32)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=128, help="50 ms chunks per step (device-resident leg)")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks per step of the host-buffer leg")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--stripe-chunks", type=int, default=128, help="stripe mode: 50 ms chunks per block (one process() call, cut into N slabs)")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "broadcast", "pull", "stripe"],
                    help="replicas: one independent capture per GPU (headline, weak scaling); broadcast: ONE capture, "
                         "each block NCCL-broadcast from rank 0 and time-sharded over the ranks (north-star-literal, strong); "
                         "pull: ONE capture, rank 0's block mapped into every rank, each rank's channelizer kernel pulls "
                         "only its weighted time slab over NVLink while it computes (no collective on the data path); "
                         "stripe: ONE capture striped at ingest — slab r of every block lands in rank r's own memory, only the "
                         "9-row halo crosses NVLink")
    ap.add_argument("--pull-local-share", type=float, default=0.0,
                    help="pull mode: share of each block rank 0 keeps (0 = local/(local+link) from --pull-rates)")
    ap.add_argument("--pull-tune-rounds", type=int, default=3,
                    help="pull mode: rounds of share re-balancing from measured per-rank kernel times (0 = keep the initial shares)")
    ap.add_argument("--pull-rates", default="207,96", help="pull mode: rank 0's resident rate and the NVLink egress bound, GS/s")
    ap.add_argument("--bcast-chunks", type=int, default=8, help="50 ms chunks per broadcast block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1-C4 `configs` legs (N=1 only)")
    ap.add_argument("--no-one-capture", action="store_true", help="N>1: skip the `one_capture` (stripe / pull) legs")
    ap.add_argument("--no-modes", action="store_true", help="skip the int16 / audio e2e modes (profiling runs)")
    ap.add_argument("--sustained-seconds", type=float, default=2.5,
                    help="extra back-to-back C5 steps after the K timed ones, for the sustained number (0 = off)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="cpu_baseline: seconds of CPU work per core (C5)")
    ap.add_argument("--cpu-samples", type=int, default=3_000_000, help="cpu_baseline samples per worker")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons DURING the timed region via NVML (5 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this rank's host threads (and therefore its pinned staging buffers, first-touch) to the CPUs NVML reports as
    local to its GPU: the end-to-end leg moves 90 GB/s per rank through host memory and loses most of it across sockets."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {local_rank}"
    except Exception as e:  # affinity is an optimisation, never a requirement
        return f"unbound ({type(e).__name__})"
    return "unbound"


def init_nccl(local_rank: int) -> None:
    """NCCL prints its version banner on fd 1 when the communicator is created; the contract is ONE JSON line on
    stdout, so fd 1 points at stderr until the first collective has run."""
    import torch
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


WORKLOAD = ("C5: 256-ch polyphase channelizer (M=256, 9 taps/arm, hop 128) + FM discriminator "
            "of every channel, 125 MS/s cf32, 50 ms chunks (6.25 M samples)")


def cpu_baseline(samples_per_worker: int, workers: int, seconds: float = 12.0) -> dict:
    """C5 on the host cores: the UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py) when it travelled with
    the snapshot (`kind: "reference"`), else the oracle port of its per-frame loop (`kind: "port"`)."""
    from oracle.cpu_baseline import channelizer_fm_cpu, config_cpu

    r = config_cpu("C5", workers, seconds=seconds)
    try:
        v = channelizer_fm_cpu(samples_per_worker * 4, workers, faithful=False)
        r["vectorized_port_msps"] = round(v["msps"], 3)
    except Exception:
        pass
    r["value"] = round(r["value"], 3)
    return r


def config_cpu_baselines(workers: int, seconds: float = 2.5) -> dict:
    """CPU arm of every `configs` entry (C1..C4), same kind selection as cpu_baseline()."""
    from oracle.cpu_baseline import config_cpu

    out = {}
    for name in ("C1", "C2", "C3", "C4-c4fm", "C4-cqpsk"):
        try:
            out[name] = config_cpu(name, workers, seconds=seconds)
        except Exception as e:
            out[name] = {"value": None, "unit": "MS/s", "cores": workers, "kind": "port", "sample": f"failed: {e}"}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the C5 path (oracle/_ref: PolyphaseChannelizer.process +
    quadrature_demod per channel, unmodified; the oracle port only if the staged copy is absent) on all host cores, same
    metric/config. Each step is a bounded sample: every core loops process() over a 512 256-sample slice of the chunk for ~`seconds`."""
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    from oracle.cpu_baseline import config_cpu

    steps = max(1, min(args.steps, 5))
    warm = max(0, min(args.warmup, 1))
    per_step_s = max(2.0, min(args.cpu_seconds, 6.0))
    for _ in range(warm):
        config_cpu("C5", workers, seconds=1.0)
    rs = [config_cpu("C5", workers, seconds=per_step_s) for _ in range(steps)]
    total = sum(r["samples"] for r in rs)
    secs = sum(r["seconds"] for r in rs)
    msps = sum(r["value"] for r in rs) / steps
    kind = rs[0]["kind"]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(msps, 3), "unit": "MS/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(1e3 * secs / steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (numpy)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": f"{workers} processes x ~{per_step_s:.0f} s of process() calls on a 512 256-sample slice of the 50 ms chunk (bounded sample)"},
        "cpu_baseline": {"value": round(msps, 3), "unit": "MS/s", "cores": workers, "kind": kind,
                         "sample": f"{steps} steps; {rs[0]['sample']}; {total} samples in total"},
        "e2e": {"value": round(msps, 3), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_broadcast(args, rank, local_rank, world):
    """ONE capture for the whole box: every step rank 0's next block (bcast-chunks x 6.25 M samples, resident in its
    HBM = the ingest GPU) is NCCL-broadcast to all ranks on a side stream, double-buffered against the compute of
    the previous block; each rank then channelizes + FM-demodulates its time slab of the block (9-frame halo, all 256
    channels). Bound: every rank must RECEIVE the whole block, so the job cannot exceed NVLink ingress / 8 B per
    sample (~770 GB/s measured peer copy -> ~96 GS/s), which is below what one GPU does on resident data."""
    import torch
    import torch.distributed as dist

    import wavecap_sdr_b200._native as N
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
    from wavecap_sdr_b200.sharding import broadcast_block, frame_slab

    torch.cuda.set_device(local_rank)
    N.init(local_rank)
    if world > 1:
        init_nccl(local_rank)
    n = args.bcast_chunks * CHUNK
    ch = PolyphaseChannelizer(FS, BW)
    bufs = [torch.empty((n,), dtype=torch.complex64, device="cuda") for _ in range(2)]
    if rank == 0:
        g = torch.Generator(device="cuda").manual_seed(4321)
        for b in bufs:
            torch.view_as_real(b).normal_(0.0, 0.5, generator=g)
    comm = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    slab = frame_slab(ch.frames_for(n), world, rank)

    def post_broadcast(i):
        with torch.cuda.stream(comm):
            comm.wait_event(freed[i & 1])
            if world > 1:
                broadcast_block(torch.view_as_real(bufs[i & 1]), src=0)
            ready[i & 1].record(comm)

    def step(i, prefetch=True):
        main_stream.wait_event(ready[i & 1])
        if prefetch:
            post_broadcast(i + 1)
        rows, _ = ch.process_slab(bufs[i & 1], world, rank, fm=True)
        freed[i & 1].record(main_stream)
        return rows

    for e in freed:
        e.record(main_stream)
    total = max(3, args.warmup) + args.steps
    post_broadcast(0)
    rows = None
    for i in range(max(3, args.warmup)):
        rows = step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(max(3, args.warmup), total):
        rows = step(i, prefetch=(i + 1 < total))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n / (ms * 1e-3) / 1e6
    if rank == 0:
        peak, peak_src = load_peaks()
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MS/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "broadcast",
            "config": {
                "workload": "C5: ONE 125 MS/s capture, block NCCL-broadcast from rank 0, time-sharded 256-ch channelizer + FM",
                "block_samples": n, "frames_per_rank": slab.n_frames, "halo_frames": 9,
                "parallelism": f"broadcast + {world} time slabs; every rank receives the whole block",
                "bound": "NVLink ingress: 8 B/sample into every rank (measured peer copy 770 GB/s -> ~96 GS/s)",
                "l2": "block (%.0f MB) exceeds the 126 MB L2" % (8 * n / 1e6),
            },
            "link": {"bytes_received_per_rank_per_step": 8 * n if world > 1 else 0,
                     "achieved_gbs": round(8 * n / (ms * 1e-3) / 1e9, 1) if world > 1 else None},
            "gpu_launches": 2 * args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_pull(args, rank, local_rank, world, standalone=True):
    """ONE capture for the whole box without a data-path collective: rank 0's blocks (bcast-chunks x 6.25 M samples,
    double-buffered in ITS HBM = the ingest GPU) live in a sharding.PeerRegion mapped into every rank. Each rank's
    channelizer kernel reads its time slab (9-frame halo) straight out of that memory — for the other ranks the bulk
    async copies that feed the FIR cross NVLink, overlapped tile by tile with the FFT/discriminator of the previous
    tile — so a rank receives only what it processes and rank 0, which reads at HBM speed, keeps the larger share
    (`slab_weights`). Flags in the region order the ranks on their streams: rank 0 publishes READY[buffer] = block
    number, rank r publishes DONE[r][buffer] when its kernel has finished reading, rank 0 re-uses a buffer only after
    every DONE. Bound: rank 0's NVLink egress feeds all other ranks, so the job runs at about local + link rate."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import wavecap_sdr_b200._native as N
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
    from wavecap_sdr_b200.sharding import PeerRegion, frame_slab, slab_weights

    if standalone:
        torch.cuda.set_device(local_rank)
        N.init(local_rank)
        if world > 1:
            init_nccl(local_rank)
        bind_to_gpu_numa(local_rank)
    n = args.bcast_chunks * CHUNK
    local_rate, link_rate = (float(v) for v in args.pull_rates.split(","))
    if world == 1:
        weights = [1.0]
    elif args.pull_local_share > 0:
        a = min(args.pull_local_share, 1.0)
        weights = [a] + [(1.0 - a) / (world - 1)] * (world - 1)
    else:
        weights = slab_weights(world, local_rate, link_rate)
    ch = PolyphaseChannelizer(FS, BW)
    region = PeerRegion(2 * 8 * n, src=0)
    if rank == 0:
        g = torch.Generator(device="cuda").manual_seed(4321)
        for b in range(2):
            torch.view_as_real(region.payload_tensor(b * n, n)).normal_(0.0, 0.5, generator=g)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    blocks = [region.payload_tensor(b * n, n) if rank == 0 else region.span(b * n, n) for b in range(2)]
    timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    READY, DONE = 0, 16
    it = [0]           # block counter: sequence numbers only grow, also across the tuning rounds

    def step(kernel_events=None, trace=None):
        i = it[0]
        it[0] += 1
        b, seq = i & 1, i + 1
        if trace is not None:
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            marks[0].record()
        if rank == 0:
            if world > 1 and i >= 2:        # the buffer is free once every reader has finished block i - 2
                region.wait_flags(DONE + 2 + b, world - 1, seq - 2, timeout_ms=5000, timed_out=timed_out, stride=2)
            region.set_flag(READY + b, seq)  # (a live capture writes block i into the buffer before this)
        else:
            region.wait_flags(READY + b, 1, seq, timeout_ms=5000, timed_out=timed_out)
        if kernel_events is not None:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
        if trace is not None:
            marks[1].record()
        rows, _ = ch.process_slab(blocks[b], world, rank, fm=True, weights=weights)
        if trace is not None:
            marks[2].record()
        if kernel_events is not None:
            eb.record()
            kernel_events.append((ea, eb))
        if rank != 0:
            region.set_flag(DONE + 2 * rank + b, seq)
        if trace is not None:
            marks[3].record()
            trace.append(marks)
        return rows

    # share tuning: a rank's rate = its share / the time its own kernel took (flag waits excluded); new shares are
    # proportional to the rates, which equalises the finish times (the pullers share rank 0's egress, so a few rounds)
    tuning = []
    for _ in range(args.pull_tune_rounds if world > 1 and args.pull_local_share <= 0 else 0):
        evs = []
        for k in range(6):
            step(evs if k >= 2 else None)
        torch.cuda.synchronize()
        mine = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        t = torch.tensor([mine], device="cuda", dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        times = [max(float(x.item()), 1e-3) for x in every]
        rates = [w_ / t_ for w_, t_ in zip(weights, times)]
        tuning.append({"weights": [round(x, 4) for x in weights], "kernel_ms": [round(x, 4) for x in times]})
        weights = [r_ / sum(rates) for r_ in rates]
    slab = frame_slab(ch.frames_for(n), world, rank, weights=weights)

    # cross-check on the real link: a slab pulled by the kernel equals the same slab staged with a plain copy first
    check = None
    if rank != 0 and slab.n_frames > 0:
        probe = PolyphaseChannelizer(FS, BW)
        pulled, _ = probe.process_slab(blocks[0], world, rank, fm=True, weights=weights)
        staged = torch.empty((n,), dtype=torch.complex64, device="cuda")
        N.check(N.lib().wc_peer_copy(C.c_void_p(staged.data_ptr() + 8 * slab.sample0), C.c_void_p(blocks[0].ptr + 8 * slab.sample0),
                                     8 * slab.n_samples, N.torch_stream_ptr()))
        probe2 = PolyphaseChannelizer(FS, BW)
        local, _ = probe2.process_slab(staged, world, rank, fm=True, weights=weights)
        check = bool(torch.equal(pulled, local)) and bool(torch.isfinite(pulled).all()) and bool(pulled.abs().sum() > 0)
        del staged, pulled, local, probe, probe2
        torch.cuda.synchronize()
    ok = torch.tensor([1 if check in (None, True) else 0], device="cuda", dtype=torch.int32)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        raise SystemExit("pull mode: a slab read through peer memory differs from the staged copy")

    w = max(3, args.warmup)
    rows = None
    sampler = ClockSampler(local_rank)
    for _ in range(w):
        rows = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    # The ranks leave the host barrier up to a few ms apart — as long as the whole timed region of a short run. Three
    # more untimed steps let the READY/DONE flags bring the STREAMS back into lockstep (the hosts enqueue ~3x faster
    # than the GPUs execute), so the start event of every rank sits at the same block boundary.
    for _ in range(3):
        rows = step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for _ in range(args.steps):
        rows = step()
    e1.record()
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps     # host enqueue time per step (no sync inside)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    # untimed trace of a few more steps: where a rank's step goes (flag wait | slab kernels | flag set), stream time
    trace = []
    for _ in range(12):
        rows = step(trace=trace)
    torch.cuda.synchronize()
    seg = [sum(m[k].elapsed_time(m[k + 1]) for m in trace[6:]) / len(trace[6:]) for k in range(3)]
    per_step = trace[6][0].elapsed_time(trace[-1][0]) / (len(trace) - 7)
    tr = torch.tensor(seg + [per_step, host_ms], device="cuda", dtype=torch.float64)
    traces = [torch.zeros_like(tr) for _ in range(world)]
    if world > 1:
        dist.all_gather(traces, tr)
    else:
        traces = [tr]
    my_ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([my_ms], device="cuda", dtype=torch.float64)
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(timed_out, op=dist.ReduceOp.MAX)
    else:
        per_rank = [t]
    if int(timed_out.item()) != 0:
        raise SystemExit("pull mode: a flag wait timed out (a rank fell behind or died)")
    ms = max(float(x.item()) for x in per_rank)
    value = n / (ms * 1e-3) / 1e6
    if rank == 0:
        slabs = [frame_slab(ch.frames_for(n), world, r, weights=weights) for r in range(world)]
        pulled_bytes = sum(8 * s.n_samples for s in slabs[1:])
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MS/s", "n_gpus": world, "steps": args.steps,
            "warmup": w, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "pull",
            "config": {
                "workload": "C5: ONE 125 MS/s capture resident on rank 0, mapped into every rank (CUDA IPC over NVLink); each rank's "
                            "256-ch channelizer + FM kernel pulls its own time slab while it computes",
                "block_samples": n, "frames_per_rank": [s.n_frames for s in slabs], "halo_frames": 9,
                "slab_weights": [round(x, 4) for x in weights], "share_tuning": tuning,
                "parallelism": f"{world} weighted time slabs, no data-path collective; READY/DONE flags in peer memory",
                "bound": "rank 0: its own HBM; the other ranks together: rank 0's NVLink egress (8 B/sample)",
                "l2": "block (%.0f MB) exceeds the 126 MB L2" % (8 * n / 1e6),
                "checked": "peer-pulled slab == staged-copy slab on every rank" if world > 1 else "single rank",
            },
            "link": {"bytes_pulled_from_rank0_per_step": pulled_bytes,
                     "rank0_egress_gbs": round(pulled_bytes / (ms * 1e-3) / 1e9, 1) if world > 1 else None,
                     "ms_per_step_by_rank": [round(float(x.item()), 4) for x in per_rank],
                     "trace_ms_by_rank[wait,slab,set,step,host_enqueue]": [[round(float(v), 4) for v in x.tolist()] for x in traces]},
            "gpu_launches": (4 if world > 1 else 3) * args.steps, "clocks": clocks,
        }
        if standalone:
            print(json.dumps(line), flush=True)
    else:
        line = None
    del rows
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    region.close()
    if world > 1 and standalone:
        dist.destroy_process_group()
    return line


def run_stripe(args, rank, local_rank, world, standalone=True):
    """ONE capture for the whole box, striped at ingest (sharding.StripedCapture): every block (stripe-chunks x 6.25 M samples
    = one process() call) is cut into N time slabs and slab r is resident in rank r's HBM when the timed region starts, as
    if each GPU's own PCIe link / DMA target had received it. Per block a rank fetches the 9 halo rows (9 KB) from the
    previous slab's owner over NVLink (rank 0: the 10 tail rows of the previous block from the last rank -> carried
    history), then runs the SAME fused kernel on [halo | slab]. No rank's egress carries the capture, so the job scales
    with N; frames are bit-equal to the unsharded call (tests/test_stripe_gpu.py). `e2e` repeats it with the ingest inside
    the timed region: every rank copies its slab from pinned host memory first (PCIe-bound)."""
    import torch
    import torch.distributed as dist

    import wavecap_sdr_b200._native as N
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
    from wavecap_sdr_b200.sharding import StripedCapture

    if standalone:
        torch.cuda.set_device(local_rank)
        N.init(local_rank)
        if world > 1:
            init_nccl(local_rank)
        bind_to_gpu_numa(local_rank)
    n = args.stripe_chunks * CHUNK
    ch = PolyphaseChannelizer(FS, BW)
    sc = StripedCapture(ch, n)
    g = torch.Generator(device="cuda").manual_seed(4321 + rank)
    for b in range(2):
        torch.view_as_real(sc.own_tensor(b)).normal_(0.0, 0.5, generator=g)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host = None

    def step(ingest=False):
        b = sc.wait_free()
        if ingest:
            sc.own_tensor(b).copy_(host, non_blocking=True)
        sc.publish()
        rows, _ = sc.process(fm=True)
        return rows

    def timed(k, ingest=False):
        barrier()
        for _ in range(3):                      # re-align the streams after the host barrier (see run_pull)
            rows = step(ingest)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            rows = step(ingest)
        if ingest:
            float(rows[-1, 17].item())          # the step's result is read on the host
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), rows

    w = max(3, args.warmup)
    for _ in range(w):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, rows = timed(args.steps)
    clocks = sampler.stop()
    finite = bool(torch.isfinite(rows).all()) and bool(rows.abs().sum() > 0)
    # e2e: the ingest inside the timed region
    host = torch.empty((sc.own_n[rank],), dtype=torch.complex64, pin_memory=True)
    torch.view_as_real(host).fill_(0.25)
    e_steps = max(2, min(args.steps, 5))
    e_ms, _ = timed(e_steps, ingest=True)
    sc.check()
    ok = torch.tensor([1 if finite else 0], device="cuda", dtype=torch.int32)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(n / (ms * 1e-3) / 1e6, 1), "unit": "MS/s", "n_gpus": world, "steps": args.steps,
            "warmup": w, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "stripe",
            "config": {
                "workload": "C5: ONE 125 MS/s capture striped at ingest: slab r of every block resident in rank r's HBM; 256-ch "
                            "channelizer + FM on [9-row halo from the previous slab's owner | own slab]",
                "block_samples": n, "frames_per_rank": [s_.n_frames for s_ in sc.slabs], "halo_rows": 9,
                "parallelism": f"{world} equal time slabs, no data-path collective; 9 KB halo per rank per block over NVLink, "
                               "READY/DONE flags in peer memory",
                "bound": "each rank: its own HBM / SMs (the single-GPU kernel rate); NVLink carries 9 KB per rank per block",
                "l2": "slab (%.0f MB per rank) exceeds the 126 MB L2" % (8 * n / world / 1e6),
                "checked": "finite, non-zero output on every rank" if int(ok.item()) == 1 else "OUTPUT CHECK FAILED",
            },
            "link": {"halo_bytes_per_rank_per_step": 8 * 9 * 128 if world > 1 else 0},
            "e2e": {"value": round(n / (e_ms * 1e-3) / 1e6, 1), "unit": "MS/s", "ms_per_step": round(e_ms, 3),
                    "h2d_bytes_per_step": int(8 * n), "d2h_bytes_per_step": 4,
                    "what": "every rank copies its slab from pinned host memory into its own buffer first (its own PCIe link)"},
            "gpu_launches": (5 if world > 1 else 3) * args.steps, "clocks": clocks,
        }
        if standalone:
            print(json.dumps(line), flush=True)
    del rows
    barrier()
    sc.close()
    if world > 1 and standalone:
        dist.destroy_process_group()
    return line


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.mode == "pull":
        run_pull(args, rank, local_rank, world)
        return
    if args.mode == "stripe":
        run_stripe(args, rank, local_rank, world)
        return
    if args.mode == "broadcast":
        run_broadcast(args, rank, local_rank, world)
        return

    # CPU baseline first (rank 0, N=1 only), before CUDA is touched in this process.
    cpu, cfg_cpu = None, {}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(args.cpu_samples, os.cpu_count() or 1, args.cpu_seconds)
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "MS/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        if not args.no_configs:
            cfg_cpu = config_cpu_baselines(os.cpu_count() or 1)

    import torch
    import torch.distributed as dist

    import wavecap_sdr_b200._native as N
    from wavecap_sdr_b200.dsp.channelizer import OUT_FM, PolyphaseChannelizer, fm_scale

    torch.cuda.set_device(local_rank)
    N.init(local_rank)
    numa = bind_to_gpu_numa(local_rank)  # after the cpu_baseline leg, which uses every host core
    if world > 1:
        init_nccl(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = N.lib()
    ch = PolyphaseChannelizer(FS, BW)
    frames = ch.frames_for(CHUNK)
    scale = fm_scale(int(ch.channel_sample_rate))
    nb = args.chunks
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    # synthetic IQ as benchmark_dsp.py:119 (randn * 0.5), generated in slabs to bound temporaries
    x = torch.empty((nb * CHUNK,), dtype=torch.complex64, device="cuda")
    xv = torch.view_as_real(x)
    for i in range(nb):
        xv[i * CHUNK:(i + 1) * CHUNK].normal_(0.0, 0.5, generator=g)
    out = torch.empty((nb * frames, 256), dtype=torch.float32, device="cuda")
    stream = N.torch_stream_ptr()

    def step():
        N.check(lib.wc_chan_process(ch._h, C.c_void_p(x.data_ptr()), CHUNK, nb, CHUNK, OUT_FM, scale,
                                    C.c_void_p(out.data_ptr()), stream))

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    samples_per_step = nb * CHUNK
    value = world * samples_per_step / (ms * 1e-3) / 1e6
    checksum = float(out[:: max(1, out.shape[0] // 997)].double().abs().sum().item())

    # sustained: the K timed steps above are a burst (tens of ms); this kernel is FMA/issue-bound, so its rate follows
    # the SM clock, which settles lower under seconds of load. Same launches back to back for >= --sustained-seconds, the
    # clock sampled inside; reported beside the burst number, not instead of it.
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / ms) + 1)
        barrier()
        s_sampler = ClockSampler(local_rank)
        s_sampler.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        barrier()
        s_clocks = s_sampler.stop()
        s_ms = s0.elapsed_time(s1) / n_sus
        ts = torch.tensor([s_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        s_ms = float(ts.item())
        sustained = {"value": round(world * samples_per_step / (s_ms * 1e-3) / 1e6, 1), "unit": "MS/s", "steps": n_sus,
                     "seconds": round(n_sus * s_ms * 1e-3, 2), "ms_per_step": round(s_ms, 4), "clocks": s_clocks}

    # secondary, informational: the channelizer alone (mode 0 = what PolyphaseChannelizer.process() returns,
    # complex64 frames, 24 algorithmic bytes per input sample) on a sub-batch that fits beside the FM buffers
    from wavecap_sdr_b200.dsp.channelizer import OUT_COMPLEX

    nb0 = max(1, min(nb, 16))
    del out
    outc = torch.empty((nb0 * frames, 256), dtype=torch.complex64, device="cuda")

    def step0():
        N.check(lib.wc_chan_process(ch._h, C.c_void_p(x.data_ptr()), CHUNK, nb0, CHUNK, OUT_COMPLEX, 0.0,
                                    C.c_void_p(outc.data_ptr()), stream))

    for _ in range(3):
        step0()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(5):
        step0()
    c1.record()
    torch.cuda.synchronize()
    ms0 = c0.elapsed_time(c1) / 5
    chan_only_msps = nb0 * CHUNK / (ms0 * 1e-3) / 1e6
    del outc

    # ---- e2e leg: host buffers through the reference-facing C-ABI call -------------------------
    eb = args.e2e_chunks
    e_steps = args.e2e_steps or min(args.steps, 5)
    hx = N.pinned_empty((eb * CHUNK,), "complex64")
    hx.view("float32")[:] = 0.25
    hx[: CHUNK] = x[:CHUNK].cpu().numpy()
    for i in range(1, eb):
        hx[i * CHUNK:(i + 1) * CHUNK] = hx[:CHUNK]
    hout = N.pinned_empty((eb * frames, 256), "float32")
    ch2 = PolyphaseChannelizer(FS, BW)

    def e2e_step():
        N.check(lib.wc_chan_process_host(ch2._h, N.np_ptr(hx), CHUNK, eb, OUT_FM, scale, N.np_ptr(hout)))
        return float(hout[-1, 17])  # the step's result is read on the host

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
    te = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_ms = float(te.item())
    e2e_value = world * eb * CHUNK / (e_ms * 1e-3) / 1e6
    g_sub = max(1, min(eb, (4 << 20) // CHUNK))  # sub-batches of the pipelined host call (channelizer.cu)

    # the same end-to-end call in the formats that shrink the PCIe traffic: int16 capture input (4 B/sample up) and the
    # fused nbfm audio /20 output (0.4 B/sample down) — SURVEY §8d "C5 + audio"; the cf32 -> FM leg above stays the headline
    from wavecap_sdr_b200.dsp.channelizer import IN_CF32, IN_CS16, OUT_AUDIO

    e2e_modes = {}
    try:
        if args.no_modes:
            raise RuntimeError("skipped (--no-modes)")
        hq = N.pinned_empty((eb * CHUNK, 2), "int16")
        hq[:CHUNK] = np.clip(np.stack([hx[:CHUNK].real, hx[:CHUNK].imag], axis=1) * 8192.0, -32767, 32767).astype(np.int16)
        for i in range(1, eb):
            hq[i * CHUNK:(i + 1) * CHUNK] = hq[:CHUNK]
        N.check(lib.wc_chan_audio_config(ch2._h, 976560, 48828))
        n_audio = int(lib.wc_chan_audio_len(ch2._h, CHUNK))
        haud = N.pinned_empty((eb * n_audio, 256), "float32")
        for name, fmt, src, mode, dst in (("cs16_to_fm", IN_CS16, hq, OUT_FM, hout), ("cf32_to_audio", IN_CF32, hx, OUT_AUDIO, haud),
                                          ("cs16_to_audio", IN_CS16, hq, OUT_AUDIO, haud)):
            def m_step():
                N.check(lib.wc_chan_process_host_ex(ch2._h, N.np_ptr(src), fmt, CHUNK, eb, mode, scale, N.np_ptr(dst)))
                return float(dst[-1, 17])
            m_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                m_step()
            torch.cuda.synchronize()
            m_ms = (time.perf_counter() - t0) * 1e3 / e_steps
            tm = torch.tensor([m_ms], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            m_ms = float(tm.item())
            e2e_modes[name] = {"value": round(world * eb * CHUNK / (m_ms * 1e-3) / 1e6, 1), "unit": "MS/s", "ms_per_step": round(m_ms, 3),
                               "h2d_bytes_per_step": int(src.nbytes), "d2h_bytes_per_step": int(dst.nbytes)}
        # device-resident rate of the audio mode (its own roofline: 8 B in + 2 x 4/20 B out per input sample)
        nba = min(nb, 32)
        daud = torch.empty((nba * n_audio, 256), dtype=torch.float32, device="cuda")

        def a_step():
            N.check(lib.wc_chan_process_ex(ch2._h, C.c_void_p(x.data_ptr()), IN_CF32, CHUNK, nba, CHUNK, OUT_AUDIO, scale,
                                           C.c_void_p(daud.data_ptr()), stream))
        for _ in range(3):
            a_step()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(5):
            a_step()
        a1.record()
        torch.cuda.synchronize()
        a_ms = a0.elapsed_time(a1) / 5
        audio_msps = nba * CHUNK / (a_ms * 1e-3) / 1e6
        e2e_modes["audio_device_resident"] = {"msps_per_gpu": round(audio_msps, 1), "alg_bytes_per_sample": 8.4,
                                              "achieved_gbs": round(8.4 * audio_msps / 1e3, 1), "ms_per_step": round(a_ms, 3),
                                              "chunks_per_step": nba,
                                              "note": "FM kernel -> discriminator rows in HBM -> /20 polyphase decimator + RMS/clip: "
                                                      "moves 24.4 B/sample, compute-bound by construction (SURVEY §8d)"}
        del daud
    except Exception as e:
        e2e_modes["error"] = f"{type(e).__name__}: {e}"

    # the ceiling the e2e leg runs under: the same bytes moved by plain pinned cudaMemcpyAsync, H2D and D2H concurrently on
    # two streams, all ranks at once, no kernel — what this box's host<->device path gives the whole job
    ceiling = None
    try:
        d_in = torch.empty((eb * CHUNK,), dtype=torch.complex64, device="cuda")
        d_o = torch.empty((eb * frames, 256), dtype=torch.float32, device="cuda")
        t_in, t_o = torch.from_numpy(hx), torch.from_numpy(hout)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_step():
            with torch.cuda.stream(s_up):
                d_in.copy_(t_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                t_o.copy_(d_o, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()

        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copy_step()
        c_ms = (time.perf_counter() - t0) * 1e3 / 3
        tc = torch.tensor([c_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        c_ms = float(tc.item())
        ceiling = {"value": round(world * eb * CHUNK / (c_ms * 1e-3) / 1e6, 1), "unit": "MS/s", "ms_per_step": round(c_ms, 3),
                   "h2d_plus_d2h_gbs_per_gpu": round((hx.nbytes + hout.nbytes) / (c_ms * 1e-3) / 1e9, 1),
                   "what": "plain pinned cudaMemcpyAsync of the step's input and output, both directions concurrently, every rank "
                           "at once, no kernel"}
        del d_in, d_o
    except Exception as e:
        ceiling = {"error": f"{type(e).__name__}: {e}"}

    # the reference-facing Python call itself: PolyphaseChannelizer.process(numpy chunk) -> list of per-frame arrays
    # (pageable input like the reference's callers hand over, 48 827 row views built on return)
    api = None
    if rank == 0:
        try:
            one = np.ascontiguousarray(hx[:CHUNK])
            ch3 = PolyphaseChannelizer(FS, BW)
            ch3.process(one)
            t0 = time.perf_counter()
            for _ in range(2):
                res = ch3.process(one)
            a_ms = (time.perf_counter() - t0) * 1e3 / 2
            api = {"value": round(CHUNK / (a_ms * 1e-3) / 1e6, 1), "unit": "MS/s", "ms_per_chunk": round(a_ms, 2),
                   "frames_returned": len(res), "what": "PolyphaseChannelizer.process(complex64 ndarray of one 50 ms chunk) -> list of "
                   "48 827 complex64[256] arrays, pageable host memory, one GPU"}
            del res, ch3
        except Exception as e:
            api = {"error": f"{type(e).__name__}: {e}"}

    # ---- C1..C4 (N=1 only): the other BASELINE configs, each with its own roofline, clocks and CPU arm ----------
    cfg_recs = None
    if rank == 0 and world == 1 and not args.no_configs:
        del x
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs

            cfg_recs = bench_configs.headline_configs(ClockSampler)
            for r in cfg_recs:
                r["cpu_baseline"] = cfg_cpu.get(r.get("config"))
                cb = r["cpu_baseline"]
                if cb and cb.get("value") and r.get("value"):
                    r["gpu_over_cpu"] = round(r["value"] / cb["value"], 1)
        except Exception as e:
            cfg_recs = [{"error": f"{type(e).__name__}: {e}"}]

    # ---- N>1: the north-star-literal split — ONE capture over all ranks (pull mode) — beside the replicas headline
    one_capture = None
    if world > 1 and not args.no_one_capture:
        try:
            del x
        except NameError:
            pass
        torch.cuda.empty_cache()
        barrier()
        one_capture = {}
        for name, fn in (("stripe", run_stripe), ("pull", run_pull)):
            try:
                oc = fn(args, rank, local_rank, world, standalone=False)
                if rank == 0 and oc is not None:
                    rec = {k: oc[k] for k in ("mode", "value", "unit", "ms_per_step", "steps", "scaling", "link", "clocks") if k in oc}
                    rec["config"] = {k: oc["config"][k] for k in ("workload", "block_samples", "slab_weights", "parallelism", "bound", "checked")
                                     if k in oc["config"]}
                    if "e2e" in oc:
                        rec["e2e"] = oc["e2e"]
                    one_capture[name] = rec
            except BaseException as e:  # noqa: BLE001 — SystemExit from a failed cross-check included
                one_capture[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
            barrier()
        if rank == 0:
            best = max((v for v in one_capture.values() if "value" in v), key=lambda v: v["value"], default=None)
            if best is not None:
                one_capture["value"], one_capture["unit"], one_capture["mode"] = best["value"], best["unit"], best["mode"]

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = ALG_BYTES_PER_SAMPLE * samples_per_step / (ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "chan_fm_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            # ncu dram bytes per input sample of the profiled launch, scaled to this launch
            traffic = round(tj["dram_bytes_per_sample"] * samples_per_step)
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MS/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "chunks_per_step": nb, "samples_per_step_per_gpu": samples_per_step,
                "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed"
                      % ((8 + 8) * samples_per_step / 1e9),
                "parallelism": f"{world} independent captures, one per GPU, no data-path collective",
                "e2e_chunks_per_step": eb, "e2e_steps": e_steps, "checksum": checksum, "host_affinity": numa,
            },
            "roofline": {
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": "wc::chan256p_kernel<1>", "alg_bytes_per_sample": ALG_BYTES_PER_SAMPLE,
                "channelizer_only": {"kernel": "wc::chan256p_kernel<0>", "alg_bytes_per_sample": 24,
                                     "msps_per_gpu": round(chan_only_msps, 1),
                                     "achieved": round(24 * chan_only_msps / 1e3, 1),
                                     "frac": round(24 * chan_only_msps / 1e3 / peak, 4),
                                     "note": "mode 0 = complex64 frames exactly as PolyphaseChannelizer.process() "
                                             "returns them, no discriminator; informational"},
            },
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_value, 1), "unit": "MS/s", "ms_per_step": round(e_ms, 3),
                    "h2d_bytes_per_step": int(eb * CHUNK * 8), "d2h_bytes_per_step": int(eb * frames * 256 * 4),
                    "copy_ceiling": ceiling,
                    "frac_of_copy_ceiling": round(e2e_value / ceiling["value"], 3) if ceiling and ceiling.get("value") else None,
                    "python_process_api": api, "modes": e2e_modes},
            "sustained": dict(sustained, roofline_frac=round(ALG_BYTES_PER_SAMPLE * sustained["value"] / world / 1e3 / peak, 4)) if sustained else None,
            "configs": cfg_recs,
            "one_capture": one_capture,
            "gpu_launches": 2 * args.steps + 2 * ((eb + g_sub - 1) // g_sub) * (e_steps + 1),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
