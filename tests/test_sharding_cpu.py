"""CPU: host-side multi-GPU logic — partition helpers, and a world-size-2 gloo run of the broadcast + time-slab
pipeline in which every rank channelizes its slab with the oracle and the gathered result equals the unsharded one."""
import os
import socket

import numpy as np
import pytest

from wavecap_sdr_b200.sharding import CHAN_HALO_FRAMES, frame_slab, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 256, 48827):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_frame_slab_geometry():
    F = 48827
    for world in (1, 2, 4, 8):
        slabs = [frame_slab(F, world, r) for r in range(world)]
        assert sum(s.n_frames for s in slabs) == F
        for s in slabs:
            assert s.skip == min(CHAN_HALO_FRAMES, s.f0) and s.sample0 == s.start_frame * 128
            assert s.sample0 + s.n_samples == (s.f1 - 1) * 128 + 256        # last emitted frame's last sample
    assert frame_slab(3, 8, 7).n_frames == 0 and frame_slab(3, 8, 7).n_samples == 0


def test_weighted_slabs():
    from wavecap_sdr_b200.sharding import slab_weights, weighted_range

    w = slab_weights(8, 207.0, 96.0)
    assert abs(sum(w) - 1.0) < 1e-12 and abs(w[0] - 207.0 / 303.0) < 1e-12 and len(set(w[1:])) == 1
    assert slab_weights(1, 207.0, 96.0) == [1.0]
    F = 390624
    for weights in (w, [1, 1, 1], [0.0, 1.0], [5, 0, 1]):
        world = len(weights)
        r = [weighted_range(F, weights, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == F and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        tot = float(sum(weights))
        assert all(abs((b - a) - F * wk / tot) <= 1.0 for (a, b), wk in zip(r, weights))
        slabs = [frame_slab(F, world, k, weights=weights) for k in range(world)]
        assert sum(s.n_frames for s in slabs) == F
        for s in slabs:
            if s.n_frames:
                assert s.skip == min(CHAN_HALO_FRAMES, s.f0) and s.sample0 + s.n_samples == (s.f1 - 1) * 128 + 256
            else:
                assert s.n_samples == 0
    assert [weighted_range(10, [1, 1], k) for k in range(2)] == [shard_range(10, 2, k) for k in range(2)]
    with pytest.raises(ValueError):
        weighted_range(10, [0, 0], 0)
    with pytest.raises(ValueError):
        frame_slab(10, 2, 0, weights=[1.0])


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist

    from oracle.channelizer import ChannelizerOracle, channelize_fm
    from wavecap_sdr_b200.sharding import broadcast_block, frame_slab

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        block = torch.zeros((n, 2), dtype=torch.float32)
        if rank == 0:
            rng = np.random.default_rng(77)
            block = torch.from_numpy(np.stack([rng.standard_normal(n), rng.standard_normal(n)], axis=1).astype(np.float32) * 0.5)
        broadcast_block(block, src=0)
        x = block.numpy().view(np.complex64).reshape(-1)
        F = (n - 256) // 128 + 1
        s = frame_slab(F, world, rank)
        o = ChannelizerOracle(125_000_000, 488281)
        y = o.process_vectorized(x[s.sample0:s.sample0 + s.n_samples])
        rate = int(o.channel_sample_rate)
        fm = channelize_fm(y, rate)[s.skip:]
        mine = torch.from_numpy(np.ascontiguousarray(fm))
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]], dtype=torch.int64))
        parts = [torch.zeros((int(k.item()), 256), dtype=torch.float32) for k in sizes]
        dist.all_gather(parts, mine) if len({int(k.item()) for k in sizes}) == 1 else None
        if rank == 0:
            q.put((x.copy(), [int(k.item()) for k in sizes], mine.numpy().copy(), s.f0, s.f1))
        else:
            q.put((None, None, mine.numpy().copy(), s.f0, s.f1))
    finally:
        dist.destroy_process_group()


def test_broadcast_time_slabs_gloo_world2():
    import torch.multiprocessing as mp

    from oracle.channelizer import ChannelizerOracle, channelize_fm

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    world, n = 2, 256 + 128 * 99
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = next(r[0] for r in res if r[0] is not None)
    o = ChannelizerOracle(125_000_000, 488281)
    full = channelize_fm(o.process_vectorized(x), int(o.channel_sample_rate))
    got = np.zeros_like(full)
    for _, _, part, f0, f1 in res:
        got[f0:f1] = part
    # slab 0 starts at frame 0 with the same (zero) history as the unsharded run; slab 1 is halo-complete
    assert np.array_equal(got, full)


# ---- one capture striped at ingest (sharding.stripe_layout / StripedCapture): host logic on CPU, gloo world 2 ------------

def test_stripe_layout_geometry():
    from wavecap_sdr_b200.sharding import stripe_layout

    for frames, world in ((48827, 8), (3124999, 8), (1500, 3), (390624, 2), (97, 1)):
        L = stripe_layout(frames, world)
        assert L.f0[0] == 0 and L.f1[-1] == frames and all(a == b for a, b in zip(L.f1[:-1], L.f0[1:]))
        for r in range(world):
            assert L.own0[r] == 128 * L.f0[r] and L.own0[r] + L.own_n[r] == 128 * (L.f1[r] - 1) + 256
            if r:
                # the halo sits entirely inside the previous rank's own part and ends where this rank's own part begins
                assert 0 <= L.halo_src(r) and L.halo_src(r) + L.halo == L.own0[r] - L.own0[r - 1] <= L.own_n[r - 1]
        assert L.tail_n == 1280 and 0 <= L.tail_src() and L.tail_src() + L.tail_n == L.own_n[-1]
    with pytest.raises(ValueError):
        stripe_layout(40, 8)


def _stripe_worker(rank, world, port, n, q):
    """every rank holds ONLY its own slab of each block; halo / tail travel by send/recv; the oracle stands in for the kernel"""
    import torch
    import torch.distributed as dist

    from oracle.channelizer import ChannelizerOracle, channelize_fm
    from wavecap_sdr_b200.sharding import stripe_layout

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F = (n - 256) // 128 + 1
        L = stripe_layout(F, world)
        out = []
        prev_tail = None
        for blk in range(2):
            rng = np.random.default_rng(500 + blk)                    # the "ingest": every rank cuts its slab out of the block
            full = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
            own = full[L.own0[rank]: L.own0[rank] + L.own_n[rank]].copy()
            del full
            o = ChannelizerOracle(125_000_000, 488281)
            rate = int(o.channel_sample_rate)
            # message order without cycles: receive my halo, serve the next rank's, the last rank ships the block's tail to
            # rank 0 (which collects it after its own work)
            halo = None
            if rank > 0:
                buf = torch.zeros(2 * L.halo, dtype=torch.float32)
                dist.recv(buf, src=rank - 1)
                halo = buf.numpy().view(np.complex64).copy()
            if rank + 1 < world:                                      # serve the next rank's halo out of my own part
                h = own[L.halo_src(rank + 1): L.halo_src(rank + 1) + L.halo]
                dist.send(torch.from_numpy(np.ascontiguousarray(h).view(np.float32).copy()), dst=rank + 1)
            if rank == world - 1 and world > 1:                       # the block's tail goes to rank 0 for its next block
                t = own[L.tail_src(): L.tail_src() + L.tail_n]
                dist.send(torch.from_numpy(np.ascontiguousarray(t).view(np.float32).copy()), dst=0)
            if rank == 0:
                if prev_tail is not None:
                    # carried history: arm_history[:, j] = the 256-sample block fed j frames ago = tail rows (8 - j, 9 - j)
                    hist = np.stack([prev_tail[(8 - j) * 128: (8 - j) * 128 + 256] for j in range(9)], axis=1)
                    o.arm_history = hist.astype(np.complex64)
                rows = channelize_fm(o.process_vectorized(own), rate)
                if world > 1:
                    buf = torch.zeros(2 * L.tail_n, dtype=torch.float32)
                    dist.recv(buf, src=world - 1)
                    prev_tail = buf.numpy().view(np.complex64).copy()
                else:
                    prev_tail = own[L.tail_src(): L.tail_src() + L.tail_n].copy()
            else:
                local = np.concatenate([halo, own])
                rows = channelize_fm(o.process_vectorized(local), rate)[9:]
            out.append((rows, L.f0[rank]))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_striped_capture_gloo_world2():
    import torch.multiprocessing as mp

    from oracle.channelizer import ChannelizerOracle, channelize_fm

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    world, n = 2, 256 + 128 * 149 + 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stripe_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = ChannelizerOracle(125_000_000, 488281)
    for blk in range(2):
        rng = np.random.default_rng(500 + blk)
        full = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
        exp = channelize_fm(o.process_vectorized(full), int(o.channel_sample_rate))
        got = np.concatenate([res[r][blk][0] for r in range(world)])
        assert got.shape == exp.shape and np.array_equal(got, exp), f"block {blk}"
