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
