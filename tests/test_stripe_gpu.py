"""GPU: ONE capture striped over several ranks at ingest (sharding.StripedCapture). W processes on the same B200 stand in
for W GPUs (CUDA IPC maps every rank's region into the others; on a multi-GPU box the same calls go over NVLink): every
rank holds only ITS time slab of each block, fetches the 9 halo rows from the previous slab's owner (rank 0: the 10 tail
rows of the previous block from the last rank) and channelizes + FM-demodulates its frames. The concatenation over ranks
must equal the unsharded `process_fm()` calls bit for bit, over three blocks (carried history, both buffers reused)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BLOCK = 256 + 128 * 1499 + 77           # 1500 frames per block, ragged tail like a real chunk
N_BLOCKS = 3


def _block(i):
    rng = np.random.default_rng(700 + i)
    return ((rng.standard_normal(BLOCK) + 1j * rng.standard_normal(BLOCK)) * 0.5).astype(np.complex64)


def _rank_main(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist

        import wavecap_sdr_b200._native as N
        from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
        from wavecap_sdr_b200.sharding import StripedCapture

        torch.cuda.set_device(0)
        N.init(0)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        ch = PolyphaseChannelizer(125_000_000, 488281)
        sc = StripedCapture(ch, BLOCK)
        parts = []
        for i in range(N_BLOCKS):
            b = sc.wait_free(timeout_ms=30000)
            sc.own_tensor(b).copy_(torch.from_numpy(sc.slab_of(_block(i))))        # the ingest: only this rank's slab
            sc.publish()
            rows, f0 = sc.process(fm=True, timeout_ms=30000)
            parts.append((rows.cpu().numpy(), f0))
        sc.check()
        dist.barrier()
        sc.close()
        dist.destroy_process_group()
        q.put((rank, "ok", parts))
    except Exception:  # noqa: BLE001 - reported to the parent
        import traceback

        q.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_striped_capture_equals_the_unsharded_calls(native, world):
    import socket

    import torch.multiprocessing as mp

    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    for _ in range(world):
        rank, status, payload = q.get(timeout=300)
        assert status == "ok", payload
        got[rank] = payload
    [p.join(timeout=60) for p in procs]
    whole = PolyphaseChannelizer(125_000_000, 488281)
    for i in range(N_BLOCKS):
        exp = whole.process_fm(_block(i))
        rows = np.concatenate([got[r][i][0] for r in range(world)])
        f0s = [got[r][i][1] for r in range(world)]
        assert f0s[0] == 0 and f0s == sorted(f0s) and rows.shape == exp.shape
        assert np.array_equal(rows, exp), f"block {i}: striped frames differ from the unsharded call"
