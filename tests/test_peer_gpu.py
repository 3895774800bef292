"""GPU: one capture pulled by several processes out of the ingest process' memory (sharding.PeerRegion over wc_peer_* /
wc_flag_*). A second PROCESS on the same B200 maps the region (CUDA IPC; on a multi-GPU box the same calls go over
NVLink), waits for the producer's flag, channelizes + FM-demodulates its weighted time slab straight out of the
mapped memory and signals completion; parent slab + child slab must equal the unsharded call bit for bit, across two
blocks (carried history) whose contents differ."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_SAMPLES = 256 + 128 * 1999 + 31
WEIGHTS = [0.68, 0.32]
READY, DONE = 0, 8          # flag words: READY + buffer, DONE + buffer


def _child(handle, payload_bytes, q):
    try:
        import torch

        import wavecap_sdr_b200._native as N
        from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
        from wavecap_sdr_b200.sharding import PeerRegion

        torch.cuda.set_device(0)
        N.init(0)
        region = PeerRegion.open(handle, payload_bytes)
        ch = PolyphaseChannelizer(125_000_000, 488281)
        timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        parts = []
        for blk in range(2):
            region.wait_flags(READY + blk, 1, blk + 1, timeout_ms=20000, timed_out=timed_out)
            rows, f0 = ch.process_slab(region.span(blk * STRIDE, N_SAMPLES), 2, 1, fm=True, weights=WEIGHTS)
            region.set_flag(DONE + blk, blk + 1)
            parts.append((rows.cpu().numpy(), f0))
        torch.cuda.synchronize()
        hist = ch.arm_history.copy()
        region.close()
        q.put(("ok", parts, int(timed_out.item()), hist))
    except Exception as e:  # noqa: BLE001 - reported to the parent
        import traceback

        q.put(("error", traceback.format_exc(), repr(e), None))


STRIDE = (N_SAMPLES + 511) // 512 * 512


def test_peer_region_two_processes_weighted_slabs(native):
    import torch
    import torch.multiprocessing as mp

    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer
    from wavecap_sdr_b200.sharding import PeerRegion

    payload = 8 * STRIDE * 2
    region = PeerRegion(payload)
    assert region.is_owner and region.world == 1
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        p = ctx.Process(target=_child, args=(region.handle, payload, q))
        p.start()
        g = torch.Generator(device="cuda").manual_seed(99)
        whole = PolyphaseChannelizer(125_000_000, 488281)
        mine = PolyphaseChannelizer(125_000_000, 488281)
        exp, got0 = [], []
        timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        for blk in range(2):
            t = region.payload_tensor(blk * STRIDE, N_SAMPLES)
            torch.view_as_real(t).normal_(0.0, 0.5, generator=g)      # the "ingest"
            region.set_flag(READY + blk, blk + 1)                     # publish after the fill, on the same stream
            exp.append(whole.process_fm(t))
            rows, f0 = mine.process_slab(t, 2, 0, fm=True, weights=WEIGHTS)
            assert f0 == 0
            got0.append(rows)
        region.wait_flags(DONE, 2, 1, timeout_ms=60000, timed_out=timed_out)    # both buffers read at least once
        region.wait_flags(DONE + 1, 1, 2, timeout_ms=60000, timed_out=timed_out)
        status, parts, child_timeout, hist = q.get(timeout=180)
        p.join(timeout=60)
        assert status == "ok", parts
        torch.cuda.synchronize()
        assert child_timeout == 0 and int(timed_out.item()) == 0
        for blk in range(2):
            rows1, f1 = parts[blk]
            full = exp[blk].cpu().numpy()
            a = got0[blk].cpu().numpy()
            assert a.shape[0] == f1 and a.shape[0] + rows1.shape[0] == full.shape[0]
            assert abs(a.shape[0] / full.shape[0] - WEIGHTS[0]) < 0.01
            assert np.array_equal(np.concatenate([a, rows1]), full)
        assert np.array_equal(hist, whole.arm_history) and np.array_equal(mine.arm_history, whole.arm_history)
    finally:
        region.close()


def test_flag_wait_times_out_instead_of_hanging(native):
    import torch

    from wavecap_sdr_b200.sharding import PeerRegion

    region = PeerRegion(4096)
    try:
        timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        region.set_flag(3, 7)
        region.wait_flags(3, 1, 7, timeout_ms=50, timed_out=timed_out)      # already there
        torch.cuda.synchronize()
        assert int(timed_out.item()) == 0
        region.wait_flags(4, 1, 1, timeout_ms=50, timed_out=timed_out)      # never set
        torch.cuda.synchronize()
        assert int(timed_out.item()) == 1
        assert int(region.flags_tensor()[3].item()) == 7
        # without an explicit flag the region's own one records the time-out and check() raises (then re-arms)
        region.wait_flags(3, 1, 7, timeout_ms=50)
        region.check()
        region.wait_flags(4, 1, 1, timeout_ms=50)
        with pytest.raises(TimeoutError):
            region.check()
        region.check()
    finally:
        region.close()
