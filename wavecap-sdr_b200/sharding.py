"""Host-side partitioning for multi-GPU runs (one process per GPU, torch.distributed for plumbing).

The hot path shards without any data-path collective:
  * independent captures / replicas: rank r owns capture r                      -> `replica_seed`
  * channels of one capture (analog chain, P25 banks, DDC bank): contiguous     -> `shard_range`
  * time slabs of one capture (channelizer): frames [f0, f1) plus a halo of
    T = 9 frames (8 for the 9-tap polyphase arms + 1 for the discriminator)     -> `frame_slab`
Only the "one capture, all GPUs" modes move IQ between GPUs: the block is broadcast from the ingest
rank (`broadcast_block`, NCCL over NVLink on GPUs, gloo on CPU for the tests).
"""
from __future__ import annotations

from dataclasses import dataclass

CHAN_HALO_FRAMES = 9  # taps_per_channel - 1 frames of FIR history + 1 frame for the FM discriminator


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Balanced contiguous partition: the first n_items % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class FrameSlab:
    f0: int          # first frame this rank EMITS
    f1: int          # one past the last frame it emits
    start_frame: int # first frame it COMPUTES (f0 - halo, clamped at 0)
    skip: int        # computed frames to drop from the front (= f0 - start_frame)
    sample0: int     # first input sample it reads (start_frame * hop)
    n_samples: int   # input samples it reads

    @property
    def n_frames(self) -> int:
        return self.f1 - self.f0


def frame_slab(n_frames: int, world: int, rank: int, channel_count: int = 256, halo: int = CHAN_HALO_FRAMES) -> FrameSlab:
    """Time shard of one channelizer call with `n_frames` frames (hop = M/2, frame b reads samples
    [b*hop, b*hop + M)). The halo makes the emitted frames independent of where the slab starts, except
    for the first `halo - 1` frames of the call itself, which depend on history carried from the previous call."""
    hop = channel_count // 2
    f0, f1 = shard_range(n_frames, world, rank)
    start = max(0, f0 - halo)
    n = 0 if f1 <= f0 else (f1 - 1 - start) * hop + channel_count
    return FrameSlab(f0, f1, start, f0 - start, start * hop, n)


def replica_seed(base_seed: int, rank: int) -> int:
    return base_seed + rank


def broadcast_block(tensor, src: int = 0, group=None, async_op: bool = False):
    """Broadcast one IQ block from the ingest rank to every rank (NCCL/NVLink on GPU tensors, gloo on CPU)."""
    import torch.distributed as dist

    return dist.broadcast(tensor, src=src, group=group, async_op=async_op)
