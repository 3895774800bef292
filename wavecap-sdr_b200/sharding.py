"""Host-side partitioning for multi-GPU runs (one process per GPU, torch.distributed for plumbing).

The hot path shards without any data-path collective:
  * independent captures / replicas: rank r owns capture r                      -> `replica_seed`
  * channels of one capture (analog chain, P25 banks, DDC bank): contiguous     -> `shard_range`
  * time slabs of one capture (channelizer): frames [f0, f1) plus a halo of
    T = 9 frames (8 for the 9-tap polyphase arms + 1 for the discriminator)     -> `frame_slab`
Only the "one capture, all GPUs" modes move IQ between GPUs. Two transports:
  * `broadcast_block`: the block is NCCL-broadcast from the ingest rank (gloo on CPU for the tests) — every rank
    receives all of it;
  * `PeerRegion`: the ingest rank's buffer is mapped into every rank (CUDA IPC over NVLink, `wc_peer_*`) and each
    rank's kernels pull ONLY their slab out of it while they compute; `slab_weights` gives the ingest rank the larger
    share its local HBM affords, flags in the region (`wc_flag_*`) order producer and consumers on the streams.
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


def weighted_range(n_items: int, weights, rank: int) -> tuple[int, int]:
    """Contiguous partition with shares proportional to `weights` (boundaries = rounded cumulative shares)."""
    w = [float(x) for x in weights]
    if not w or not (0 <= rank < len(w)) or min(w) < 0 or sum(w) <= 0:
        raise ValueError(f"bad rank {rank} / weights {weights}")
    tot = sum(w)
    edges = [0]
    acc = 0.0
    for x in w:
        acc += x
        edges.append(max(edges[-1], min(n_items, int(round(n_items * acc / tot)))))
    edges[-1] = n_items
    return edges[rank], edges[rank + 1]


def slab_weights(world: int, local_rate: float, link_rate: float, src: int = 0) -> list[float]:
    """Shares that finish together when the ingest rank `src` works from its own HBM at `local_rate` and the other
    ranks together can pull at most `link_rate` (its NVLink egress) — same unit, e.g. GS/s: src gets
    local/(local+link), the rest split the remainder evenly."""
    if world == 1:
        return [1.0]
    a = local_rate / (local_rate + link_rate)
    return [a if r == src else (1.0 - a) / (world - 1) for r in range(world)]


def frame_slab(n_frames: int, world: int, rank: int, channel_count: int = 256, halo: int = CHAN_HALO_FRAMES,
               weights=None) -> FrameSlab:
    """Time shard of one channelizer call with `n_frames` frames (hop = M/2, frame b reads samples
    [b*hop, b*hop + M)). The halo makes the emitted frames independent of where the slab starts, except
    for the first `halo - 1` frames of the call itself, which depend on history carried from the previous call.
    `weights` (one per rank) makes the shares unequal (`slab_weights`)."""
    hop = channel_count // 2
    if weights is None:
        f0, f1 = shard_range(n_frames, world, rank)
    else:
        if len(weights) != world:
            raise ValueError("one weight per rank")
        f0, f1 = weighted_range(n_frames, weights, rank)
    start = max(0, f0 - halo)
    n = 0 if f1 <= f0 else (f1 - 1 - start) * hop + channel_count
    return FrameSlab(f0, f1, start, f0 - start, start * hop, n)


def replica_seed(base_seed: int, rank: int) -> int:
    return base_seed + rank


def broadcast_block(tensor, src: int = 0, group=None, async_op: bool = False):
    """Broadcast one IQ block from the ingest rank to every rank (NCCL/NVLink on GPU tensors, gloo on CPU)."""
    import torch.distributed as dist

    return dist.broadcast(tensor, src=src, group=group, async_op=async_op)


@dataclass(frozen=True)
class DeviceSpan:
    """`count` complex64 samples at device address `ptr` — possibly another GPU's memory mapped here (PeerRegion).
    Quacks like the part of a CUDA tensor the slab entry points use."""
    ptr: int
    count: int

    def data_ptr(self) -> int:
        return self.ptr

    def numel(self) -> int:
        return self.count


class PeerRegion:
    """`nbytes` of device memory on rank `src`, mapped into every rank of the process group (wc_peer_alloc /
    wc_peer_open: CUDA IPC, NVLink P2P). The first FLAG_BYTES hold uint32 sequence flags (zeroed), the rest is payload.
    Collective: every rank constructs it at the same point."""

    FLAG_BYTES = 4096

    def __init__(self, payload_bytes: int, src: int = 0, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as N

        self.src = src
        self.payload_bytes = int(payload_bytes)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.is_owner = self.rank == src
        N.ensure_init()
        total = self.FLAG_BYTES + self.payload_bytes
        ptr = C.c_void_p()
        box = [None]
        if self.is_owner:
            handle = (C.c_ubyte * 64)()
            N.check(N.lib().wc_peer_alloc(total, C.byref(ptr), handle))
            box[0] = bytes(handle)
            self.base = int(ptr.value)
            self.flags_tensor().zero_()
            torch.cuda.synchronize()
        if self.world > 1:
            dist.broadcast_object_list(box, src=src, group=group)
        self.handle = box[0]
        if not self.is_owner:
            N.check(N.lib().wc_peer_open(C.c_char_p(self.handle), C.byref(ptr)))
            self.base = int(ptr.value)
        self._open = True

    @classmethod
    def open(cls, handle: bytes, payload_bytes: int) -> "PeerRegion":
        """Map a region exported by another process (its `.handle`), outside any process group."""
        import ctypes as C

        from . import _native as N

        self = cls.__new__(cls)
        self.src, self.rank, self.world, self.is_owner = -1, -1, 0, False
        self.payload_bytes, self.handle = int(payload_bytes), handle
        N.ensure_init()
        ptr = C.c_void_p()
        N.check(N.lib().wc_peer_open(C.c_char_p(handle), C.byref(ptr)))
        self.base = int(ptr.value)
        self._open = True
        return self

    # -- addressing ---------------------------------------------------------------------------------
    def flag_ptr(self, index: int) -> int:
        assert 0 <= index < self.FLAG_BYTES // 4
        return self.base + 4 * index

    def payload_ptr(self, byte_offset: int = 0) -> int:
        assert 0 <= byte_offset <= self.payload_bytes
        return self.base + self.FLAG_BYTES + byte_offset

    def span(self, sample_offset: int, count: int) -> DeviceSpan:
        assert 8 * (sample_offset + count) <= self.payload_bytes
        return DeviceSpan(self.payload_ptr(8 * sample_offset), count)

    def _owner_tensor(self, ptr: int, shape, typestr: str):
        import torch

        assert self.is_owner, "only the owning rank may wrap the region as a tensor"

        class _Arr:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Arr(), device="cuda")

    def flags_tensor(self):
        return self._owner_tensor(self.base, (self.FLAG_BYTES // 4,), "<i4")

    def payload_tensor(self, sample_offset: int, count: int):
        """complex64 view of the payload (owner only; other ranks address it through `span`)."""
        import torch

        return torch.view_as_complex(self._owner_tensor(self.payload_ptr(8 * sample_offset), (count, 2), "<f4"))

    # -- stream-ordered flags -------------------------------------------------------------------------
    def set_flag(self, index: int, value: int) -> None:
        import ctypes as C

        from . import _native as N

        N.check(N.lib().wc_flag_set(C.c_void_p(self.flag_ptr(index)), value & 0xFFFFFFFF, N.torch_stream_ptr()))

    def wait_flags(self, index: int, count: int, value: int, timeout_ms: int = 2000, timed_out=None, stride: int = 1) -> None:
        """Stream waits until flags index, index+stride, ... (count of them) reach `value`. A wait that gives up (dead or
        stalled peer) sets `timed_out` (int32 CUDA tensor; default: a flag this region owns) instead of hanging the GPU —
        whatever the stream computes after that read a buffer that was never published, so the caller MUST look at the
        flag before using any result: `check()` raises, and results produced since the last clean check are void."""
        import ctypes as C

        from . import _native as N

        if timed_out is None:
            timed_out = self._own_timeout_flag()
        N.check(N.lib().wc_flag_wait(C.c_void_p(self.flag_ptr(index)), count, stride, value & 0xFFFFFFFF, timeout_ms,
                                     C.c_void_p(timed_out.data_ptr()), N.torch_stream_ptr()))

    def _own_timeout_flag(self):
        import torch

        if getattr(self, "_timed_out", None) is None:
            self._timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        return self._timed_out

    def check(self) -> None:
        """Synchronise the current stream and raise if any wait_flags() since the last check timed out."""
        import torch

        torch.cuda.current_stream().synchronize()
        t = getattr(self, "_timed_out", None)
        if t is not None and int(t.item()) != 0:
            t.zero_()
            raise TimeoutError("PeerRegion: a flag wait timed out — the peer never published the block; "
                               "discard every result computed since the last check()")

    def close(self) -> None:
        import ctypes as C

        from . import _native as N

        if not getattr(self, "_open", False):
            return
        self._open = False
        if self.is_owner:
            N.check(N.lib().wc_peer_free(C.c_void_p(self.base)))
        else:
            N.check(N.lib().wc_peer_close(C.c_void_p(self.base)))
