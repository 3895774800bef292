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


@dataclass(frozen=True)
class StripeLayout:
    """Where the samples of one block (one process() call of `frames` frames, hop 128, frame 256) live when the block is
    striped over `world` ranks, and what each rank fetches from its neighbour (all offsets in samples of the block)."""
    frames: int
    f0: tuple        # first frame each rank emits
    f1: tuple        # one past the last
    own0: tuple      # first sample of each rank's own part: 128 * f0
    own_n: tuple     # samples of the own part: its frames' rows, (f1 - 1 - f0) * 128 + 256
    halo: int        # samples a rank > 0 needs in front of its own part (from rank r - 1): 9 hop rows
    tail0: int       # first sample of the (T + 1)-row tail of the block that rank 0 installs as carried history next block
    tail_n: int

    def halo_src(self, r: int) -> int:
        """offset of rank r's halo inside rank r - 1's own part"""
        return self.own0[r] - self.halo - self.own0[r - 1]

    def tail_src(self) -> int:
        """offset of the block's tail inside the last rank's own part"""
        return self.tail0 - self.own0[-1]


def stripe_layout(frames: int, world: int) -> StripeLayout:
    slabs = [frame_slab(frames, world, r, halo=0) for r in range(world)]
    if any(s.n_frames <= CHAN_HALO_FRAMES + 1 for s in slabs):
        raise ValueError("block too short for this many ranks")
    return StripeLayout(frames, tuple(s.f0 for s in slabs), tuple(s.f1 for s in slabs), tuple(128 * s.f0 for s in slabs),
                        tuple((s.f1 - 1 - s.f0) * 128 + 256 for s in slabs), CHAN_HALO_FRAMES * 128,
                        (frames - CHAN_HALO_FRAMES) * 128, (CHAN_HALO_FRAMES + 1) * 128)


class StripedCapture:
    """ONE capture over all ranks, striped at ingest: every block (= one `PolyphaseChannelizer.process()` call of
    `block_samples` samples) is cut into `world` time slabs and slab r is delivered straight into rank r's memory (its own
    PCIe link / DMA target), so no rank's NVLink egress or HBM carries the whole capture. What a rank lacks is only the
    9 hop rows in front of its slab (8 rows of polyphase history + 1 for the FM discriminator): 1152 samples = 9 KB per
    block, copied from the previous slab's owner over NVLink (peer-mapped regions, stream-ordered flags); rank 0 takes
    the 10 tail rows of the previous block from the last rank instead and installs them as the carried history
    (`wc_chan_carry_tail`). Emitted frames are bit-equal to the unsharded call's.

    Flags in rank r's region (owner r): READY[b] = block number whose slab is in buffer b; DONE[b] = block number whose
    tail the reader (rank (r + 1) % world) has finished copying — the owner may overwrite buffer b after that.
    Collective: every rank constructs it at the same point."""

    HALO_ROWS = CHAN_HALO_FRAMES          # rows a rank > 0 needs in front of its slab
    AREA = (CHAN_HALO_FRAMES + 1) * 128   # samples reserved in front of the own part (rank 0's tail fetch needs T + 1 rows)
    READY, DONE = 0, 16

    def __init__(self, chan, block_samples: int, group=None):
        import torch.distributed as dist

        assert chan.channel_count == 256 and chan.taps_per_channel == 9
        self.chan = chan
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.block_samples = int(block_samples)
        self.frames = chan.frames_for(self.block_samples)
        self.slabs = [frame_slab(self.frames, self.world, r, halo=0) for r in range(self.world)]
        self.layout = stripe_layout(self.frames, self.world)
        # own part of rank r: the samples its frames [f0, f1) cover = [128 f0, 128 (f1 - 1) + 256)
        self.own0 = list(self.layout.own0)
        self.own_n = list(self.layout.own_n)
        self.stride = self.AREA + ((max(self.own_n) + 15) & ~15)      # samples per buffer
        self.regions = [PeerRegion(8 * 2 * self.stride, src=r, group=group) for r in range(self.world)]
        self.mine = self.regions[self.rank]
        self.block = 0                        # blocks processed so far

    # -- addressing (sample offsets inside a region's payload) ----------------------------------------
    def _own_off(self, b: int) -> int:
        return b * self.stride + self.AREA

    def own_tensor(self, b: int):
        """complex64 view of this rank's slab in buffer b: where the ingest writes."""
        return self.mine.payload_tensor(self._own_off(b), self.own_n[self.rank])

    def slab_of(self, block_tensor_or_array, r: int | None = None):
        """the samples of a whole block that belong to rank r's slab (for the ingest side / the tests)"""
        r = self.rank if r is None else r
        return block_tensor_or_array[self.own0[r]: self.own0[r] + self.own_n[r]]

    # -- one block --------------------------------------------------------------------------------------
    def publish(self, timeout_ms: int = 5000) -> int:
        """Call after the ingest of the next block's slab into buffer (block & 1) has been enqueued on the current
        stream: marks it READY. Returns the buffer index. (Before overwriting a buffer the ingest side calls
        `wait_free` first.)"""
        b = self.block & 1
        self.mine.set_flag(self.READY + b, self.block + 1)
        return b

    def wait_free(self, timeout_ms: int = 5000) -> int:
        """stream waits until the reader of this rank's buffer (block & 1) has copied the tail of block - 2"""
        b = self.block & 1
        if self.block >= 2 and self.world > 1:
            self.mine.wait_flags(self.DONE + b, 1, self.block - 1, timeout_ms=timeout_ms)
        return b

    def process(self, fm: bool = True, timeout_ms: int = 5000):
        """Channelize (+ FM-demodulate) this rank's slab of the current block; returns (rows [n_frames, 256], f0)."""
        import ctypes as C

        import torch

        from . import _native as N
        from .dsp.channelizer import OUT_COMPLEX, OUT_FM, fm_scale

        i, r, W = self.block, self.rank, self.world
        b, seq = i & 1, i + 1
        lib, st = N.lib(), N.torch_stream_ptr()
        s = self.slabs[r]
        own_ptr = self.mine.payload_ptr(8 * self._own_off(b))
        mode = OUT_FM if fm else OUT_COMPLEX
        scale = fm_scale(int(self.chan.channel_sample_rate)) if fm else 0.0
        if r == 0:
            # history = the T + 1 tail rows of the previous block, owned by the last rank (previous buffer)
            if i > 0:
                last = self.regions[W - 1]
                pb = (i - 1) & 1
                if W > 1:
                    last.wait_flags(self.READY + pb, 1, seq - 1, timeout_ms=timeout_ms)
                tail_off = pb * self.stride + self.AREA + self.layout.tail_src()
                dst = self.mine.payload_ptr(8 * (b * self.stride))          # the area in front of my own part
                N.check(lib.wc_peer_copy(C.c_void_p(dst), C.c_void_p(last.payload_ptr(8 * tail_off)), 8 * self.AREA, st))
                if W > 1:
                    last.set_flag(self.DONE + pb, seq - 1)
                N.check(lib.wc_chan_carry_tail(self.chan._h, C.c_void_p(dst), st))
            else:
                self.chan.reset()
            rows = torch.empty((s.n_frames, 256), device="cuda", dtype=torch.float32 if fm else torch.complex64)
            N.check(lib.wc_chan_process(self.chan._h, C.c_void_p(own_ptr), self.own_n[0], 1, self.own_n[0], mode, scale,
                                        C.c_void_p(rows.data_ptr()), st))
            out = rows
        else:
            prev = self.regions[r - 1]
            prev.wait_flags(self.READY + b, 1, seq, timeout_ms=timeout_ms)
            halo = self.HALO_ROWS * 128
            src_off = b * self.stride + self.AREA + self.layout.halo_src(r)
            dst = own_ptr - 8 * halo
            N.check(lib.wc_peer_copy(C.c_void_p(dst), C.c_void_p(prev.payload_ptr(8 * src_off)), 8 * halo, st))
            prev.set_flag(self.DONE + b, seq)
            n_local = halo + self.own_n[r]
            rows = torch.empty((s.n_frames + self.HALO_ROWS, 256), device="cuda", dtype=torch.float32 if fm else torch.complex64)
            N.check(lib.wc_chan_process(self.chan._h, C.c_void_p(dst), n_local, 1, n_local, mode, scale,
                                        C.c_void_p(rows.data_ptr()), st))
            out = rows[self.HALO_ROWS:]       # the first 9 local frames only rebuilt the FIR / discriminator state
        self.block += 1
        return out, s.f0

    def check(self) -> None:
        for reg in self.regions:
            reg.check()

    def close(self) -> None:
        for reg in self.regions:
            reg.close()
