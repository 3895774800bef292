"""One-call analog chain: Python face of `wc_analog_plan_*` / `wc_analog_run` (include/wcsdr_b200.h).

`capture.process_channels_batch` builds (and caches) one `AnalogPlan` per distinct set of channel chains and then makes ONE
C call per batch of chunks instead of chaining the stage operators from Python. Filter design stays here, with scipy, exactly
where the reference does it (dsp/fm.py:143, dsp/filters.py:59-61,82, scipy.signal.resample_poly's firwin).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _native as N
from .dsp import _stages as S
from .dsp import agc as AGC

KIND_NONE, KIND_FM, KIND_AM = 0, 1, 2


def _dptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class AnalogPlan:
    """chains: per channel a tuple (kind, iir_stages, agc?, agc_target_db, audio_rate) — `capture._chain_signature` without the
    blanker / noise-reduction fields; kind "fm" | "am" | anything else (metrics only)."""

    def __init__(self, sample_rate: int, chunk_len: int, in_fmt: int, modes, offsets_hz, bfo_hz, squelch_db, chains):
        N.ensure_init()
        self.sample_rate, self.chunk_len, self.in_fmt, self.n_ch = int(sample_rate), int(chunk_len), int(in_fmt), len(modes)
        m = np.ascontiguousarray(modes, dtype=np.int32)
        o = np.ascontiguousarray(offsets_hz, dtype=np.float64)
        b = np.ascontiguousarray(bfo_hz, dtype=np.float64)
        sq = np.ascontiguousarray([np.nan if v is None else float(v) for v in squelch_db], dtype=np.float32)
        h = C.c_void_p()
        lib = N.lib()
        N.check(lib.wc_analog_plan_create(self.sample_rate, self.chunk_len, self.in_fmt, self.n_ch, _dptr(m), _dptr(o), _dptr(b),
                                          _dptr(sq), C.byref(h)))
        self._h = h
        c = 0
        while c < self.n_ch:                        # runs of adjacent channels with identical chains
            e = c + 1
            while e < self.n_ch and chains[e] == chains[c]:
                e += 1
            sig = chains[c]
            kind = KIND_FM if sig[0] == "fm" else KIND_AM if sig[0] == "am" else KIND_NONE
            if kind == KIND_NONE:
                N.check_nonneg(lib.wc_analog_plan_add_run(h, c, e - c, 0, 1, 1, None, 0))
            else:
                audio_rate = int(sig[4])
                if audio_rate == self.sample_rate:
                    run = lib.wc_analog_plan_add_run(h, c, e - c, kind, 1, 1, None, 0)
                else:
                    up, down = S.rate_ratio(self.sample_rate, audio_rate)
                    taps = np.ascontiguousarray(S.resample_taps(up, down), dtype=np.float64)
                    run = lib.wc_analog_plan_add_run(h, c, e - c, kind, up, down, _dptr(taps), taps.size)
                N.check_nonneg(run)
                for bb, aa in sig[1]:
                    bb = np.ascontiguousarray(np.atleast_1d(bb), dtype=np.float64)
                    aa = np.ascontiguousarray(np.atleast_1d(aa), dtype=np.float64)
                    N.check(lib.wc_analog_plan_add_iir(h, run, _dptr(bb), bb.size, _dptr(aa), aa.size))
                if kind == KIND_AM and sig[2]:
                    (ba, aa), (br, ar), target, max_gain = AGC.agc_params(self.sample_rate, float(sig[3]), 5.0, 50.0)
                    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (ba, aa, br, ar)]
                    N.check(lib.wc_analog_plan_set_agc(h, run, _dptr(arrs[0]), _dptr(arrs[1]), _dptr(arrs[2]), _dptr(arrs[3]),
                                                       float(target), float(max_gain)))
            c = e
        N.check(lib.wc_analog_plan_finish(h))
        self.audio_floats = int(lib.wc_analog_plan_audio_floats(h))
        self.audio_len = [int(lib.wc_analog_plan_audio_len(h, i)) for i in range(self.n_ch)]
        self.audio_off = [int(lib.wc_analog_plan_audio_offset(h, i)) for i in range(self.n_ch)]
        self._bufs: dict[int, tuple] = {}           # n_chunks -> (staged input, audio, metrics): fixed addresses, so the
        #                                             C side replays its captured graph for repeated calls

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_analog_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def _buffers(self, n_chunks: int):
        import torch

        if n_chunks not in self._bufs:
            shape = (n_chunks, self.chunk_len, 2) if self.in_fmt == S.FMT_CS16 else (n_chunks, self.chunk_len)
            dt = torch.int16 if self.in_fmt == S.FMT_CS16 else torch.complex64
            self._bufs[n_chunks] = (torch.empty(shape, dtype=dt, device="cuda"),
                                    torch.empty((max(1, n_chunks * self.audio_floats),), dtype=torch.float32, device="cuda"),
                                    torch.empty((3, self.n_ch, n_chunks), dtype=torch.float32, device="cuda"))
        return self._bufs[n_chunks]

    def run(self, x, n_chunks: int):
        """x: CUDA tensor [n_chunks, chunk_len] complex64 / [n_chunks, chunk_len, 2] int16 (used in place), or a numpy
        array of that shape (staged into a plan-owned device buffer). Returns (audio CUDA float32 buffer, metrics CUDA
        float32 [3][n_ch][n_chunks]); both are plan-owned and overwritten by the next run() with the same n_chunks."""
        import torch

        staged, audio, metrics = self._buffers(n_chunks)
        if N.is_torch_cuda(x):
            src = x
        else:
            staged.copy_(torch.from_numpy(np.ascontiguousarray(x)).reshape(staged.shape), non_blocking=False)
            src = staged
        N.check(N.lib().wc_analog_run(self._h, C.c_void_p(src.data_ptr()), int(n_chunks), C.c_void_p(audio.data_ptr()),
                                      C.c_void_p(metrics.data_ptr()), S.stream()))
        return audio, metrics

    def channel_audio(self, audio, channel: int, n_chunks: int):
        """view [n_chunks, audio_len[channel]] of channel's audio inside the run's audio buffer"""
        n_a = self.audio_len[channel]
        off = n_chunks * self.audio_off[channel]
        return audio[off: off + n_chunks * n_a].view(n_chunks, n_a)


# Plans carry per-call state (staging and result buffers, captured graphs) and the reference calls the stateless chain from a
# 3-worker pool (capture.py:1906-1925): one cache per host thread, no sharing.
import threading

_TLS = threading.local()
_MAX_PLANS = 16


def get_plan(sample_rate: int, chunk_len: int, in_fmt: int, modes, offsets_hz, bfo_hz, squelch_db, chains) -> AnalogPlan:
    key = (int(sample_rate), int(chunk_len), int(in_fmt), tuple(modes), tuple(float(v) for v in offsets_hz),
           tuple(float(v) for v in bfo_hz), tuple(None if v is None else float(v) for v in squelch_db), tuple(chains))
    plans = getattr(_TLS, "plans", None)
    if plans is None:
        plans = _TLS.plans = OrderedDict()
    plan = plans.get(key)
    if plan is None:
        plan = AnalogPlan(sample_rate, chunk_len, in_fmt, modes, offsets_hz, bfo_hz, squelch_db, chains)
        plans[key] = plan
        while len(plans) > _MAX_PLANS:
            plans.popitem(last=False)
    else:
        plans.move_to_end(key)
    return plan
