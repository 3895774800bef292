"""GPU polyphase channelizer with the call surface of `wavecapsdr.dsp.channelizer`.

Mirrors (names, argument meaning, return contracts) wavecapsdr/dsp/channelizer.py:
`PolyphaseChannelizer` (:28-158), `ChannelCalculator` (:161-231), `channelize_samples` (:234-268).
All arithmetic runs in csrc/channelizer.cu through the C ABI (`wc_chan_*`).

Extra, GPU-only entry points (not in the reference): `process_array`, `process_fm`,
`process_batch` and device-tensor inputs.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .. import _native as N

DEFAULT_CHANNEL_BANDWIDTH = 25000
DEFAULT_TAPS_PER_CHANNEL = 9
PERFECT_RECONSTRUCTION_GAIN = 0.5

OUT_COMPLEX = 0
OUT_FM = 1
OUT_AUDIO = 2
IN_CF32 = 0
IN_CS16 = 1


def fm_scale(sample_rate: int) -> float:
    """Discriminator scale of dsp/fm.py:94: float32(fs / (2*pi*75000))."""
    return float(np.float32(sample_rate / (2.0 * np.pi * 75000.0)))


class PolyphaseChannelizer:
    """2x-oversampled polyphase filter bank; state carries across `process()` calls."""

    def __init__(self, sample_rate: float, channel_bandwidth: int = DEFAULT_CHANNEL_BANDWIDTH,
                 taps_per_channel: int = DEFAULT_TAPS_PER_CHANNEL):
        N.ensure_init()
        self.sample_rate = sample_rate
        self.channel_bandwidth = channel_bandwidth
        self.taps_per_channel = taps_per_channel
        h = C.c_void_p()
        N.check(N.lib().wc_chan_create(float(sample_rate), int(channel_bandwidth), int(taps_per_channel),
                                       C.byref(h)))
        self._h = h
        m, rate, t = C.c_int(), C.c_double(), C.c_int()
        N.check(N.lib().wc_chan_info(self._h, C.byref(m), C.byref(rate), C.byref(t)))
        self.channel_count = m.value
        self.channel_sample_rate = rate.value
        self.block_counter = 0  # unused by the reference as well (channelizer.py:67)
        arms = np.empty((self.channel_count, self.taps_per_channel), dtype=np.float64)
        N.check(N.lib().wc_chan_get_arms(self._h, N.np_ptr(arms)))
        self.arms = arms

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_chan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- reference attribute ---------------------------------------------------------------------
    @property
    def arm_history(self) -> np.ndarray:
        """complex64 [channel_count, taps_per_channel]; column j = block fed j frames ago."""
        out = np.empty((self.channel_count, self.taps_per_channel), dtype=np.complex64)
        N.check(N.lib().wc_chan_get_history(self._h, N.np_ptr(out)))
        return out

    def frames_for(self, n_samples: int) -> int:
        return int(N.lib().wc_chan_frames_for(self._h, int(n_samples)))

    # -- core ------------------------------------------------------------------------------------
    def _run(self, samples, mode: int, n_chunks: int = 1, scale: float = 0.0):
        """numpy in -> numpy out, CUDA tensor in -> CUDA tensor out. complex input is complex64 IQ; int16 input
        ([..., 2] or flat interleaved I,Q) is the raw capture format, scaled by 1/32768 on the device (cli.py:449-453)."""
        m = self.channel_count
        out_dtype = np.complex64 if mode == OUT_COMPLEX else np.float32
        if N.is_torch_cuda(samples):
            import torch

            x = samples
            cs16 = x.dtype == torch.int16
            if not cs16 and x.dtype != torch.complex64:
                x = x.to(torch.complex64)
            x = x.contiguous()
            n = (x.numel() // 2 if cs16 else x.numel()) // n_chunks
            rows = self._rows_for(n, mode)
            out = torch.empty((rows * n_chunks, m), device=x.device,
                              dtype=torch.complex64 if mode == OUT_COMPLEX else torch.float32)
            if rows:
                N.check(N.lib().wc_chan_process_ex(self._h, C.c_void_p(x.data_ptr()), IN_CS16 if cs16 else IN_CF32, n, n_chunks, n,
                                                   mode, scale, C.c_void_p(out.data_ptr()), N.torch_stream_ptr()))
            return out
        a = np.asarray(samples)
        cs16 = a.dtype == np.int16
        x = np.ascontiguousarray(a, dtype=np.int16 if cs16 else np.complex64)
        n = (x.size // 2 if cs16 else x.size) // n_chunks
        rows = self._rows_for(n, mode)
        out = np.empty((rows * n_chunks, m), dtype=out_dtype)
        if rows:
            N.check(N.lib().wc_chan_process_host_ex(self._h, N.np_ptr(x), IN_CS16 if cs16 else IN_CF32, n, n_chunks, mode, scale,
                                                    N.np_ptr(out)))
        return out

    def _rows_for(self, n_samples: int, mode: int) -> int:
        if mode == OUT_AUDIO:
            return int(N.lib().wc_chan_audio_len(self._h, int(n_samples)))
        return self.frames_for(n_samples)

    def process_audio(self, samples, demod_sample_rate: int | None = None, audio_rate: int | None = None, n_chunks: int = 1):
        """`nbfm_demod(extract_channel(process(samples), k), demod_sample_rate, audio_rate)` (dsp/fm.py:317-406, defaults) for
        every channel k in one pass: float32 [n_audio, channel_count] per chunk, n_audio = ceil(frames / D). Built for
        integer decimation D = demod_sample_rate / audio_rate; defaults: the channel rate rounded down to a multiple of
        20 and a twentieth of it (976560 -> 48828 for the 125 MS/s / 256-channel grid)."""
        if demod_sample_rate is None:
            demod_sample_rate = (int(self.channel_sample_rate) // 20) * 20
        if audio_rate is None:
            audio_rate = demod_sample_rate // 20
        key = (int(demod_sample_rate), int(audio_rate))
        if getattr(self, "_audio_key", None) != key:
            N.check(N.lib().wc_chan_audio_config(self._h, key[0], key[1]))
            self._audio_key = key
        return self._run(samples, OUT_AUDIO, n_chunks=n_chunks)

    def process(self, samples) -> list:
        """One array of `channel_count` complex64 values per output frame (channelizer.py:91-137)."""
        return list(self.process_array(samples))

    def process_array(self, samples):
        """Same result as `process()` as one [frames, channel_count] array (numpy in -> numpy out,
        CUDA tensor in -> CUDA tensor out)."""
        return self._run(samples, OUT_COMPLEX)

    def process_fm(self, samples, demod_sample_rate: int | None = None):
        """Fused `quadrature_demod(extract_channel(process(samples), k), rate)` for every k:
        float32 [frames, channel_count]; row 0 is zero (dsp/fm.py:90-91)."""
        rate = int(self.channel_sample_rate) if demod_sample_rate is None else int(demod_sample_rate)
        return self._run(samples, OUT_FM, scale=fm_scale(rate))

    def process_batch(self, samples, n_chunks: int, fm: bool = False, demod_sample_rate: int | None = None):
        """`n_chunks` back-to-back `process()` calls of equal length in one launch."""
        rate = int(self.channel_sample_rate) if demod_sample_rate is None else int(demod_sample_rate)
        return self._run(samples, OUT_FM if fm else OUT_COMPLEX, n_chunks=n_chunks,
                         scale=fm_scale(rate) if fm else 0.0)

    def process_slab(self, block, world: int, rank: int, fm: bool = False, demod_sample_rate: int | None = None,
                     weights=None):
        """Time shard of ONE `process(block)` call for multi-GPU runs: this rank emits frames [f0, f1) of the call
        (`sharding.frame_slab`, shares proportional to `weights` when given), computing a 9-frame halo in front so the
        rows equal the unsharded call's. `block` is the whole call's input: a CUDA tensor (every rank holds it after
        the NCCL broadcast) or a `sharding.DeviceSpan` into the ingest rank's memory mapped here (`PeerRegion`) — the
        kernel then pulls just this slab over NVLink while it computes. The carried history is then advanced from the
        true end of the block, as if the whole call had run here.
        Returns (rows [f1 - f0, channel_count], f0)."""
        import torch

        from ..sharding import DeviceSpan, frame_slab

        if isinstance(block, DeviceSpan):
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            assert N.is_torch_cuda(block) and block.dtype == torch.complex64 and block.is_contiguous()
            device = block.device
        n = int(block.numel())
        m = self.channel_count
        s = frame_slab(self.frames_for(n), world, rank, channel_count=m, halo=self.taps_per_channel, weights=weights)
        mode = OUT_FM if fm else OUT_COMPLEX
        rate = int(self.channel_sample_rate) if demod_sample_rate is None else int(demod_sample_rate)
        rows = s.f1 - s.start_frame
        out = torch.empty((max(rows, 0), m), device=device, dtype=torch.float32 if fm else torch.complex64)
        if s.n_frames > 0:
            peer = isinstance(block, DeviceSpan)
            if peer:
                # halo rows of every CTA run cross NVLink again (8 per run: 8 % at 96 frames): one wave of long runs, not six of short ones
                slots = 4 * torch.cuda.get_device_properties(device).multi_processor_count
                N.check(N.lib().wc_chan_set_run_frames(self._h, min(256, max(96, -(-rows // slots)))))
            try:
                N.check(N.lib().wc_chan_process(self._h, C.c_void_p(block.data_ptr() + 8 * s.sample0), s.n_samples, 1,
                                                s.n_samples, mode, fm_scale(rate) if fm else 0.0,
                                                C.c_void_p(out.data_ptr()), N.torch_stream_ptr()))
            finally:
                if peer:
                    N.check(N.lib().wc_chan_set_run_frames(self._h, 0))
        N.check(N.lib().wc_chan_carry_from(self._h, C.c_void_p(block.data_ptr()), n, N.torch_stream_ptr()))
        return out[s.skip:], s.f0

    def reset(self) -> None:
        N.check(N.lib().wc_chan_reset(self._h))
        self.block_counter = 0

    def extract_channel(self, channel_results, channel_index: int):
        """Samples of one channel across frames (channelizer.py:144-158)."""
        if isinstance(channel_results, np.ndarray):
            return np.ascontiguousarray(channel_results[:, channel_index], dtype=np.complex64)
        if N.is_torch_cuda(channel_results):
            return channel_results[:, channel_index].contiguous()
        return np.array([fr[channel_index] for fr in channel_results], dtype=np.complex64)


class ChannelCalculator:
    """Frequency <-> FFT-bin bookkeeping (channelizer.py:161-231); pure host arithmetic."""

    def __init__(self, center_frequency: float, sample_rate: float,
                 channel_bandwidth: int = DEFAULT_CHANNEL_BANDWIDTH):
        self.center_frequency = center_frequency
        self.sample_rate = sample_rate
        self.channel_bandwidth = channel_bandwidth
        count = int(sample_rate / channel_bandwidth)
        self.channel_count = count - (count % 2)

    def get_channel_index(self, target_frequency: float) -> int:
        steps = int(round((target_frequency - self.center_frequency) / self.channel_bandwidth))
        return self.channel_count + steps if steps < 0 else steps % self.channel_count

    def get_channel_center_frequency(self, channel_index: int) -> float:
        signed = channel_index if channel_index < self.channel_count // 2 else channel_index - self.channel_count
        return self.center_frequency + signed * self.channel_bandwidth


def channelize_samples(samples, sample_rate: float, target_frequency: float, center_frequency: float,
                       channel_bandwidth: int = DEFAULT_CHANNEL_BANDWIDTH):
    """Extract one channel from wideband IQ (channelizer.py:234-268)."""
    chan = PolyphaseChannelizer(sample_rate, channel_bandwidth)
    idx = ChannelCalculator(center_frequency, sample_rate, channel_bandwidth).get_channel_index(target_frequency)
    frames = chan.process_array(samples)
    return chan.extract_channel(frames, idx), chan.channel_sample_rate
