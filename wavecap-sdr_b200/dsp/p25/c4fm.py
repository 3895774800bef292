"""GPU C4FM demodulator with the call surface of `wavecapsdr.dsp.p25.c4fm`.

Mirrors wavecapsdr/dsp/p25/c4fm.py: `C4FMDemodulator` (:2379-2807: `demodulate`, `reset`,
`get_timing_offset`), `design_baseband_lpf` (:95-132), `design_rrc_filter` (:135-183). All signal
arithmetic runs in csrc/p25.cu through the C ABI (`wc_c4fm_*`); filters are designed on the host at
construction, like the reference does (scipy at plan creation, never on the data path).

GPU-only addition: `C4FMBank` — C independent stateful demodulators advanced by one call (what the
reference does with one Python object per control/voice channel).
"""
from __future__ import annotations

import ctypes as C

import os

import numpy as np

from ... import _native as N

NUMBA_AVAILABLE = False  # name kept for benchmark_dsp.py:19; nothing here is JIT-compiled Python

EQUALIZER_LOOP_GAIN = 0.15
MAXIMUM_PLL = np.pi / 3.0
MAXIMUM_GAIN = 1.25
INITIAL_GAIN = 1.219


def design_baseband_lpf(sample_rate: float, passband_hz: float = 5200.0, stopband_hz: float = 6500.0,
                        num_taps: int = 63) -> np.ndarray:
    """c4fm.py:95-132 — remez with the legacy `Hz=` keyword inside try/except, windowed-sinc fallback
    (with scipy >= 1.15 the fallback is what runs, in the reference as here)."""
    from scipy import signal

    try:
        h = signal.remez(num_taps, [0, passband_hz, stopband_hz, sample_rate / 2.0], [1, 0], Hz=sample_rate)
    except Exception:
        h = signal.firwin(num_taps, passband_hz, fs=sample_rate, window="hamming")
    return np.asarray(h, dtype=np.float32)


def design_rrc_filter(samples_per_symbol: float, num_taps: int = 101, alpha: float = 0.2) -> np.ndarray:
    """c4fm.py:135-183 — root raised cosine, normalised to unit sum, float32."""
    if num_taps % 2 == 0:
        num_taps += 1
    t = (np.arange(num_taps) - (num_taps - 1) / 2) / samples_per_symbol
    h = np.zeros(num_taps, dtype=np.float64)
    for i, ti in enumerate(t):
        if ti == 0:
            h[i] = 1 - alpha + 4 * alpha / np.pi
        elif abs(ti) == 1 / (4 * alpha):
            h[i] = (alpha / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha))
                                           + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))) / (
                np.pi * ti * (1 - (4 * alpha * ti) ** 2))
    return (h / np.sum(h)).astype(np.float32)


class C4FMBank:
    """`n_channels` C4FM demodulators with persistent per-channel state, one launch sequence per call."""

    def __init__(self, n_channels: int, sample_rate: int = 19200, symbol_rate: int = 4800, wide_pulse: bool = False):
        N.ensure_init()
        self.n_channels = int(n_channels)
        self.sample_rate = sample_rate
        self.symbol_rate = symbol_rate
        self.samples_per_symbol = sample_rate / symbol_rate
        self.wide_pulse = wide_pulse
        pb, sb, alpha = (10000.0, 12000.0, 0.5) if wide_pulse else (5200.0, 6500.0, 0.2)
        self._baseband_lpf = design_baseband_lpf(sample_rate, passband_hz=pb, stopband_hz=sb)
        self._rrc_filter = design_rrc_filter(self.samples_per_symbol,
                                             num_taps=int(16 * self.samples_per_symbol) + 1, alpha=alpha)
        h = C.c_void_p()
        N.check(N.lib().wc_c4fm_create(self.n_channels, int(sample_rate), int(symbol_rate), int(bool(wide_pulse)),
                                       N.np_ptr(self._baseband_lpf), int(self._baseband_lpf.size),
                                       N.np_ptr(self._rrc_filter), int(self._rrc_filter.size), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_c4fm_destroy(h)
            except Exception:
                pass

    def max_symbols(self, n_samples: int) -> int:
        return int(N.lib().wc_c4fm_max_symbols(self._h, int(n_samples)))

    def reset(self, channel: int = -1) -> None:
        N.check(N.lib().wc_c4fm_reset(self._h, int(channel)))

    def state(self, channel: int = 0) -> dict:
        s = np.zeros(8, dtype=np.float64)
        N.check(N.lib().wc_c4fm_get_state(self._h, int(channel), N.np_ptr(s)))
        return {"pll": s[0], "gain": s[1], "sample_point": s[2], "buffer_pointer": int(s[3]), "fine_sync": bool(s[4]),
                "symbols_since_sync": int(s[5]), "sync_count": int(s[6]), "events_last_call": int(s[7])}

    def demodulate(self, iq):
        """iq: complex64 [n_channels][n] (numpy, or a torch CUDA tensor) ->
        (dibits uint8 [C][max_sym], soft float32 [C][max_sym], counts int32 [C]); row c is valid up to counts[c]."""
        if N.is_torch_cuda(iq):
            import torch

            x = iq.to(torch.complex64).contiguous()
            assert x.dim() == 2 and x.shape[0] == self.n_channels, x.shape
            n = int(x.shape[1])
            ms = max(1, self.max_symbols(n))
            dib = torch.zeros((self.n_channels, ms), dtype=torch.uint8, device=x.device)
            soft = torch.zeros((self.n_channels, ms), dtype=torch.float32, device=x.device)
            cnt = torch.zeros((self.n_channels,), dtype=torch.int32, device=x.device)
            if n:
                N.check(N.lib().wc_c4fm_demod(self._h, C.c_void_p(x.data_ptr()), n, n, C.c_void_p(dib.data_ptr()),
                                              C.c_void_p(soft.data_ptr()), C.c_void_p(cnt.data_ptr()), ms,
                                              N.torch_stream_ptr()))
            return dib, soft, cnt
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        assert x.ndim == 2 and x.shape[0] == self.n_channels, x.shape
        n = int(x.shape[1])
        ms = max(1, self.max_symbols(n))
        dib = np.zeros((self.n_channels, ms), dtype=np.uint8)
        soft = np.zeros((self.n_channels, ms), dtype=np.float32)
        cnt = np.zeros((self.n_channels,), dtype=np.int32)
        if n:
            N.check(N.lib().wc_c4fm_demod_host(self._h, N.np_ptr(x), n, N.np_ptr(dib), N.np_ptr(soft), N.np_ptr(cnt), ms))
        return dib, soft, cnt


    def demodulate_discriminator(self, audio):
        """audio: [n_channels][n] discriminator audio (any float dtype; cast to float32 like the reference, audio[:, 0]
        kept in float64 for the first-call filter state) -> (dibits, soft, counts) as `demodulate`."""
        a = np.asarray(audio)
        assert a.ndim == 2 and a.shape[0] == self.n_channels, a.shape
        n = int(a.shape[1])
        ms = max(1, self.max_symbols(n))
        dib = np.zeros((self.n_channels, ms), dtype=np.uint8)
        soft = np.zeros((self.n_channels, ms), dtype=np.float32)
        cnt = np.zeros((self.n_channels,), dtype=np.int32)
        if n:
            first = np.ascontiguousarray(a[:, 0].astype(np.float64))
            x = np.ascontiguousarray(a.astype(np.float32))
            N.check(N.lib().wc_c4fm_demod_disc_host(self._h, N.np_ptr(x), n, N.np_ptr(first), N.np_ptr(dib), N.np_ptr(soft),
                                                    N.np_ptr(cnt), ms))
        return dib, soft, cnt


class _EqualizerView:
    """`demod._equalizer.pll` / `.gain` as the reference's CLI prints them (cli.py:759-760): read-only view of the
    channel's equaliser state on the device (c4fm.py:199-272)."""

    def __init__(self, bank, channel: int = 0):
        self._bank, self._channel = bank, channel

    @property
    def pll(self) -> float:
        return float(self._bank.state(self._channel)["pll"])

    @property
    def gain(self) -> float:
        return float(self._bank.state(self._channel)["gain"])


class C4FMDemodulator:
    """Drop-in for wavecapsdr.dsp.p25.c4fm.C4FMDemodulator (one channel)."""

    # c4fm.py:2408-2410 (the kernels use the first two: csrc/p25.cu C4_THRESH)
    SYNC_THRESHOLD_DETECTION = 100.0
    SYNC_THRESHOLD_OPTIMIZED = 100.0
    SYNC_THRESHOLD_EQUALIZED = 179.0

    def __init__(self, sample_rate: int = 19200, symbol_rate: int = 4800, wide_pulse: bool = False, **kwargs):
        self._bank = C4FMBank(1, sample_rate, symbol_rate, wide_pulse)
        self.sample_rate = sample_rate
        self.symbol_rate = symbol_rate
        self.samples_per_symbol = sample_rate / symbol_rate
        self.wide_pulse = wide_pulse
        self._baseband_lpf = self._bank._baseband_lpf
        self._rrc_filter = self._bank._rrc_filter

    def reset(self) -> None:
        self._bank.reset(0)

    def demodulate(self, iq):
        """c4fm.py:2528-2807: complex IQ -> (dibits uint8, soft symbols float32)."""
        iq = np.asarray(iq)
        if iq.size == 0:
            return np.array([], dtype=np.uint8), np.array([], dtype=np.float32)
        dib, soft, cnt = self._bank.demodulate(iq.reshape(1, -1))
        n = int(cnt[0])
        return dib[0, :n].copy(), soft[0, :n].copy()

    def get_timing_offset(self) -> float:
        return float(self._bank.state(0)["pll"])

    def demodulate_discriminator(self, disc_audio):
        """c4fm.py:2817-2992: discriminator audio (radians/sample) -> (dibits uint8, soft symbols float32). RRC in
        float64 with scipy's carried state (first call: lfilter_zi * audio[0]), phases = filtered * sps, the shared
        symbol recovery, and the discriminator flavour of the sync loop (csrc/p25.cu c4fm_sync_kernel<true>)."""
        a = np.asarray(disc_audio)
        if a.size == 0:
            return np.array([], dtype=np.uint8), np.array([], dtype=np.float32)
        if a.ndim > 1:  # "take first channel if stereo"
            a = a[:, 0]
        dib, soft, cnt = self._bank.demodulate_discriminator(a.reshape(1, -1))
        n = int(cnt[0])
        return dib[0, :n].copy(), soft[0, :n].copy()

    @property
    def _equalizer(self) -> _EqualizerView:
        return _EqualizerView(self._bank, 0)

    @property
    def _sample_point(self) -> float:
        return float(self._bank.state(0)["sample_point"])

    @property
    def _ted_phase(self) -> float:
        """API compatibility of the reference (c4fm.py:2523-2526): fixed symbol timing, no Gardner phase."""
        return 0.0

    @property
    def _sync_count(self) -> int:
        return self._bank.state(0)["sync_count"]

    @property
    def _fine_sync(self) -> bool:
        return self._bank.state(0)["fine_sync"]


def c4fm_demod_simple(iq, sample_rate: int = 19200, symbol_rate: int = 4800):
    """Stateless single-shot form (c4fm.py:2995-3015): dibits of one fresh demodulator."""
    dibits, _ = C4FMDemodulator(sample_rate=sample_rate, symbol_rate=symbol_rate).demodulate(iq)
    return dibits


# ---- helper classes timed by backend/benchmark_dsp.py:17-114 (same names and call shapes) -----------------------

class _FMDemodulator:
    """Symbol-spaced differential demodulator (c4fm.py:276-395) on the GPU. `symbol_delay` is accepted as an alias of
    `samples_per_symbol` — benchmark_dsp.py:27 still passes the old keyword, which raises TypeError in the reference."""

    def __init__(self, samples_per_symbol: float = 10.0, symbol_delay: float | None = None):
        if symbol_delay is not None:
            samples_per_symbol = float(symbol_delay)
        self.samples_per_symbol = float(samples_per_symbol)
        self._bank = C4FMBank(1, int(round(self.samples_per_symbol * 4800)), 4800)
        sps = C.c_double()
        N.check(N.lib().wc_c4fm_info(self._bank._h, None, C.byref(sps), None, None))
        assert abs(sps.value - self.samples_per_symbol) < 1e-9, "samples_per_symbol must be sample_rate / 4800 for an integer rate"

    def reset(self) -> None:
        self._bank.reset(0)

    def demodulate(self, i, q):
        import torch

        i = np.asarray(i, dtype=np.float32)
        q = np.asarray(q, dtype=np.float32)
        n = len(i)
        if n == 0:
            return np.array([], dtype=np.float32)
        pairs = torch.from_numpy(np.ascontiguousarray(np.stack([i, q], axis=1))).cuda()
        out = torch.empty((n,), dtype=torch.float32, device="cuda")
        N.check(N.lib().wc_c4fm_diffdemod(self._bank._h, C.c_void_p(pairs.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                          N.torch_stream_ptr()))
        return out.cpu().numpy()


class _Interpolator:
    """8-tap, 128-step fractional interpolator (c4fm.py:891-2253); `filter_batch` evaluates many positions per launch."""

    NTAPS = 8
    NSTEPS = 128
    # the 129 x 8 table of the reference class (c4fm.py:907-2202), extracted by oracle/make_tables.py; the kernels carry the
    # same numbers (csrc/interp_taps_129x8.inc)
    TAPS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "interp_taps_129x8.npy"))

    def filter_batch(self, samples, offsets, mus) -> np.ndarray:
        import torch

        N.ensure_init()
        x = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float32)).cuda()
        o = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int32)).cuda()
        m = torch.from_numpy(np.ascontiguousarray(mus, dtype=np.float64)).cuda()
        out = torch.empty((o.numel(),), dtype=torch.float64, device="cuda")
        N.check(N.lib().wc_c4fm_interp(C.c_void_p(x.data_ptr()), int(x.numel()), C.c_void_p(o.data_ptr()),
                                       C.c_void_p(m.data_ptr()), int(o.numel()), C.c_void_p(out.data_ptr()),
                                       N.torch_stream_ptr()))
        return out.cpu().numpy()

    def filter(self, samples, offset: int, mu: float) -> float:
        return float(self.filter_batch(samples, [int(offset)], [float(mu)])[0])


class _SoftSyncDetector:
    """24-symbol soft sync correlator (c4fm.py:2268-2329); `process_block` scores a whole block per launch."""

    SYNC_PATTERN = 0x5575F5FF77FF
    SYNC_THRESHOLD = 130.0

    def __init__(self) -> None:
        N.ensure_init()
        self.reset()

    def reset(self) -> None:
        self._hist = np.zeros(24, dtype=np.float32)

    def process_block(self, soft) -> np.ndarray:
        import torch

        s = torch.from_numpy(np.ascontiguousarray(soft, dtype=np.float32).reshape(-1)).cuda()
        n = int(s.numel())
        if n == 0:
            return np.zeros(0, dtype=np.float64)
        h = torch.from_numpy(self._hist).cuda()
        nh = torch.empty_like(h)
        sc = torch.empty((n,), dtype=torch.float64, device="cuda")
        N.check(N.lib().wc_c4fm_sync_scores(C.c_void_p(s.data_ptr()), n, C.c_void_p(h.data_ptr()), C.c_void_p(sc.data_ptr()),
                                            C.c_void_p(nh.data_ptr()), N.torch_stream_ptr()))
        self._hist = nh.cpu().numpy()
        return sc.cpu().numpy()

    def process(self, soft_symbol: float) -> float:
        return float(self.process_block([soft_symbol])[0])
