"""GPU CQPSK/LSM demodulator with the call surface of `wavecapsdr.decoders.p25.CQPSKDemodulator`.

Mirrors wavecapsdr/decoders/p25.py:190-669 (constructor arguments, `demodulate(iq) -> uint8 dibits`,
empty input -> empty output). All signal arithmetic runs in csrc/cqpsk.cu through the C ABI
(`wc_cqpsk_*`); the 63-tap low-pass is designed on the host at construction like the reference
(scipy.signal.firwin, :375-388), never on the data path.

GPU-only addition: `CQPSKBank` — C independent stateful demodulators advanced by one call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _native as N


def _baseband_taps(sample_rate: int, cutoff_hz: float = 7250, num_taps: int = 63) -> np.ndarray:
    from scipy import signal

    norm = min(0.99, max(0.01, cutoff_hz / (sample_rate / 2)))
    return np.asarray(signal.firwin(num_taps, norm, window="hamming"), dtype=np.float32)


def _mmse_taps() -> np.ndarray:
    """decoders/p25.py:289-323: 129 x 8 Hann-windowed sinc rows normalised to unit sum (float32)."""
    steps = np.arange(129, dtype=np.float64) / 128
    taps = np.zeros((129, 8), dtype=np.float32)
    for step in range(129):
        for tap in range(8):
            t = tap - 3 - steps[step]
            if abs(t) < 1e-6:
                taps[step, tap] = 1.0
            else:
                w = 0.5 * (1 + np.cos(np.pi * t / 4)) if abs(t) < 4 else 0
                taps[step, tap] = (np.sin(np.pi * t) / (np.pi * t)) * w
        s = np.sum(taps[step])
        if abs(s) > 1e-6:
            taps[step] /= s
    return taps


class CQPSKBank:
    """`n_channels` CQPSK demodulators with persistent per-channel state."""

    BASEBAND_CUTOFF_HZ = 7250

    def __init__(self, n_channels: int, sample_rate: int = 19200, symbol_rate: int = 4800):
        N.ensure_init()
        self.n_channels = int(n_channels)
        self.sample_rate = sample_rate
        self.symbol_rate = symbol_rate
        self.samples_per_symbol = sample_rate / symbol_rate
        self._baseband_taps = _baseband_taps(sample_rate, self.BASEBAND_CUTOFF_HZ)
        self._mmse_taps = _mmse_taps()
        h = C.c_void_p()
        N.check(N.lib().wc_cqpsk_create(self.n_channels, int(sample_rate), int(symbol_rate), N.np_ptr(self._baseband_taps),
                                        N.np_ptr(self._mmse_taps), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_cqpsk_destroy(h)
            except Exception:
                pass

    def max_symbols(self, n_samples: int) -> int:
        return int(N.lib().wc_cqpsk_max_symbols(self._h, int(n_samples)))

    def reset(self, channel: int = -1) -> None:
        N.check(N.lib().wc_cqpsk_reset(self._h, int(channel)))

    def state(self, channel: int = 0) -> dict:
        s = np.zeros(6, dtype=np.float64)
        N.check(N.lib().wc_cqpsk_get_state(self._h, int(channel), N.np_ptr(s)))
        return {"freq_offset": s[0], "phase_acc": s[1], "symbol_clock": s[2], "symbol_time": s[3], "agc_gain": s[4],
                "omega": s[5]}

    def demodulate(self, iq):
        """iq complex64 [n_channels][n] (numpy or torch CUDA) -> (dibits uint8 [C][max_sym], counts int32 [C])."""
        if N.is_torch_cuda(iq):
            import torch

            x = iq.to(torch.complex64).contiguous()
            assert x.dim() == 2 and x.shape[0] == self.n_channels, x.shape
            n = int(x.shape[1])
            ms = max(1, self.max_symbols(n))
            dib = torch.zeros((self.n_channels, ms), dtype=torch.uint8, device=x.device)
            cnt = torch.zeros((self.n_channels,), dtype=torch.int32, device=x.device)
            if n:
                N.check(N.lib().wc_cqpsk_demod(self._h, C.c_void_p(x.data_ptr()), n, n, C.c_void_p(dib.data_ptr()),
                                               C.c_void_p(cnt.data_ptr()), ms, N.torch_stream_ptr()))
            return dib, cnt
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        assert x.ndim == 2 and x.shape[0] == self.n_channels, x.shape
        n = int(x.shape[1])
        ms = max(1, self.max_symbols(n))
        dib = np.zeros((self.n_channels, ms), dtype=np.uint8)
        cnt = np.zeros((self.n_channels,), dtype=np.int32)
        if n:
            N.check(N.lib().wc_cqpsk_demod_host(self._h, N.np_ptr(x), n, N.np_ptr(dib), N.np_ptr(cnt), ms))
        return dib, cnt


class CQPSKDemodulator:
    """Drop-in for wavecapsdr.decoders.p25.CQPSKDemodulator (one channel)."""

    BASEBAND_CUTOFF_HZ = 7250
    MMSE_NTAPS = 32
    MMSE_NSTEPS = 128
    EQUALIZER_GAIN = 1.0   # decoders/p25.py:216

    def __init__(self, sample_rate: int = 19200, symbol_rate: int = 4800) -> None:
        self.quarter_pi, self.half_pi, self.three_quarter_pi = np.pi / 4, np.pi / 2, 3 * np.pi / 4   # :232-234
        self._bank = CQPSKBank(1, sample_rate, symbol_rate)
        self.sample_rate = sample_rate
        self.symbol_rate = symbol_rate
        self.samples_per_symbol = sample_rate / symbol_rate
        self._baseband_taps = self._bank._baseband_taps
        self._mmse_taps = self._bank._mmse_taps

    def demodulate(self, iq):
        """decoders/p25.py:413-479: complex IQ -> dibits (uint8)."""
        iq = np.asarray(iq)
        if iq.size == 0:
            return np.array([], dtype=np.uint8)
        if not np.iscomplexobj(iq):  # the reference's interleaved-real rescue (:428-433)
            if len(iq) % 2 == 0:
                iq = iq[::2] + 1j * iq[1::2]
            else:
                return np.array([], dtype=np.uint8)
        dib, cnt = self._bank.demodulate(iq.reshape(1, -1))
        return dib[0, : int(cnt[0])].copy()

    @property
    def _freq_offset(self) -> float:
        return float(self._bank.state(0)["freq_offset"])

    @property
    def _agc_gain(self) -> float:
        return float(self._bank.state(0)["agc_gain"])

    @property
    def _symbol_clock(self) -> float:
        return float(self._bank.state(0)["symbol_clock"])


# ---------------------------------------------------------------------------------------------------
# DiscriminatorDemodulator (decoders/p25.py:1105-1345) on the GPU: csrc/discdemod.cu
# ---------------------------------------------------------------------------------------------------
def _discriminator_lpf(sample_rate: int) -> np.ndarray:
    """_design_baseband_filter (decoders/p25.py:1188-1195): firwin(65, 5200 Hz, hamming), float32."""
    from scipy import signal

    cutoff = min(5200 / (sample_rate / 2), 0.99)
    return np.asarray(signal.firwin(65, cutoff, window="hamming"), dtype=np.float32)


class DiscriminatorBank:
    """C stateful discriminator-audio demodulators advanced by one call: audio [C][n] -> (dibits uint8 [C][max_sym],
    soft float32 [C][max_sym] (the slicer input), counts int32 [C])."""

    def __init__(self, n_channels: int, sample_rate: int = 48000, symbol_rate: int = 4800):
        N.ensure_init()
        self.n_channels, self.sample_rate, self.symbol_rate = int(n_channels), int(sample_rate), int(symbol_rate)
        self._mmse_taps = np.ascontiguousarray(_mmse_taps())
        self._baseband_taps = np.ascontiguousarray(_discriminator_lpf(sample_rate))
        h = C.c_void_p()
        N.check(N.lib().wc_discdemod_create(self.n_channels, self.sample_rate, self.symbol_rate, N.np_ptr(self._mmse_taps),
                                            N.np_ptr(self._baseband_taps), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                N.lib().wc_discdemod_destroy(h)
            except Exception:
                pass
            self._h = None

    def max_symbols(self, n_samples: int) -> int:
        return int(N.lib().wc_discdemod_max_symbols(self._h, int(n_samples)))

    def reset(self, channel: int = -1) -> None:
        N.check(N.lib().wc_discdemod_reset(self._h, int(channel)))

    def state(self, channel: int = 0) -> dict:
        s = np.zeros(8, dtype=np.float64)
        N.check(N.lib().wc_discdemod_get_state(self._h, int(channel), N.np_ptr(s)))
        keys = ("input_gain", "dc_estimate", "symbol_clock", "symbol_spread", "fine_freq_correction",
                "coarse_freq_correction", "clock_is_python_float", "symbol_count")
        return dict(zip(keys, (float(v) for v in s)))

    def demodulate(self, audio):
        if N.is_torch_cuda(audio):
            import torch

            x = audio.to(torch.float32).contiguous()
            assert x.dim() == 2 and x.shape[0] == self.n_channels, x.shape
            n = int(x.shape[1])
            ms = max(1, self.max_symbols(n))
            dib = torch.zeros((self.n_channels, ms), dtype=torch.uint8, device=x.device)
            soft = torch.zeros((self.n_channels, ms), dtype=torch.float32, device=x.device)
            cnt = torch.zeros((self.n_channels,), dtype=torch.int32, device=x.device)
            if n:
                N.check(N.lib().wc_discdemod_demod(self._h, C.c_void_p(x.data_ptr()), n, n, C.c_void_p(dib.data_ptr()),
                                                   C.c_void_p(soft.data_ptr()), C.c_void_p(cnt.data_ptr()), ms,
                                                   N.torch_stream_ptr()))
            return dib, soft, cnt
        x = np.ascontiguousarray(np.asarray(audio).astype(np.float32, copy=False))
        assert x.ndim == 2 and x.shape[0] == self.n_channels, x.shape
        n = int(x.shape[1])
        ms = max(1, self.max_symbols(n))
        dib = np.zeros((self.n_channels, ms), dtype=np.uint8)
        soft = np.zeros((self.n_channels, ms), dtype=np.float32)
        cnt = np.zeros((self.n_channels,), dtype=np.int32)
        if n:
            N.check(N.lib().wc_discdemod_demod_host(self._h, N.np_ptr(x), n, N.np_ptr(dib), N.np_ptr(soft), N.np_ptr(cnt), ms))
        return dib, soft, cnt


class DiscriminatorDemodulator:
    """Drop-in for wavecapsdr.decoders.p25.DiscriminatorDemodulator (one channel)."""

    BASEBAND_CUTOFF_HZ = 5200
    MMSE_NTAPS = 8
    MMSE_NSTEPS = 128

    def __init__(self, sample_rate: int = 48000, symbol_rate: int = 4800) -> None:
        self._bank = DiscriminatorBank(1, sample_rate, symbol_rate)
        self.sample_rate = sample_rate
        self.symbol_rate = symbol_rate
        self.samples_per_symbol = sample_rate / symbol_rate
        self._mmse_taps = self._bank._mmse_taps
        self._baseband_taps = self._bank._baseband_taps

    def demodulate(self, audio):
        """decoders/p25.py:1197-1236: mono discriminator audio -> dibits (uint8)."""
        a = np.asarray(audio)
        if a.size == 0:
            return np.array([], dtype=np.uint8)
        dib, _soft, cnt = self._bank.demodulate(a.reshape(1, -1))
        return dib[0, : int(cnt[0])].copy()

    def reset(self) -> None:
        self._bank.reset(0)

    @property
    def _symbol_spread(self) -> float:
        return self._bank.state(0)["symbol_spread"]

    @property
    def _symbol_clock(self) -> float:
        return self._bank.state(0)["symbol_clock"]

    @property
    def _input_gain(self) -> float:
        return self._bank.state(0)["input_gain"]

    @property
    def _symbol_count(self) -> int:
        return int(self._bank.state(0)["symbol_count"])


class P25TrellisDecoder:
    """Drop-in for wavecapsdr.decoders.p25.P25TrellisDecoder (decoders/p25.py:1348-1393): 1/2-rate Viterbi on the GPU,
    truncated to the 48 dibits of a TSBK."""

    def __init__(self) -> None:
        from ..dsp.fec.trellis import TrellisDecoder

        self._decoder = TrellisDecoder()

    def decode(self, dibits):
        if len(dibits) < 4:
            return None, -1
        decoded, error_metric = self._decoder.decode(dibits)
        if len(decoded) == 0:
            return None, -1
        return decoded[:48], int(error_metric)
