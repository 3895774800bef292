"""Control-channel scanner on the GPU — call surface of `wavecapsdr.trunking.cc_scanner` (ChannelMeasurement,
ControlChannelScanner.scan_all / get_best_channel / get_channel_ranking / should_roam). One `scan_all` is ONE library call
(`wc_ccscan_measure`, csrc/ccscan.cu): every in-range candidate plus the two band-edge noise probes are shifted, low-pass
filtered, decimated, measured and sync-correlated from the same device-resident block; the reference runs the whole chain,
noise probes included, once per candidate on the CPU (trunking/cc_scanner.py:166-277).
"""
from __future__ import annotations

import asyncio
import ctypes as C
import logging
import time
from dataclasses import dataclass, field

import numpy as np

from . import _native as N
from .dsp import _stages as S

_log = logging.getLogger(__name__)


@dataclass
class ChannelMeasurement:
    frequency_hz: float
    power_db: float
    peak_power_db: float
    noise_floor_db: float
    snr_db: float
    sync_detected: bool
    measurement_time: float
    sample_count: int

    def __str__(self) -> str:
        return (f"{self.frequency_hz / 1e6:.4f} MHz: power={self.power_db:.1f} dB, SNR={self.snr_db:.1f} dB, "
                f"{'SYNC' if self.sync_detected else '----'}")


def _scanner_taps(decim: int) -> np.ndarray:
    from scipy import signal

    return np.ascontiguousarray(signal.firwin(65, 0.8 / decim, window=("kaiser", 6.0)), dtype=np.float64)


def measure_offsets(iq, sample_rate: int, offsets_hz):
    """Raw measurement for a list of frequency offsets: (power_mean, power_max, correlation, sample_count) arrays."""
    import torch

    x = S.to_device(iq, np.complex64)
    n = int(x.numel())
    offs = np.ascontiguousarray(np.asarray(offsets_hz, dtype=np.float64))
    k = int(offs.size)
    lib = N.lib()
    m = int(lib.wc_ccscan_out_len(n, int(sample_rate)))
    decim = max(1, int(sample_rate) // 48000)
    y = torch.empty((k, max(m, 1)), dtype=torch.complex128, device=x.device)
    psum = torch.empty((k,), dtype=torch.float64, device=x.device)
    pmax = torch.empty((k,), dtype=torch.float64, device=x.device)
    corr = torch.empty((k,), dtype=torch.float64, device=x.device)
    scratch = torch.empty((8 * k,), dtype=torch.uint8, device=x.device)
    taps = _scanner_taps(decim)
    N.check(lib.wc_ccscan_measure(S.ptr(x), n, int(sample_rate), N.np_ptr(offs), k, N.np_ptr(taps), S.ptr(y), S.ptr(psum),
                                  S.ptr(pmax), S.ptr(corr), S.ptr(scratch), S.stream()))
    mean = psum.cpu().numpy() / max(m, 1)
    return mean, pmax.cpu().numpy(), corr.cpu().numpy(), m


@dataclass
class ControlChannelScanner:
    center_hz: float
    sample_rate: int
    control_channels: list
    measurement_samples: int = 48000 * 2
    channel_bandwidth: float = 12500
    min_snr_db: float = 6.0
    sync_check_enabled: bool = True
    _last_scan_time: float = 0.0
    _measurements: dict = field(default_factory=dict)
    _current_channel_hz: float | None = None
    # fields of the reference dataclass kept for its callers (trunking/system.py reads and assigns scanner attributes):
    # the reference never fills the buffer fields either (cc_scanner.py:91-98)
    _iq_buffer: list = field(default_factory=list)
    _buffer_samples: int = 0
    _measurement_in_progress: bool = False
    _measurement_complete: asyncio.Event = field(default_factory=asyncio.Event)
    _sync_pattern: np.ndarray | None = None

    def __post_init__(self) -> None:
        N.ensure_init()
        # +3 -> dibit 1, -3 -> dibit 3 (cc_scanner.py:102-107); the kernel carries the same word (csrc/ccscan.cu)
        self._sync_pattern = np.array([1, 1, 1, 1, 1, 3, 1, 1, 3, 3, 1, 1, 3, 3, 3, 3, 1, 3, 1, 3, 3, 3, 3, 3], dtype=np.uint8)
        _log.info(f"ControlChannelScanner initialized: center={self.center_hz / 1e6:.4f} MHz, "
                  f"channels={len(self.control_channels)}")

    def get_channel_offset(self, freq_hz: float) -> float:
        return freq_hz - self.center_hz

    def is_channel_in_range(self, freq_hz: float) -> bool:
        return abs(self.get_channel_offset(freq_hz)) <= self.sample_rate / 2 - self.channel_bandwidth

    def scan_all(self, iq) -> dict:
        """cc_scanner.py:128-164; out-of-range frequencies are skipped like there."""
        freqs = [f for f in self.control_channels if self.is_channel_in_range(f)]
        out: dict = {}
        if freqs:
            edge = self.sample_rate / 2 - 15000   # cc_scanner.py:217-221
            offsets = [self.get_channel_offset(f) for f in freqs] + [-edge + 25000, edge - 25000]
            mean, peak, corr, m = measure_offsets(iq, self.sample_rate, offsets)
            noise = min(mean[-2], mean[-1])
            eps = 1e-12
            floor_db = 10 * np.log10(noise + eps)
            now = time.time()
            for i, f in enumerate(freqs):
                power_db = 10 * np.log10(mean[i] + eps)
                snr_db = power_db - floor_db
                sync = bool(self.sync_check_enabled and m > 0 and snr_db >= 8.0 and abs(corr[i]) > 0.6)
                out[f] = ChannelMeasurement(frequency_hz=f, power_db=float(power_db),
                                            peak_power_db=float(10 * np.log10(peak[i] + eps)),
                                            noise_floor_db=float(floor_db), snr_db=float(snr_db), sync_detected=sync,
                                            measurement_time=now, sample_count=m)
        self._measurements.update(out)
        self._last_scan_time = time.time()
        return out

    def _ranked(self) -> list:
        items = list(self._measurements.items())
        key = lambda kv: kv[1].snr_db
        return (sorted((kv for kv in items if kv[1].sync_detected), key=key, reverse=True)
                + sorted((kv for kv in items if not kv[1].sync_detected), key=key, reverse=True))

    def get_best_channel(self):
        r = self._ranked()
        return r[0] if r else None

    def get_channel_ranking(self) -> list:
        return self._ranked()

    def should_roam(self, current_freq_hz: float, roam_threshold_db: float = 6.0):
        cur = self._measurements.get(current_freq_hz)
        best = self.get_best_channel()
        if cur is None or best is None or best[0] == current_freq_hz:
            return None
        if (not cur.sync_detected and best[1].sync_detected) or best[1].snr_db - cur.snr_db >= roam_threshold_db:
            return best[0]
        return None

    def log_scan_results(self) -> None:
        """cc_scanner.py:471-485"""
        if not self._measurements:
            _log.info("No control channel measurements available")
            return
        _log.info("Control Channel Scan Results:")
        _log.info("-" * 60)
        for i, (_freq, m) in enumerate(self.get_channel_ranking()):
            _log.info(f"  {i + 1}. {m}{' *' if i == 0 else ''}")
        _log.info("-" * 60)

    def get_stats(self) -> dict:
        """cc_scanner.py:487-505"""
        return {
            "channels_configured": len(self.control_channels),
            "channels_measured": len(self._measurements),
            "last_scan_time": float(self._last_scan_time),
            "current_channel_hz": float(self._current_channel_hz) if self._current_channel_hz else None,
            "measurements": {f"{freq / 1e6:.4f}_MHz": {"power_db": float(m.power_db), "snr_db": float(m.snr_db),
                                                      "sync_detected": bool(m.sync_detected)}
                             for freq, m in self._measurements.items()},
        }
