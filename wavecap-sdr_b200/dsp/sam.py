"""GPU synchronous-AM demodulation with the call surface of `wavecapsdr.dsp.sam` (dsp/sam.py).

The carrier-recovery loop is sequential per sequence (csrc/analog.cu sam_pll_kernel, one thread per sequence; float64 like
the reference's Python loop by default, `EXACT = False` selects the float32-detector flavour); the batch path in capture.py runs all (channel, chunk) sequences of a call side by side. The
tail after the sideband selection is the AM tail (`am.am_tail`): same stages, same order (dsp/sam.py:223-258).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .. import _native as N
from . import _stages as S
from . import am as AM

SIDEBAND = {"dsb": 0, "usb": 1, "lsb": 2}
# True (default): float64 replay of the reference's per-sample arithmetic (coherent I/Q bit-equal to the reference on the
# goldens). False: float32 oscillator / mixer / phase detector around the float64 loop state, several times faster per sample;
# same audio to ~1e-6 relative RMS on a single carrier, up to 2e-4 when strong neighbours drive the mixed vector through the
# origin, where the reference's arctan2(Q, |I|) detector amplifies any last-bit difference (tests/test_sam_gpu.py).
EXACT = True


def _n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def pll_coefficients(sample_rate: float, loop_bandwidth: float, damping: float) -> tuple[float, float]:
    """alpha (proportional), beta (integral) of the second-order loop (dsp/sam.py:54-65)."""
    omega_n = 2 * np.pi * loop_bandwidth
    return float(2 * damping * omega_n / sample_rate), float((omega_n ** 2) / (sample_rate ** 2))


def pll_rows(rows, alpha: float, beta: float, sideband: int = 0, state=None, want_coherent: bool = False, exact=None):
    """CarrierRecoveryPLL.process on every row of a CUDA complex64 [n_seq, n] tensor at once.
    state: CUDA float64 [n_seq, 3] (phase, frequency, integrator), updated in place; None = fresh PLLs.
    -> (audio float32 [n_seq, n] after the sideband selection, state, (coherent_i, coherent_q) | None)."""
    import torch

    n_seq, n = int(rows.shape[0]), int(rows.shape[1])
    if state is None:
        state = torch.zeros((n_seq, 3), dtype=torch.float64, device=rows.device)
    audio = torch.empty((n_seq, n), dtype=torch.float32, device=rows.device)
    ci = cq = None
    if want_coherent:
        ci, cq = torch.empty_like(audio), torch.empty_like(audio)
    N.check(N.lib().wc_sam_pll(S.ptr(rows), int(rows.stride(0)), n, n_seq, float(alpha), float(beta), int(sideband),
                               int(EXACT if exact is None else exact), S.ptr(state),
                               S.ptr(audio), S.ptr(ci), S.ptr(cq), S.stream()))
    return audio, state, ((ci, cq) if want_coherent else None)


@dataclass
class CarrierRecoveryPLL:
    """Second-order PLL for AM carrier recovery (dsp/sam.py:25-129): same fields, same state attributes."""

    sample_rate: float
    loop_bandwidth: float = 50.0
    damping: float = 0.707

    _phase: float = field(default=0.0, init=False)
    _frequency: float = field(default=0.0, init=False)
    _integrator: float = field(default=0.0, init=False)
    _alpha: float = field(default=0.0, init=False)
    _beta: float = field(default=0.0, init=False)

    def __post_init__(self) -> None:
        self._compute_coefficients()

    def _compute_coefficients(self) -> None:
        self._alpha, self._beta = pll_coefficients(self.sample_rate, self.loop_bandwidth, self.damping)

    def set_bandwidth(self, bandwidth_hz: float) -> None:
        self.loop_bandwidth = bandwidth_hz
        self._compute_coefficients()

    def _run(self, iq, sideband: int, want_coherent: bool):
        import torch

        x = S.to_device(iq, np.complex64).reshape(1, -1)
        st = torch.tensor([[self._phase, self._frequency, self._integrator]], dtype=torch.float64, device=x.device)
        audio, st, coh = pll_rows(x, self._alpha, self._beta, sideband, st, want_coherent)
        self._phase, self._frequency, self._integrator = (float(v) for v in st[0].cpu().numpy())
        return audio, coh, self._frequency * self.sample_rate / (2 * np.pi)

    def process(self, iq):
        """-> (coherent_i float32, coherent_q float32, freq_offset_hz) (dsp/sam.py:73-123)."""
        if _n(iq) == 0:
            return np.empty(0, dtype=np.float32), np.empty(0, dtype=np.float32), 0.0
        _, (ci, cq), f = self._run(iq, 0, True)
        return S.like_input(ci.reshape(-1), iq), S.like_input(cq.reshape(-1), iq), f

    def reset(self) -> None:
        self._phase = self._frequency = self._integrator = 0.0


def sam_demod(iq, sample_rate: int, audio_rate: int = 48_000, sideband: str = "dsb", pll_bandwidth: float = 50.0,
              pll_damping: float = 0.707, enable_agc: bool = True, enable_highpass: bool = True, highpass_hz: float = 100.0,
              enable_lowpass: bool = True, lowpass_hz: float = 5000.0, enable_noise_blanker: bool = False,
              noise_blanker_threshold_db: float = 10.0, agc_target_db: float = -20.0, notch_frequencies=None, pll_state=None):
    """-> (audio float32, carrier offset in Hz, pll_state) (dsp/sam.py:132-270)."""
    if _n(iq) == 0:
        return np.empty(0, dtype=np.float32), 0.0, pll_state
    if pll_state is None:
        pll_state = CarrierRecoveryPLL(sample_rate=float(sample_rate), loop_bandwidth=pll_bandwidth, damping=pll_damping)
    elif pll_state.loop_bandwidth != pll_bandwidth:
        pll_state.set_bandwidth(pll_bandwidth)
    audio, _, freq_offset = pll_state._run(iq, SIDEBAND.get(sideband.lower(), 0), False)
    stages = AM.am_post_chain(sample_rate, enable_highpass, highpass_hz, enable_lowpass, lowpass_hz, notch_frequencies)
    out = AM.am_tail(audio, int(sample_rate), int(audio_rate), stages, enable_agc, agc_target_db,
                     blanker_db=noise_blanker_threshold_db if enable_noise_blanker else None)
    return S.like_input(out.reshape(-1), iq), freq_offset, pll_state


def sam_demod_simple(iq, sample_rate: int, audio_rate: int = 48_000, sideband: str = "dsb", pll_bandwidth: float = 50.0,
                     enable_agc: bool = True, enable_highpass: bool = True, highpass_hz: float = 100.0,
                     enable_lowpass: bool = True, lowpass_hz: float = 5000.0, agc_target_db: float = -20.0):
    """Stateless wrapper returning only the audio (dsp/sam.py:273-318)."""
    audio, _, _ = sam_demod(iq=iq, sample_rate=sample_rate, audio_rate=audio_rate, sideband=sideband, pll_bandwidth=pll_bandwidth,
                            enable_agc=enable_agc, enable_highpass=enable_highpass, highpass_hz=highpass_hz,
                            enable_lowpass=enable_lowpass, lowpass_hz=lowpass_hz, agc_target_db=agc_target_db)
    return audio
