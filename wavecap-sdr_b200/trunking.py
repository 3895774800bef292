"""Trunking fan-out on the GPU: K narrowband channels out of one wideband chunk per call.

What `TrunkingSystem.on_raw_iq_callback` (wavecapsdr/trunking/system.py:1558-1870) does for its control
channel — phase-continuous NCO (:1434-1466), Kaiser(7.857) 157-tap /s1 and 73-tap /s2 decimators
(:1392-1406, :1753-1779) — and what every `VoiceRecorder.process_iq` (:561-656) repeats per active call,
batched over channels in one pass that reads the wideband chunk from HBM once (csrc/firdec.cu, `wc_ddc_*`).

`DDCBank(..., flavor="control")` reproduces the control-channel arithmetic (fir_decimate state quirk, complex64
between stages); `flavor="voice"` the recorder's (scipy lfilter steady-state start, complex128 throughout).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class DDCBank:
    def __init__(self, n_channels: int, sample_rate: int, decim1: int, decim2: int = 1, flavor: str = "control",
                 taps1=None, taps2=None):
        N.ensure_init()
        assert flavor in ("control", "voice", "zeros")
        self.n_channels, self.sample_rate, self.decim1, self.decim2 = int(n_channels), int(sample_rate), int(decim1), int(decim2)
        self.flavor = flavor
        init_mode = {"zeros": 0, "control": 1, "voice": 2}[flavor]
        self.out_dtype = np.complex128 if flavor == "voice" else np.complex64
        t1 = None if taps1 is None else np.ascontiguousarray(taps1, dtype=np.float64)
        t2 = None if taps2 is None else np.ascontiguousarray(taps2, dtype=np.float64)
        h = C.c_void_p()
        N.check(N.lib().wc_ddc_create(self.n_channels, self.sample_rate, None if t1 is None else N.np_ptr(t1),
                                      0 if t1 is None else len(t1), self.decim1, None if t2 is None else N.np_ptr(t2),
                                      0 if t2 is None else len(t2), self.decim2, init_mode,
                                      1 if flavor == "voice" else 0, C.byref(h)))
        self._h = h
        self.offsets_hz = np.zeros(self.n_channels, dtype=np.float64)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_ddc_destroy(h)
            except Exception:
                pass

    @property
    def output_rate(self) -> int:
        return self.sample_rate // self.decim1 // self.decim2

    def taps(self):
        n1, n2 = C.c_int(), C.c_int()
        N.check(N.lib().wc_ddc_get_taps(self._h, None, None, C.byref(n1), C.byref(n2)))
        t1, t2 = np.zeros(n1.value), np.zeros(max(n2.value, 1))
        N.check(N.lib().wc_ddc_get_taps(self._h, N.np_ptr(t1), N.np_ptr(t2), None, None))
        return t1, (t2[: n2.value] if n2.value else None)

    def set_offsets(self, offsets_hz) -> None:
        o = np.ascontiguousarray(offsets_hz, dtype=np.float64)
        assert o.shape == (self.n_channels,)
        self.offsets_hz = o
        N.check(N.lib().wc_ddc_set_offsets(self._h, N.np_ptr(o)))

    def reset(self, channel: int = -1) -> None:
        N.check(N.lib().wc_ddc_reset(self._h, int(channel)))

    def out_len(self, n_samples: int) -> int:
        return int(N.lib().wc_ddc_out_len(self._h, int(n_samples)))

    def process(self, iq):
        """iq complex64 [n] (numpy or torch CUDA) -> [K][out_len(n)] narrowband IQ (complex64, or complex128 for voice)."""
        if N.is_torch_cuda(iq):
            import torch

            x = iq.to(torch.complex64).contiguous().reshape(-1)
            n = int(x.numel())
            m = self.out_len(n)
            out = torch.empty((self.n_channels, m), dtype=torch.complex128 if self.flavor == "voice" else torch.complex64,
                              device=x.device)
            if n:
                N.check(N.lib().wc_ddc_process(self._h, C.c_void_p(x.data_ptr()), n, C.c_void_p(out.data_ptr()), m,
                                               N.torch_stream_ptr()))
            return out
        x = np.ascontiguousarray(iq, dtype=np.complex64).reshape(-1)
        n = int(x.size)
        m = self.out_len(n)
        out = np.empty((self.n_channels, m), dtype=self.out_dtype)
        if n:
            N.check(N.lib().wc_ddc_process_host(self._h, N.np_ptr(x), n, N.np_ptr(out)))
        return out


class VoiceDiscriminator:
    """FM discriminator of `VoiceRecorder.process_iq` (trunking/system.py:708-717) for C channels at once:
    np.diff(np.unwrap([last_phase | np.angle(iq)])) with the last phase carried across calls. Input [C][n] complex64 or
    complex128 (numpy or torch CUDA); output float64 [C][n] (same residency as the input)."""

    def __init__(self, n_channels: int):
        import torch

        N.ensure_init()
        self.n_channels = int(n_channels)
        self._last = torch.zeros((self.n_channels,), dtype=torch.float64, device="cuda")

    def reset(self) -> None:
        self._last.zero_()

    def process(self, iq):
        import torch

        on_dev = N.is_torch_cuda(iq)
        x = iq if on_dev else torch.from_numpy(np.ascontiguousarray(iq)).cuda()
        assert x.dim() == 2 and x.shape[0] == self.n_channels, x.shape
        if x.dtype not in (torch.complex64, torch.complex128):
            x = x.to(torch.complex64)
        x = x.contiguous()
        n = int(x.shape[1])
        out = torch.empty((self.n_channels, n), dtype=torch.float64, device=x.device)
        if n:
            N.check(N.lib().wc_fm_discriminator(C.c_void_p(x.data_ptr()), 1 if x.dtype == torch.complex128 else 0, n, n,
                                                self.n_channels, C.c_void_p(self._last.data_ptr()),
                                                C.c_void_p(out.data_ptr()), N.torch_stream_ptr()))
        return out if on_dev else out.cpu().numpy()
