"""The "cuda" FFT backend: fused window -> FFT -> |X| -> dB -> fftshift (-> K-frame dB mean) kernel
(csrc/spectrum.cu). Takes the registry slot the reference fills with CuPy/cuFFT
(dsp/fft/cupy_backend.py:35-125, registry.py:167-174)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _native as N
from .base import FFTBackend, FFTResult


class CudaFFTBackend(FFTBackend):
    def __init__(self, fft_size: int = 2048):
        super().__init__(fft_size)
        N.ensure_init()
        h = C.c_void_p()
        N.check(N.lib().wc_spectrum_create(int(fft_size), C.byref(h)))
        self._h = h
        w = np.empty(fft_size, dtype=np.float32)
        N.check(N.lib().wc_spectrum_window(self._h, N.np_ptr(w)))
        self._window = w

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().wc_spectrum_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def name(self) -> str:
        return "cuda"

    def freqs(self, sample_rate: int) -> np.ndarray:
        """fftshift(fftfreq(N, 1/fs)) as float32 (scipy_backend.py:64,78)."""
        n = self.fft_size
        k = np.arange(-(n // 2), n - n // 2, dtype=np.float64)
        return (k / (n * (1.0 / sample_rate))).astype(np.float32)  # numpy.fft.fftfreq: k / (n*d)

    def execute(self, iq, sample_rate: int) -> FFTResult:
        """First fft_size samples -> dB spectrum; zeros when fewer samples (scipy_backend.py:48-79)."""
        n = self.fft_size
        size = int(iq.numel()) if hasattr(iq, "numel") else int(np.asarray(iq).size)
        if size < n:
            return FFTResult(np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32), sample_rate / n)
        power = self.execute_frames(iq, n_frames=1, frame_stride=n, avg=1)[0]
        return FFTResult(power_db=power, freqs=self.freqs(sample_rate), bin_hz=sample_rate / n)

    def execute_frames(self, iq, n_frames: int, frame_stride: int | None = None, avg: int = 1):
        """dB spectra of `n_frames` frames (frame f starts at f*frame_stride, only its first fft_size
        samples are used, exactly like one execute() per chunk); consecutive groups of `avg` frames
        are averaged in dB. numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
        n = self.fft_size
        stride = n if frame_stride is None else int(frame_stride)
        groups = (n_frames + avg - 1) // avg
        if N.is_torch_cuda(iq):
            import torch

            x = iq.to(torch.complex64).contiguous()
            assert x.numel() >= (n_frames - 1) * stride + n
            out = torch.empty((groups, n), dtype=torch.float32, device=x.device)
            N.check(N.lib().wc_spectrum_execute(self._h, C.c_void_p(x.data_ptr()), stride, n_frames, avg,
                                                C.c_void_p(out.data_ptr()), N.torch_stream_ptr()))
            return out
        x = np.ascontiguousarray(iq, dtype=np.complex64).reshape(-1)
        assert x.size >= (n_frames - 1) * stride + n
        out = np.empty((groups, n), dtype=np.float32)
        N.check(N.lib().wc_spectrum_execute_host(self._h, N.np_ptr(x), stride, n_frames, avg, N.np_ptr(out)))
        return out
