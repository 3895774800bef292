"""Spectrum FFT backends (call surface of wavecapsdr.dsp.fft)."""
from .base import FFTBackend, FFTResult
from .registry import available_backends, get_backend, register

__all__ = ["FFTBackend", "FFTResult", "available_backends", "get_backend", "register"]
