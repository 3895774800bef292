"""FFT backend registry with the call surface of wavecapsdr/dsp/fft/registry.py:27-176.

This package carries exactly one backend ("cuda", csrc/spectrum.cu) — no multi-backend dispatch and
no CPU fallback. `register` is kept so callers can add their own classes; asking for a backend that
is not registered raises instead of silently falling back to scipy."""
from __future__ import annotations

from typing import Any, Callable

from .base import FFTBackend

_BACKENDS: dict[str, type[FFTBackend]] = {}


def register(name: str) -> Callable[[type[FFTBackend]], type[FFTBackend]]:
    def decorator(cls: type[FFTBackend]) -> type[FFTBackend]:
        _BACKENDS[name] = cls
        return cls

    return decorator


def _ensure_registered() -> None:
    if "cuda" not in _BACKENDS:
        from .cuda_backend import CudaFFTBackend

        _BACKENDS["cuda"] = CudaFFTBackend


def get_backend(accelerator: str = "auto", fft_size: int = 2048, **kwargs: Any) -> FFTBackend:
    _ensure_registered()
    name = "cuda" if accelerator == "auto" else accelerator
    if name not in _BACKENDS:
        raise ValueError(f"FFT backend '{accelerator}' is not part of wavecap_sdr_b200 (registered: {sorted(_BACKENDS)})")
    return _BACKENDS[name](fft_size=fft_size, **kwargs)


def available_backends() -> list[str]:
    _ensure_registered()
    return sorted(_BACKENDS)
