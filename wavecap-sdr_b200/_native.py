"""ctypes binding of libwcsdr_b200.so (the C-ABI declared in include/wcsdr_b200.h).

There is NO CPU fallback: if the shared library is missing or no B200 is present, every compute
entry point raises. Only `tests/`, `bench.py --impl reference` and `smoke()` may touch `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libwcsdr_b200.so"

_lib = None
_lock = threading.Lock()
_inited_device: int | None = None


class NativeError(RuntimeError):
    """Raised when the CUDA library reports an error (or is not built)."""


def lib() -> C.CDLL:
    """Load the shared library (once). Raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "wavecap_sdr_b200 has no CPU fallback."
            )
        import torch  # noqa: F401  (loads the process's libcudart first; the library links the shared runtime)

        l = C.CDLL(str(LIB_PATH))
        _declare(l)
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().wc_last_error().decode("utf-8", "replace")
        raise NativeError(f"wcsdr_b200 error {rc}: {msg}")


def check_nonneg(rc: int) -> int:
    """for entry points that return an index (>= 0) or a negative error code"""
    if rc < 0:
        msg = lib().wc_last_error().decode("utf-8", "replace")
        raise NativeError(f"wcsdr_b200 error {rc}: {msg}")
    return rc


def init(device: int | None = None) -> None:
    """Select the CUDA device (default: LOCAL_RANK or 0) and verify it is sm_100."""
    global _inited_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device == device and getattr(_tls, "device", None) == device:
        return
    check(lib().wc_init(device))
    _inited_device = device
    _tls.device = device


_tls = threading.local()


def ensure_init() -> None:
    """The CUDA current device is per host thread: the reference's DSP pool threads (capture.py:1906-1925) never select
    one, so every thread that reaches the library binds itself to the device init() chose."""
    if _inited_device is None:
        init()
    elif getattr(_tls, "device", None) != _inited_device:
        import torch

        torch.cuda.set_device(_inited_device)      # cudaSetDevice for this thread (torch and the C ABI share the runtime)
        _tls.device = _inited_device


# ---- pointer helpers -------------------------------------------------------------------------------

def np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(getattr(x, "is_cuda"))


def torch_stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by CUDA pinned host memory (torch is only the allocator)."""
    import torch

    tdt = {np.dtype(np.complex64): torch.complex64, np.dtype(np.float32): torch.float32,
           np.dtype(np.int16): torch.int16, np.dtype(np.uint8): torch.uint8,
           np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    t = torch.empty(shape, dtype=tdt, pin_memory=True)
    a = t.numpy()
    _keepalive[id(a)] = t
    return a


_keepalive: dict[int, object] = {}


# ---- prototypes ------------------------------------------------------------------------------------

def _declare(l: C.CDLL) -> None:
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
    P = C.POINTER

    def fn(name, res, *args):
        f = getattr(l, name)
        f.restype = res
        f.argtypes = list(args)

    fn("wc_init", i32, i32)
    fn("wc_last_error", C.c_char_p)
    fn("wc_version", C.c_char_p)
    fn("wc_device_info", i32, P(i32), P(i32), P(i32), P(i64))
    # channelizer
    fn("wc_chan_create", i32, f64, i32, i32, P(vp))
    fn("wc_chan_destroy", None, vp)
    fn("wc_chan_info", i32, vp, P(i32), P(f64), P(i32))
    fn("wc_chan_get_arms", i32, vp, vp)
    fn("wc_chan_get_history", i32, vp, vp)
    fn("wc_chan_frames_for", i64, vp, i64)
    fn("wc_chan_reset", i32, vp)
    fn("wc_chan_process", i32, vp, vp, i64, i32, i64, i32, f32, vp, vp)
    fn("wc_chan_carry_from", i32, vp, vp, i64, vp)
    fn("wc_chan_carry_tail", i32, vp, vp, vp)
    fn("wc_chan_set_run_frames", i32, vp, i32)
    fn("wc_chan_process_host", i32, vp, vp, i64, i32, i32, f32, vp)
    fn("wc_chan_audio_config", i32, vp, i32, i32)
    fn("wc_chan_audio_len", i64, vp, i64)
    fn("wc_chan_process_ex", i32, vp, vp, i32, i64, i32, i64, i32, f32, vp, vp)
    fn("wc_chan_process_host_ex", i32, vp, vp, i32, i64, i32, i32, f32, vp)
    # analog chain stages
    fn("wc_front_chan_scratch_bytes", i32, i32)
    fn("wc_front_run", i32, vp, i32, i32, i32, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp)
    fn("wc_front_run_ex", i32, vp, i32, i32, i32, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp)
    fn("wc_iir_create", i32, vp, i32, vp, i32, P(vp))
    fn("wc_iir_destroy", None, vp)
    fn("wc_iir_is_sequential", i32, vp)
    fn("wc_iir_kind", i32, vp)
    fn("wc_iir_lfilter", i32, vp, vp, vp, i32, i64, i32, i32, vp)
    fn("wc_sumsq", i32, vp, i32, i64, i32, vp, vp)
    fn("wc_elementwise", i32, vp, vp, i64, i32, f32, vp)
    fn("wc_agc_apply", i32, vp, vp, vp, vp, i64, f32, f32, vp)
    fn("wc_resampler_create", i32, i32, i32, vp, i32, P(vp))
    fn("wc_resampler_destroy", None, vp)
    fn("wc_resampler_out_len", i64, vp, i64)
    fn("wc_resampler_run", i32, vp, vp, i32, i64, i32, vp, i32, vp, f32, f32, vp, vp, f32, vp)
    fn("wc_finalize", i32, vp, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp)
    fn("wc_sam_pll", i32, vp, i64, i32, i32, f64, f64, i32, i32, vp, vp, vp, vp, vp)
    # analog plan
    fn("wc_analog_plan_create", i32, i32, i32, i32, i32, vp, vp, vp, vp, P(vp))
    fn("wc_analog_plan_add_run", i32, vp, i32, i32, i32, i32, i32, vp, i32)
    fn("wc_analog_plan_add_iir", i32, vp, i32, vp, i32, vp, i32)
    fn("wc_analog_plan_set_agc", i32, vp, i32, vp, vp, vp, vp, f32, f32)
    fn("wc_analog_plan_finish", i32, vp)
    fn("wc_analog_plan_destroy", None, vp)
    fn("wc_analog_plan_audio_floats", i64, vp)
    fn("wc_analog_plan_audio_len", i32, vp, i32)
    fn("wc_analog_plan_audio_offset", i64, vp, i32)
    fn("wc_analog_plan_use_graph", i32, vp, i32)
    fn("wc_analog_run", i32, vp, vp, i32, vp, vp, vp)
    # spectrum
    fn("wc_spectrum_create", i32, i32, P(vp))
    fn("wc_spectrum_destroy", None, vp)
    fn("wc_spectrum_window", i32, vp, vp)
    fn("wc_spectrum_execute", i32, vp, vp, i64, i32, i32, vp, vp)
    fn("wc_spectrum_execute_host", i32, vp, vp, i64, i32, i32, vp)
    # P25 C4FM
    u8p = vp
    fn("wc_c4fm_create", i32, i32, i32, i32, i32, vp, i32, vp, i32, P(vp))
    fn("wc_c4fm_destroy", None, vp)
    fn("wc_c4fm_info", i32, vp, P(i32), P(f64), P(i32), P(i32))
    fn("wc_c4fm_get_taps", i32, vp, vp, vp)
    fn("wc_c4fm_max_symbols", i32, vp, i32)
    fn("wc_c4fm_reset", i32, vp, i32)
    fn("wc_c4fm_demod", i32, vp, vp, i64, i32, u8p, vp, vp, i32, vp)
    fn("wc_c4fm_demod_host", i32, vp, vp, i32, u8p, vp, vp, i32)
    fn("wc_c4fm_get_state", i32, vp, i32, vp)
    fn("wc_c4fm_demod_disc", i32, vp, vp, i64, i32, vp, u8p, vp, vp, i32, vp)
    fn("wc_c4fm_demod_disc_host", i32, vp, vp, i32, vp, u8p, vp, vp, i32)
    fn("wc_c4fm_diffdemod", i32, vp, vp, i32, vp, vp)
    fn("wc_c4fm_interp", i32, vp, i32, vp, vp, i32, vp, vp)
    fn("wc_c4fm_sync_scores", i32, vp, i32, vp, vp, vp, vp)
    # P25 CQPSK
    fn("wc_cqpsk_create", i32, i32, i32, i32, vp, vp, P(vp))
    fn("wc_cqpsk_destroy", None, vp)
    fn("wc_cqpsk_max_symbols", i32, vp, i32)
    fn("wc_cqpsk_reset", i32, vp, i32)
    fn("wc_cqpsk_demod", i32, vp, vp, i64, i32, vp, vp, i32, vp)
    fn("wc_cqpsk_demod_host", i32, vp, vp, i32, vp, vp, i32)
    fn("wc_cqpsk_get_state", i32, vp, i32, vp)
    # streaming FIR / trunking fan-out
    fn("wc_fir_complex", i32, vp, i32, vp, i32, i32, vp, vp, vp, vp)
    fn("wc_ddc_create", i32, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, P(vp))
    fn("wc_ddc_destroy", None, vp)
    fn("wc_ddc_get_taps", i32, vp, vp, vp, P(i32), P(i32))
    fn("wc_ddc_set_offsets", i32, vp, vp)
    fn("wc_ddc_reset", i32, vp, i32)
    fn("wc_ddc_out_len", i32, vp, i32)
    fn("wc_ddc_process", i32, vp, vp, i32, vp, i64, vp)
    fn("wc_ddc_process_host", i32, vp, vp, i32, vp)
    # optional audio clean-up stages
    fn("wc_noise_blanker", i32, vp, vp, i32, i64, i32, f32, i32, vp)
    fn("wc_spectral_nr_out_len", i32, i32)
    fn("wc_spectral_nr", i32, vp, i32, i64, i32, f32, vp, i64, vp)
    fn("wc_pack", i32, vp, vp, i64, i32, vp)
    fn("wc_audio_levels", i32, vp, i32, i64, i32, vp, vp, vp, vp)
    fn("wc_signal_metrics", i32, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp)
    # P25 framing
    fn("wc_bch_decode", i32, vp, vp, i32, vp, vp, vp)
    fn("wc_bch_decode_host", i32, vp, vp, i32, vp, vp)
    fn("wc_trellis12_decode", i32, vp, i64, vp, i32, vp, i32, vp, i64, vp, vp, vp)
    fn("wc_tsbk_decode", i32, vp, i32, vp, vp, vp, vp, vp)
    fn("wc_tsbk_decode_host", i32, vp, i32, vp, vp, vp, vp)
    fn("wc_p25framer_create", i32, i32, P(vp))
    fn("wc_p25framer_destroy", None, vp)
    fn("wc_p25framer_reset", i32, vp, i32, i32)
    fn("wc_p25framer_max_msgs", i32, i32)
    fn("wc_p25framer_pool_bytes", i32, i32)
    fn("wc_p25framer_process", i32, vp, vp, vp, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp)
    fn("wc_p25framer_process_host", i32, vp, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp)
    fn("wc_p25framer_get_state", i32, vp, i32, vp)
    # control-channel scanner
    fn("wc_ccscan_out_len", i32, i32, i32)
    fn("wc_ccscan_measure", i32, vp, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp)
    # voice-channel discriminator path
    fn("wc_fm_discriminator", i32, vp, i32, i64, i32, i32, vp, vp, vp)
    fn("wc_discdemod_create", i32, i32, i32, i32, vp, vp, P(vp))
    fn("wc_discdemod_destroy", None, vp)
    fn("wc_discdemod_reset", i32, vp, i32)
    fn("wc_discdemod_max_symbols", i32, vp, i32)
    fn("wc_discdemod_get_taps", i32, vp, vp)
    fn("wc_discdemod_demod", i32, vp, vp, i64, i32, vp, vp, vp, i32, vp)
    fn("wc_discdemod_demod_host", i32, vp, vp, i32, vp, vp, vp, i32)
    fn("wc_discdemod_get_state", i32, vp, i32, vp)
    # peer memory (one capture, many GPUs)
    fn("wc_peer_alloc", i32, i64, P(vp), vp)
    fn("wc_peer_free", i32, vp)
    fn("wc_peer_open", i32, vp, P(vp))
    fn("wc_peer_close", i32, vp)
    fn("wc_peer_copy", i32, vp, vp, i64, vp)
    fn("wc_flag_set", i32, vp, C.c_uint, vp)
    fn("wc_flag_wait", i32, vp, i32, i64, C.c_uint, i32, vp, vp)
    for extra in _EXTRA_DECLS:
        extra(l, fn)


_EXTRA_DECLS: list = []


def exported_symbols() -> list[str]:
    """Every `wc_*` function declared in include/wcsdr_b200.h (parsed from the header)."""
    import re

    hdr = (_PKG.parent / "include" / "wcsdr_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(wc_[a-z0-9_]+)\s*\(", hdr)))
