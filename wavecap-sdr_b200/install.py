"""Rebind the reference's hot-path names to the GPU implementations.

    import wavecap_sdr_b200.install as b200
    b200.install()             # needs an importable `wavecapsdr`, libwcsdr_b200.so and a B200

Every right-hand side has the reference's signature and array contracts. `uninstall()` restores the originals.
"""
from __future__ import annotations

import importlib
import inspect
import sys

from . import _native as N

# (reference module, attribute) -> (our module, attribute)
REBIND = [
    ("wavecapsdr.dsp.channelizer", "PolyphaseChannelizer", "wavecap_sdr_b200.dsp.channelizer", "PolyphaseChannelizer"),
    ("wavecapsdr.dsp.channelizer", "channelize_samples", "wavecap_sdr_b200.dsp.channelizer", "channelize_samples"),
    *[("wavecapsdr.dsp.fm", n, "wavecap_sdr_b200.dsp.fm", n) for n in
      ("quadrature_demod", "deemphasis_filter", "lpf_audio", "resample_poly", "rms_normalize", "soft_clip", "wbfm_demod",
       "nbfm_demod")],
    *[("wavecapsdr.dsp.am", n, "wavecap_sdr_b200.dsp.am", n) for n in ("freq_shift", "am_demod", "ssb_demod")],
    *[("wavecapsdr.dsp.sam", n, "wavecap_sdr_b200.dsp.sam", n) for n in ("CarrierRecoveryPLL", "sam_demod", "sam_demod_simple")],
    ("wavecapsdr.capture", "sam_demod_simple", "wavecap_sdr_b200.dsp.sam", "sam_demod_simple"),   # bound by name at capture.py:45
    *[("wavecapsdr.dsp.agc", n, "wavecap_sdr_b200.dsp.agc", n) for n in ("apply_agc", "apply_simple_agc", "soft_clip")],
    *[("wavecapsdr.dsp.filters", n, "wavecap_sdr_b200.dsp.filters", n) for n in
      ("highpass_filter", "lowpass_filter", "bandpass_filter", "notch_filter", "fir_filter_complex", "fir_decimate")],
    *[("wavecapsdr.capture", n, "wavecap_sdr_b200.capture", n) for n in
      ("freq_shift", "decimate_iq_for_p25", "_process_channel_dsp_stateless")],
    *[("wavecapsdr.dsp.p25.c4fm", n, "wavecap_sdr_b200.dsp.p25.c4fm", n) for n in
      ("C4FMDemodulator", "c4fm_demod_simple", "_FMDemodulator", "_Interpolator", "_SoftSyncDetector")],
    ("wavecapsdr.decoders.p25", "CQPSKDemodulator", "wavecap_sdr_b200.decoders.p25", "CQPSKDemodulator"),
    # P25 framing. decoders/p25.py binds P25P1MessageFramer / P25P1Message / P25P1DataUnitID by name
    # at import time (decoders/p25.py:25-29), so they are rebound there as well as in their home module.
    ("wavecapsdr.dsp.fec.bch", "bch_decode", "wavecap_sdr_b200.dsp.fec.bch", "bch_decode"),
    ("wavecapsdr.dsp.fec.trellis", "trellis_decode", "wavecap_sdr_b200.dsp.fec.trellis", "trellis_decode"),
    ("wavecapsdr.decoders.p25", "P25TrellisDecoder", "wavecap_sdr_b200.decoders.p25", "P25TrellisDecoder"),
    ("wavecapsdr.decoders.p25_framer", "P25P1SoftSyncDetector", "wavecap_sdr_b200.decoders.p25_framer", "P25P1SoftSyncDetector"),
    ("wavecapsdr.decoders.p25_framer", "P25P1MessageFramer", "wavecap_sdr_b200.decoders.p25_framer", "P25P1MessageFramer"),
    ("wavecapsdr.decoders.p25", "P25P1MessageFramer", "wavecap_sdr_b200.decoders.p25_framer", "P25P1MessageFramer"),
    # voice-channel discriminator path
    ("wavecapsdr.decoders.p25", "DiscriminatorDemodulator", "wavecap_sdr_b200.decoders.p25", "DiscriminatorDemodulator"),
    # control-channel scanner: one library call per scan instead of one CPU chain per candidate (trunking/system.py:997)
    ("wavecapsdr.trunking.cc_scanner", "ControlChannelScanner", "wavecap_sdr_b200.cc_scanner", "ControlChannelScanner"),
    ("wavecapsdr.trunking.cc_scanner", "ChannelMeasurement", "wavecap_sdr_b200.cc_scanner", "ChannelMeasurement"),
]

_saved: list[tuple[object, str, object]] = []
_saved_backends: list[tuple[dict, str, object]] = []


def install(device: int | None = None) -> list[str]:
    """Returns the list of rebound names; raises NativeError when the library or a B200 is missing."""
    N.init(device)
    done = []
    pairs = []
    for ref_mod, ref_attr, our_mod, our_attr in REBIND:
        rm = importlib.import_module(ref_mod)
        om = importlib.import_module(our_mod)
        old, new = getattr(rm, ref_attr), getattr(om, our_attr)
        _saved.append((rm, ref_attr, old))
        setattr(rm, ref_attr, new)
        pairs.append((old, new))
        done.append(f"{ref_mod}.{ref_attr}")
    # The package binds many of these objects BY NAME at import time — `from wavecapsdr.dsp.p25.c4fm import C4FMDemodulator
    # as DSPC4FMDemodulator` (trunking/control_channel.py:29), `... as _WorkingC4FMDemodulator` (decoders/p25.py:39),
    # `from .dsp.fm import nbfm_demod, quadrature_demod, wbfm_demod` (capture.py:42), `bch_decode` / `trellis_decode` in
    # decoders/p25_frames.py:23-25 ... Those aliases still point at the CPU objects: every module of the package that is
    # already loaded is swept for them (modules imported later pick the new objects up by themselves).
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "wavecapsdr" or mod_name.startswith("wavecapsdr.")):
            continue
        for attr, val in list(vars(mod).items()):
            if not (inspect.isfunction(val) or inspect.isclass(val)):
                continue
            for old, new in pairs:
                if val is old:
                    _saved.append((mod, attr, old))
                    setattr(mod, attr, new)
                    done.append(f"{mod_name}.{attr} (alias)")
                    break
    # FFT registry: take the "cuda" slot (dsp/fft/registry.py:167-174)
    try:
        reg = importlib.import_module("wavecapsdr.dsp.fft.registry")
        from .dsp.fft.cuda_backend import CudaFFTBackend

        # the reference fills its table lazily and only while it is EMPTY (registry.py:139-142): registering "cuda" first
        # would leave "scipy" (its documented always-available fallback) unregistered for good
        ensure = getattr(reg, "_ensure_registered", None)
        if callable(ensure):
            ensure()
        backends = getattr(reg, "_BACKENDS", None)
        if isinstance(backends, dict):
            _saved_backends.append((backends, "cuda", backends.get("cuda")))
        reg.register("cuda")(CudaFFTBackend)
        done.append("wavecapsdr.dsp.fft.registry['cuda']")
    except Exception:  # registry shape differs: leave the reference's backends alone
        pass
    return done


def uninstall() -> None:
    while _saved:
        mod, attr, val = _saved.pop()
        setattr(mod, attr, val)
    while _saved_backends:
        table, name, val = _saved_backends.pop()
        if val is None:
            table.pop(name, None)
        else:
            table[name] = val
