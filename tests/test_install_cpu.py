"""CPU: the rebinding table of wavecap_sdr_b200.install names things that exist on both sides, with the same
call signature (needs /root/reference, so it only runs in the build container)."""
import importlib
import inspect

import pytest

from oracle import refenv


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_rebind_table_resolves_and_signatures_match():
    refenv.load()
    from wavecap_sdr_b200.install import REBIND

    for ref_mod, ref_attr, our_mod, our_attr in REBIND:
        r = getattr(importlib.import_module(ref_mod), ref_attr)
        o = getattr(importlib.import_module(our_mod), our_attr)
        if inspect.isclass(r):
            r, o = r.__init__, o.__init__
        rp = [p for p in inspect.signature(r).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
        op = [p for p in inspect.signature(o).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
        assert [p.name for p in rp] == [p.name for p in op][: len(rp)], (ref_mod, ref_attr)
        for a, b in zip(rp, op):
            if a.default is not inspect.Parameter.empty:
                assert a.default == b.default, (ref_mod, ref_attr, a.name)


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_install_keeps_the_reference_fft_backends_registered(monkeypatch):
    """install() adds the "cuda" slot to the reference's FFT registry WITHOUT keeping the reference's own backends out of
    it: the registry fills itself lazily and only while empty (dsp/fft/registry.py:139-142), and `get_backend()` falls back
    to `_BACKENDS["scipy"]` unconditionally (:121). Found by running the reference's own tests/unit/test_fft_backends.py
    against install(). Host logic only: the device selection is stubbed out, no kernel runs."""
    refenv.load()
    import wavecapsdr.dsp.fft.registry as reg
    import wavecap_sdr_b200._native as N
    import wavecap_sdr_b200.install as b200

    saved = dict(reg._BACKENDS)
    reg._BACKENDS.clear()                      # the state of a fresh process
    monkeypatch.setattr(N, "init", lambda device=None: None)
    try:
        names = b200.install(0)
        assert "wavecapsdr.dsp.fft.registry['cuda']" in names
        assert "scipy" in reg._BACKENDS and reg._BACKENDS["cuda"].__module__.startswith("wavecap_sdr_b200")
        assert reg.get_backend("scipy", fft_size=1024).name == "scipy"
        assert reg.get_backend("auto", fft_size=1024).name in ("scipy", "fftw")   # small sizes stay on the CPU (:87-104)
    finally:
        b200.uninstall()
    assert "scipy" in reg._BACKENDS and "cuda" not in reg._BACKENDS      # the slot is given back (no CuPy in this image)
    reg._BACKENDS.clear()
    reg._BACKENDS.update(saved)


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_rebound_classes_keep_the_reference_public_surface():
    """Every public name of a rebound reference class — methods, properties, class constants, and the attributes its
    __init__ assigns — exists on the class that replaces it, plus the private ones the reference itself reads from outside
    the class (`demod._sync_count`, `demod._equalizer.pll/.gain` in cli.py:759-760; `_ted_phase` in its tests)."""
    import re

    refenv.load()
    from wavecap_sdr_b200.install import REBIND

    for ref_mod, ref_attr, our_mod, our_attr in REBIND:
        r = getattr(importlib.import_module(ref_mod), ref_attr)
        o = getattr(importlib.import_module(our_mod), our_attr)
        if not inspect.isclass(r):
            continue
        want = {n for n in dir(r) if not n.startswith("_")}
        try:
            want |= {n for n in re.findall(r"self\.([A-Za-z][A-Za-z0-9_]*)\s*[:=][^=]", inspect.getsource(r.__init__))}
        except (OSError, TypeError):   # a dataclass: the generated __init__ has no source, its fields are the attributes
            want |= {f for f in getattr(r, "__dataclass_fields__", {}) if not f.startswith("_")}
        have = set(dir(o)) | set(re.findall(r"self\.([A-Za-z_][A-Za-z0-9_]*)", inspect.getsource(o)))
        have |= set(getattr(o, "__dataclass_fields__", {}))
        assert not sorted(want - have), (ref_mod, ref_attr, sorted(want - have))
    import wavecap_sdr_b200.dsp.p25.c4fm as c4

    for name in ("_sync_count", "_fine_sync", "_ted_phase", "_equalizer", "_sample_point"):
        assert hasattr(c4.C4FMDemodulator, name), name


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_install_reaches_the_by_name_aliases_of_the_package(monkeypatch):
    """The reference binds the demodulators and DSP functions by name in the modules that use them
    (trunking/control_channel.py:28-29, decoders/p25.py:37-39, decoders/p25_frames.py:23-25, capture.py:38-45, ...): install()
    must replace those aliases too, or the trunking control-channel monitor keeps building the CPU demodulators. Host logic
    only (device selection stubbed)."""
    refenv.load()
    import wavecapsdr.capture as rcap
    import wavecapsdr.decoders.p25 as rp25
    import wavecapsdr.decoders.p25_frames as rfr
    import wavecapsdr.trunking.control_channel as rcc
    import wavecap_sdr_b200._native as N
    import wavecap_sdr_b200.install as b200

    before = (rcc.DSPC4FMDemodulator, rcc.P25CQPSKDemodulator, rp25._WorkingC4FMDemodulator, rfr.bch_decode, rfr.trellis_decode,
              rcap.wbfm_demod, rcap.nbfm_demod, rcap.am_demod)
    assert all(o.__module__.startswith("wavecapsdr") for o in before)
    monkeypatch.setattr(N, "init", lambda device=None: None)
    try:
        names = b200.install(0)
        after = (rcc.DSPC4FMDemodulator, rcc.P25CQPSKDemodulator, rp25._WorkingC4FMDemodulator, rfr.bch_decode, rfr.trellis_decode,
                 rcap.wbfm_demod, rcap.nbfm_demod, rcap.am_demod)
        assert all(o.__module__.startswith("wavecap_sdr_b200") for o in after), [o.__module__ for o in after]
        assert "wavecapsdr.trunking.control_channel.DSPC4FMDemodulator (alias)" in names
    finally:
        b200.uninstall()
    restored = (rcc.DSPC4FMDemodulator, rcc.P25CQPSKDemodulator, rp25._WorkingC4FMDemodulator, rfr.bch_decode, rfr.trellis_decode,
                rcap.wbfm_demod, rcap.nbfm_demod, rcap.am_demod)
    assert all(a is b for a, b in zip(before, restored))
