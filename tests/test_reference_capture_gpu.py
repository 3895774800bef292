"""GPU: the reference's OWN per-capture caller — `Capture._process_channels_parallel` (capture.py:2489-2597), which hands
every running channel of a capture to a 3-worker ThreadPoolExecutor (`_get_dsp_executor`, :1906-1925) — run unmodified, first
with the reference's CPU `_process_channel_dsp_stateless`, then with the function `install()` binds in its place. Same Capture
object, same Channel objects, same chunks; the audio of every channel must agree to the north star's 1e-4 relative RMS.
This is the call path a deployed WaveCap-SDR takes per 50 ms chunk, worker threads included (each thread selects the device
on its first call). The reference travels in oracle/_ref/reference_backend.tar; without it the test is skipped."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import parity_note, rel_rms
from oracle import build_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref.staged(), reason="oracle/_ref not staged")]

def _carriers(fs, n, n_chunks, plan, seed):
    """one carrier per channel on a noise floor (FM: sine-modulated phase; AM / SSB: 50 % sine envelope), fresh per chunk"""
    rng = np.random.default_rng(seed)
    t = np.arange(n * n_chunks) / fs
    x = 0.02 * (rng.standard_normal(t.size) + 1j * rng.standard_normal(t.size))
    for i, (mode, off) in enumerate(plan):
        tone = 600.0 + 250.0 * i
        if mode in ("wbfm", "nbfm"):
            dev = 75_000.0 if mode == "wbfm" else 5_000.0
            x += 0.3 * np.exp(1j * (2 * np.pi * off * t + (dev / tone) * np.sin(2 * np.pi * tone * t)))
        else:
            x += 0.3 * (1.0 + 0.5 * np.sin(2 * np.pi * tone * t)) * np.exp(1j * 2 * np.pi * off * t)
    return x.astype(np.complex64).reshape(n_chunks, n)


# the capture loop's chunk is max(8192, fs // 20) samples (capture.py:3035). AM / SSB run at 48 kS/s: at MS/s rates the
# reference's own tf-form 100 Hz high-pass moves by more than 1e-4 under a 1-ulp input change
CASES = [
    ("fm 2.4 MS/s", 2_400_000, [("wbfm", 200_000.0), ("nbfm", -350_000.0), ("nbfm", 612_500.0), ("nbfm", -800_000.0)]),
    ("am/ssb 48 kS/s", 48_000, [("am", 6_000.0), ("ssb", -9_000.0), ("nbfm", 15_000.0), ("am", 0.0)]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_capture_parallel_dsp_matches_the_reference_through_its_own_thread_pool(native, case):
    label, fs, plan = case
    n = max(8192, fs // 20)
    build_ref.load()
    import wavecapsdr.capture as rc
    from wavecapsdr.devices.fake import FakeDriver
    import wavecap_sdr_b200.install as b200

    cap = rc.Capture(cfg=rc.CaptureConfig(id="c1", device_id="fake0", center_hz=100e6, sample_rate=fs), driver=FakeDriver())
    for i, (mode, off) in enumerate(plan):
        ch = rc.Channel(rc.ChannelConfig(id=f"ch{i}", capture_id="c1", mode=mode, offset_hz=off))
        ch.start()
        cap._channels[ch.cfg.id] = ch
    chunks = _carriers(fs, n, 3, plan, seed=5)

    def run():
        out = []
        with ThreadPoolExecutor(max_workers=3, thread_name_prefix="DSP-c1-") as ex:   # _get_dsp_executor's shape
            for k in range(chunks.shape[0]):
                res = cap._process_channels_parallel(chunks[k], ex, timeout=120.0)
                assert len(res) == len(plan)
                out.append({ch.cfg.id: audio for ch, audio in res})
        return out

    ref = run()                                   # the reference's own numpy/scipy path on the box's CPU
    names = b200.install(0)
    try:
        assert "wavecapsdr.capture._process_channel_dsp_stateless" in names
        assert rc._process_channel_dsp_stateless.__module__.startswith("wavecap_sdr_b200")
        run()                                     # first calls: plans, filter tables, per-thread device selection
        got = run()
    finally:
        b200.uninstall()
    worst = 0.0
    for k in range(len(ref)):
        for i, (mode, _) in enumerate(plan):
            a, b = ref[k][f"ch{i}"], got[k][f"ch{i}"]
            assert a is not None and b is not None, (k, mode)
            a, b = np.asarray(a), np.asarray(b)
            assert a.dtype == b.dtype == np.float32 and a.shape == b.shape, (k, mode, a.shape, b.shape)
            e = rel_rms(b, a)
            worst = max(worst, e)
            assert e < 1e-4, (label, k, mode, e)
    parity_note(f"reference Capture._process_channels_parallel, {label} (3 worker threads, {'+'.join(m for m, _ in plan)}, "
                f"3 chunks of {n} samples): untouched vs install(), worst audio rel-RMS {worst:.1e}")


@pytest.mark.parametrize("fft_size", [2048, 65536])
def test_capture_calculate_fft_through_the_registry(native, fft_size):
    """The spectrum producer of the reference's capture loop, `Capture._calculate_fft` -> `get_backend(accelerator, fft_size)`
    -> `FFTBackend.execute` (capture.py:2353-2404), unmodified: `fft_accelerator="scipy"` on the untouched reference against
    `"cuda"` (and `"auto"`, which picks cuda above 4096 points, registry.py:92-104) after install() registered its backend."""
    build_ref.load()
    import wavecapsdr.capture as rc
    from wavecapsdr.devices.fake import FakeDriver
    import wavecap_sdr_b200.install as b200

    fs = 61_440_000
    rng = np.random.default_rng(9)
    n = max(8192, fft_size) * 2
    t = np.arange(n) / fs
    x = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) + 0.5 * np.exp(2j * np.pi * 7.3e6 * t)
         + 0.01 * np.exp(-2j * np.pi * 19.1e6 * t)).astype(np.complex64)

    def spectrum(accelerator):
        cap = rc.Capture(cfg=rc.CaptureConfig(id="c1", device_id="fake0", center_hz=100e6, sample_rate=fs,
                                              fft_accelerator=accelerator), driver=FakeDriver())
        cap._calculate_fft(x, fs, fft_size)
        return cap._fft_backend.name, np.asarray(cap._fft_power), np.asarray(cap._fft_freqs)

    name0, p0, f0 = spectrum("scipy")
    assert name0 == "scipy"
    b200.install(0)
    try:
        name1, p1, f1 = spectrum("cuda")
        name2, p2, _ = spectrum("auto")
        import wavecapsdr.dsp.fft as rfft

        assert "scipy" in rfft.available_backends() and "cuda" in rfft.available_backends()
    finally:
        b200.uninstall()
    assert name1 == "cuda" and (name2 == "cuda") == (fft_size > 4096), (name1, name2)
    assert p1.shape == p0.shape == (fft_size,) and p1.dtype == p0.dtype
    assert np.array_equal(f1, f0)
    e = rel_rms(p1, p0)
    assert e < 1e-4 and rel_rms(p2, p0) < 1e-4, e
    assert int(np.argmax(p1)) == int(np.argmax(p0))
    parity_note(f"reference Capture._calculate_fft, {fft_size} points: scipy backend vs the installed cuda backend, dB rel-RMS {e:.1e}")


@pytest.mark.parametrize("fs,mode,offset", [(2_400_000, "wbfm", 200_000.0), (2_400_000, "nbfm", -350_000.0),
                                            (48_000, "am", 6_000.0), (48_000, "ssb", -9_000.0)])
def test_channel_process_iq_chunk_after_install(native, fs, mode, offset):
    """`Channel.process_iq_chunk` (capture.py:950-1200), the per-channel entry of the reference that calls `freq_shift`,
    `wbfm_demod` / `nbfm_demod` / `am_demod` / `ssb_demod` through the names capture.py bound at import time (:38-45):
    install() replaces those aliases, so this unmodified coroutine runs the CUDA chain. Audio handed to `_broadcast` and
    the channel's `signal_power_db` before / after."""
    import asyncio

    build_ref.load()
    import wavecapsdr.capture as rc
    import wavecap_sdr_b200.install as b200

    n = max(8192, fs // 20)
    chunks = _carriers(fs, n, 2, [(mode, offset)], seed=11)

    def run():
        ch = rc.Channel(rc.ChannelConfig(id="ch0", capture_id="c1", mode=mode, offset_hz=offset))
        ch.start()
        got = []

        async def keep(audio):
            got.append(np.array(audio, copy=True))

        ch._broadcast = keep

        async def go():
            for k in range(chunks.shape[0]):
                await ch.process_iq_chunk(chunks[k], fs)

        asyncio.run(go())
        return got, ch.signal_power_db

    ref, p0 = run()
    assert len(ref) == 2 and rc.wbfm_demod.__module__.startswith("wavecapsdr")
    b200.install(0)
    try:
        assert rc.wbfm_demod.__module__.startswith("wavecap_sdr_b200") and rc.am_demod.__module__.startswith("wavecap_sdr_b200")
        run()
        got, p1 = run()
    finally:
        b200.uninstall()
    assert len(got) == len(ref)
    worst = max(rel_rms(b, a) for a, b in zip(ref, got))
    assert all(a.dtype == b.dtype and a.shape == b.shape for a, b in zip(ref, got))
    assert worst < 1e-4 and abs(p1 - p0) < 1e-3, (mode, worst, p0, p1)
    parity_note(f"reference Channel.process_iq_chunk ({mode}, {fs} S/s): audio before / after install() rel-RMS {worst:.1e}")
