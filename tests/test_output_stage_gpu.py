"""GPU parity: wire packers, audio level metering and the framer's soft sync detector (row 1)
vs numpy restatements of capture.py:102-144, :633-661 and decoders/p25_framer.py:125-231."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_wire_packers(native):
    from wavecap_sdr_b200.capture import pack_f32, pack_iq16, pack_pcm16

    rng = np.random.default_rng(4)
    a = (rng.standard_normal(10007) * 0.7).astype(np.float32)
    a[:4] = [1.0, -1.0, 1.5, -2.0]
    exp16 = (np.clip(a, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16).tobytes()
    assert pack_pcm16(a) == exp16
    assert pack_f32(a) == np.clip(a, -1.0, 1.0).astype(np.float32).tobytes()
    z = ((rng.standard_normal(5001) + 1j * rng.standard_normal(5001)) * 0.6).astype(np.complex64)
    zi = z.copy().view(np.float32)
    assert pack_iq16(z) == (np.clip(zi, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16).tobytes()
    assert pack_pcm16(np.zeros(0, np.float32)) == b"" and pack_iq16(np.zeros(0, np.complex64)) == b""


def test_audio_levels(native):
    from wavecap_sdr_b200.capture import audio_levels

    rng = np.random.default_rng(5)
    x = (rng.standard_normal((7, 2400)) * 0.4).astype(np.float32)
    x[3] = 0.0
    rms_db, peak_db, clips = audio_levels(x)
    for i in range(7):
        rms = float(np.sqrt(np.mean(x[i] ** 2)))
        peak = float(np.max(np.abs(x[i])))
        assert abs(rms_db[i] - (20 * np.log10(rms) if rms > 1e-10 else -100.0)) < 1e-4
        assert abs(peak_db[i] - (20 * np.log10(peak) if peak > 1e-10 else -100.0)) < 1e-6
        assert clips[i] == int(np.sum(np.abs(x[i]) > 0.95))


def test_framer_soft_sync_detector(native):
    from wavecap_sdr_b200.decoders.p25_framer import P25P1SoftSyncDetector

    det = P25P1SoftSyncDetector()
    pat = det.SYNC_PATTERN_SYMBOLS
    rng = np.random.default_rng(6)
    s = np.concatenate([rng.standard_normal(40) * 2, pat, rng.standard_normal(30) * 2]).astype(np.float32)
    ext = np.concatenate([np.zeros(24, np.float32), s])
    exp = np.correlate(ext, pat, mode="valid")[-len(s):]
    got = np.concatenate([det.process_batch(s[:10]), det.process_batch(s[10:11]), det.process_batch(s[11:])])
    assert got.dtype == np.float32 and np.max(np.abs(got - exp)) < 1e-3
    assert int(np.argmax(got)) == 40 + 23 and abs(got[63] - 216.0) < 1e-3      # perfect sync word scores 24 * 9
    det.reset()
    assert abs(det.process(3.0) - float(pat[23] * 3.0)) < 1e-5


def test_signal_metrics_vs_reference_formulas(native):
    """Channel.update_signal_metrics (capture.py:749-798): RSSI and the np.partition SNR estimate for 6 channels of one
    chunk (cf32 and cs16), vs the same numpy formulas on the oracle's freq_shift."""
    from oracle.analog import freq_shift as o_shift
    from wavecap_sdr_b200.capture import SignalMeter, signal_metrics

    rng = np.random.default_rng(8)
    fs, n = 2_400_000, 120_000
    t = np.arange(n) / fs
    x = (0.2 * np.exp(2j * np.pi * 300_000.0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 700 * t))
         + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    offs = [0.0, 300_000.0, -450_000.5, 12_345.0, 1_000_000.0, 0.4]
    rssi, snr = signal_metrics(x, fs, offs)
    for c, off in enumerate(offs):
        sh = x if off == 0.0 else o_shift(x, off, fs)
        mag = np.abs(sh)
        ref_rssi = float(10.0 * np.log10(np.mean(mag ** 2) + 1e-10))
        part = np.partition(mag, [n // 10, n - n // 10 - 1])
        ref_snr = float(10.0 * np.log10(part[n - n // 10 - 1] ** 2 / part[n // 10] ** 2))
        assert abs(rssi[c] - ref_rssi) <= 1e-4, (c, rssi[c], ref_rssi)
        assert abs(snr[c] - ref_snr) <= 2e-4, (c, snr[c], ref_snr)
    # int16 input, too-short input, throttle
    q = np.stack([np.round(x.real * 20000), np.round(x.imag * 20000)], axis=-1).astype(np.int16)
    r16, _ = signal_metrics(q, fs, [0.0], in_fmt="cs16")
    xq = (q[:, 0].astype(np.float32) / 32768.0 + 1j * (q[:, 1].astype(np.float32) / 32768.0)).astype(np.complex64)
    assert abs(r16[0] - float(10.0 * np.log10(np.mean(np.abs(xq) ** 2) + 1e-10))) <= 1e-4
    assert signal_metrics(x[:15], fs, [0.0])[1][0] is not None   # k_noise == 1, k_signal == 13: usable (capture.py:783)
    assert signal_metrics(x[:5], fs, [0.0])[1] == [None]           # k_noise == 0: the reference leaves snr_db None
    m = SignalMeter(offs)
    for i in range(10):
        m.update(x, fs)
        assert (m.snr_db[0] is None) == (i < 9)
    assert abs(m.rssi_db[1] - rssi[1]) < 1e-9
