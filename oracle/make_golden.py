"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:  python oracle/make_golden.py [name ...]
The fixtures are small on purpose (they are committed); full-size parity uses the oracle
restatements, which tests/test_oracle_*.py pin against these fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refenv  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cnoise(rng, n, scale=0.5):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def gen_channelizer_c5():
    from wavecapsdr.dsp.channelizer import PolyphaseChannelizer
    from wavecapsdr.dsp.fm import quadrature_demod

    rng = np.random.default_rng(5)
    fs, bw = 125_000_000.0, 488281
    n, cut = 256 + 128 * 40 + 50, 256 + 128 * 17 + 3
    x = cnoise(rng, n)
    # an FM carrier on bin 12 so the discriminator output is not just noise
    t = np.arange(n)
    x += (0.4 * np.exp(1j * (2 * np.pi * 12 * 488281.25 / fs * t + 8.0 * np.sin(2 * np.pi * 3e3 / fs * t)))).astype(np.complex64)
    ch = PolyphaseChannelizer(fs, bw)
    f1 = np.array(ch.process(x[:cut]))
    f2 = np.array(ch.process(x[cut:]))
    rate = int(ch.channel_sample_rate)
    fm2 = np.stack([quadrature_demod(ch.extract_channel(list(f2), k), rate) for k in range(256)], axis=1)
    np.savez_compressed(os.path.join(OUT, "channelizer_c5.npz"), fs=fs, bw=bw, cut=cut, x=x, frames1=f1,
                        frames2=f2, fm2=fm2.astype(np.float32), arm_history=ch.arm_history, arms=ch.arms,
                        demod_rate=rate)


def channel_calc_cases():
    """(center, fs, bw, targets) grids for ChannelCalculator: even and odd raw counts, offsets on / between bin centres,
    beyond +-Nyquist and beyond -channel_count (which the reference does not wrap)."""
    cases = []
    for center, fs, bw in ((851.0e6, 6.0e6, 12500), (460.0e6, 1.0e6, 25000), (100.0e6, 125.0e6, 488281), (155.0e6, 2.4e6, 7000)):
        count = int(fs / bw)
        steps = np.concatenate([np.arange(-count - 3, count + 4, max(1, count // 37)), [0, 1, -1, count // 2, -count // 2]])
        frac = np.array([0.0, 0.49, -0.49, 0.5, -0.5, 0.51])
        targets = (center + (steps[:, None] + frac[None, :]) * bw).reshape(-1)
        cases.append((center, fs, bw, targets))
    return cases


def channelize_samples_input():
    rng = np.random.default_rng(61)
    fs, bw, center = 1.0e6, 25000, 460.0e6
    n = 40 * 9 + 20 * 57 + 11
    t = np.arange(n)
    x = cnoise(rng, n, 0.05) + (0.5 * np.exp(2j * np.pi * (-75000.0 / fs) * t)).astype(np.complex64)
    return x.astype(np.complex64), fs, bw, center


def gen_channel_calc():
    """ChannelCalculator / channelize_samples (dsp/channelizer.py:161-268)."""
    from wavecapsdr.dsp.channelizer import ChannelCalculator, channelize_samples

    out = {}
    for i, (center, fs, bw, targets) in enumerate(channel_calc_cases()):
        calc = ChannelCalculator(center, fs, bw)
        out[f"count{i}"] = np.int64(calc.channel_count)
        out[f"index{i}"] = np.array([calc.get_channel_index(float(f)) for f in targets], dtype=np.int64)
        out[f"center{i}"] = np.array([calc.get_channel_center_frequency(k) for k in range(calc.channel_count)], dtype=np.float64)
    x, fs, bw, center = channelize_samples_input()
    for j, target in enumerate((center - 75000.0, center + 200000.0, center)):
        y, rate = channelize_samples(x, fs, target, center, bw)
        out[f"chan{j}"] = np.asarray(y)
        out[f"rate{j}"] = np.float64(rate)
    np.savez_compressed(os.path.join(OUT, "channel_calc.npz"), **out)


def gen_analog():
    """C1 (WBFM, 2.4 MS/s cf32), C2 (16 NBFM from one 10 MS/s int16 capture) and 48 kS/s AM/SSB/AGC
    through the reference's own capture._process_channel_dsp_stateless / dsp functions. Inputs are
    regenerated from seeds by oracle.analog.synth_*; only outputs are stored."""
    import wavecapsdr.capture as rc
    from wavecapsdr.dsp import agc as ragc
    from wavecapsdr.dsp import am as ram
    from wavecapsdr.dsp import fm as rfm
    from oracle import analog as oa

    out = {}
    # C1: two consecutive chunks (per-chunk restarts)
    for i in range(2):
        x = oa.synth_c1(seed=1, n=120_000, t0=i * 120_000)
        cfg = rc.ChannelConfig(id="c1", capture_id="c", mode="wbfm", offset_hz=200000.0)
        a, m = rc._process_channel_dsp_stateless(x, 2_400_000, cfg)
        out[f"c1_audio{i}"] = a
        out[f"c1_metrics{i}"] = np.array([m["rssi_db"], m["signal_power_db"]])
    # C2: one chunk, 16 channels, carriers 3 and 12 keyed off
    q, offs = oa.synth_c2(seed=2, n=500_000, keyed_off=(3, 12))
    xc = oa.cs16_to_cf32(q)
    aud, met = [], []
    for off in offs:
        cfg = rc.ChannelConfig(id="c2", capture_id="c", mode="nbfm", offset_hz=float(off))
        rc.CaptureManager._apply_mode_defaults(None, "nbfm", cfg)
        a, m = rc._process_channel_dsp_stateless(xc, 10_000_000, cfg)
        aud.append(a)
        met.append([m["rssi_db"], m["signal_power_db"]])
    out["c2_audio"] = np.stack(aud)
    out["c2_metrics"] = np.array(met)
    out["c2_offsets"] = np.array(offs, dtype=np.float64)
    # stage functions at 48 kS/s (AM/SSB/AGC are unstable at MS/s rates)
    rng = np.random.default_rng(3)
    n, fs = 9600, 48000
    t = np.arange(n) / fs
    xa = ((1 + 0.5 * np.sin(2 * np.pi * 700 * t)) * 0.3 * np.exp(1j * 0.3)
          + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    out["am_x"] = xa
    out["am_audio"] = ram.am_demod(xa, fs, 16000)
    out["am_audio_noagc"] = ram.am_demod(xa, fs, 16000, enable_agc=False)
    out["ssb_audio"] = ram.ssb_demod(xa, fs, 16000)
    out["ssb_audio_lsb"] = ram.ssb_demod(xa, fs, 16000, mode="lsb", enable_agc=False)
    f = np.real(xa).astype(np.float32)
    out["agc"] = ragc.apply_agc(f, fs)
    out["deemph"] = rfm.deemphasis_filter(f, fs)
    out["lpf"] = rfm.lpf_audio(f, fs, 5000)
    out["resamp_3_10"] = rfm.resample_poly(f, 48000, 14400)
    out["resamp_up"] = rfm.resample_poly(f[:1000], 8000, 48000)
    out["quad"] = rfm.quadrature_demod(xa, fs)
    out["fshift"] = rc.freq_shift(xa, 1234.4, fs)
    out["am_fshift"] = ram.freq_shift(xa, 1500.0, fs)
    np.savez_compressed(os.path.join(OUT, "analog.npz"), **out)


def gen_spectrum():
    """ScipyFFTBackend.execute (the reference's default backend) on seeded C3 input; 65536-point
    frames are stored every 16th bin (plus the argmax neighbourhood) to keep the fixture small."""
    from wavecapsdr.dsp.fft.scipy_backend import ScipyFFTBackend
    from oracle import spectrum as osp

    out = {}
    fs = 61_440_000
    be = ScipyFFTBackend(65536)
    frames = []
    for i in range(4):
        r = be.execute(osp.synth_c3(seed=3, n=65536, t0=i * 65536), fs)
        frames.append(r.power_db)
    frames = np.stack(frames)
    out["c3_frame0"] = frames[0]
    out["c3_frames_dec16"] = frames[:, ::16]
    out["c3_freqs_dec16"] = r.freqs[::16]
    out["c3_bin_hz"] = np.float64(r.bin_hz)
    out["c3_window_dec16"] = be.window[::16]
    be2 = ScipyFFTBackend(2048)
    r2 = be2.execute(osp.synth_c3(seed=4, n=5000, fs=2_400_000), 2_400_000)
    out["s2048_power"] = r2.power_db
    out["s2048_freqs"] = r2.freqs
    r3 = ScipyFFTBackend(512).execute(osp.synth_c3(seed=5, n=100), 48000)   # too few samples -> zeros
    out["short_power"] = r3.power_db
    np.savez_compressed(os.path.join(OUT, "spectrum.npz"), **out)


def c4fm_cases():
    """(name, sample_rate, chunk, n_frames, snr_db, cfo_hz, timing, seed) — shared with the tests."""
    return [
        ("c4fm_48k_2400", 48000, 2400, 8, 25.0, 120.0, 0.30, 41),      # generic path, 50 ms chunks
        ("c4fm_50k_2500", 50000, 2500, 8, 22.0, -150.0, 0.65, 42),     # fractional sps (10.4167)
        ("c4fm_48k_72000", 48000, 72000, 44, 28.0, 60.0, 0.10, 43),    # control-channel path: chunk > buffer half
        ("c4fm_48k_ragged", 48000, 1777, 6, 30.0, 0.0, 0.80, 44),      # chunk not a multiple of sps
    ]


def gen_p25_c4fm():
    """C4FMDemodulator.demodulate of the reference on seeded C4FM signals, replaying fixed chunk
    sequences (the output is chunking dependent). Inputs are stored (complex64) because they come
    out of float64 transcendental code whose last bit may differ between CPUs."""
    from wavecapsdr.dsp.p25.c4fm import C4FMDemodulator
    from oracle import c4fm as oc

    out = {}
    for name, fs, chunk, nfr, snr, cfo, timing, seed in c4fm_cases():
        rng = np.random.default_rng(seed)
        dib = oc.random_frames(rng, n_frames=nfr, payload=150, gap=40)
        x = oc.modulate_c4fm(dib, fs, snr_db=snr, cfo_hz=cfo, timing=timing, seed=seed)
        d = C4FMDemodulator(sample_rate=fs)
        ds, ss, cnt = [], [], []
        for s in range(0, len(x), chunk):
            a, b = d.demodulate(x[s:s + chunk])
            ds.append(a)
            ss.append(b)
            cnt.append(len(a))
        out[name + "_x"] = x
        out[name + "_dibits"] = np.concatenate(ds).astype(np.uint8)
        out[name + "_soft"] = np.concatenate(ss).astype(np.float32)
        out[name + "_counts"] = np.array(cnt, dtype=np.int32)
        out[name + "_sync_count"] = np.int32(d._sync_count)
        out[name + "_tx"] = dib
    np.savez_compressed(os.path.join(OUT, "p25_c4fm.npz"), **out)


def cqpsk_cases():
    """(name, sample_rate, symbol_rate, chunk, n_dibits, snr_db, cfo_hz, timing, seed)."""
    return [
        ("cqpsk_48k_2400", 48000, 4800, 2400, 2400, 25.0, 40.0, 0.30, 51),
        ("cqpsk_50k_2500", 50000, 4800, 2500, 2400, 22.0, -60.0, 0.70, 52),
        ("cqpsk_48k_72000", 48000, 4800, 72000, 9000, 28.0, 25.0, 0.10, 53),
        ("cqpsk_48k_6000baud_ragged", 48000, 6000, 1777, 2400, 30.0, 0.0, 0.55, 54),
        ("cqpsk_48k_tiny", 48000, 4800, 50, 300, 25.0, 30.0, 0.20, 55),     # chunks shorter than the 63-tap filter
    ]


def gen_p25_cqpsk():
    """decoders.p25.CQPSKDemodulator.demodulate of the reference on seeded pi/4-DQPSK signals, fixed chunk
    sequences. Inputs are stored (complex64), see gen_p25_c4fm."""
    import warnings

    from wavecapsdr.decoders.p25 import CQPSKDemodulator
    from oracle import cqpsk as oq

    warnings.filterwarnings("ignore")
    out = {}
    for name, fs, sr, chunk, nd, snr, cfo, timing, seed in cqpsk_cases():
        rng = np.random.default_rng(seed)
        tx = rng.integers(0, 4, nd).astype(np.uint8)
        x = oq.modulate_cqpsk(tx, fs, sr, snr_db=snr, cfo_hz=cfo, timing=timing, seed=seed)
        d = CQPSKDemodulator(sample_rate=fs, symbol_rate=sr)
        ds, cnt = [], []
        for s in range(0, len(x), chunk):
            a = d.demodulate(x[s:s + chunk])
            ds.append(a)
            cnt.append(len(a))
        out[name + "_x"] = x
        out[name + "_dibits"] = np.concatenate(ds).astype(np.uint8)
        out[name + "_counts"] = np.array(cnt, dtype=np.int32)
        out[name + "_state"] = np.array([float(d._freq_offset), float(d._phase_acc), float(d._symbol_clock),
                                         float(d._symbol_time), float(d._agc_gain)], dtype=np.float64)
        out[name + "_tx"] = tx
    np.savez_compressed(os.path.join(OUT, "p25_cqpsk.npz"), **out)


def gen_ddc():
    """wavecapsdr.dsp.filters.fir_filter_complex / fir_decimate (numba kernels) of the reference on seeded input,
    streamed over three calls with carried state."""
    from wavecapsdr.dsp.filters import fir_decimate, fir_filter_complex
    from scipy import signal as sg

    rng = np.random.default_rng(6)
    x = ((rng.standard_normal(30000) + 1j * rng.standard_normal(30000)) * 0.3).astype(np.complex64)
    taps = sg.firwin(157, 0.8 / 30, window=("kaiser", 7.857))
    out = {"x": x, "taps": taps}
    zi = sg.lfilter_zi(taps, 1.0).astype(np.complex128) * x[0]
    ys, cuts = [], [0, 12000, 12077, 30000]          # 12000 > 10000 takes the prange kernel, 77 < 156 taps
    for a, b in zip(cuts[:-1], cuts[1:]):
        y, zi = fir_decimate(x[a:b], taps, 30, zi=zi)
        ys.append(np.asarray(y))
    out["dec30"] = np.concatenate(ys)
    out["dec30_counts"] = np.array([len(y) for y in ys])
    out["dec30_zi"] = zi
    y, z = fir_filter_complex(x[:5000], taps[:73].copy(), None)
    out["filt73"] = np.asarray(y)
    out["filt73_zi"] = z
    np.savez_compressed(os.path.join(OUT, "ddc.npz"), **out)


def gen_audiofx():
    """dsp.filters.noise_blanker / spectral_noise_reduction and the FM chains with those flags on."""
    from wavecapsdr.dsp import filters as rf
    from wavecapsdr.dsp import fm as rfm
    from oracle import analog as oa

    rng = np.random.default_rng(12)
    x = (rng.standard_normal(20000) * 0.1 + 0.3 * np.sin(2 * np.pi * 700 / 48000 * np.arange(20000))).astype(np.float32)
    x[[100, 101, 5000, 19999]] = [3, -4, 2.5, 5]
    out = {"x": x}
    out["nb"] = rf.noise_blanker(x, 10.0, 3)
    out["nb_w0"] = rf.noise_blanker(x[:4001], 6.0, 0)
    out["nr"] = rf.spectral_noise_reduction(x, 48000, 12.0)
    out["nr18"] = rf.spectral_noise_reduction(x[:3000], 48000, 18.0)
    iq = oa.synth_c1(seed=1, n=120_000)
    out["wbfm_nb_nr"] = rfm.wbfm_demod(oa_shift(iq), 2_400_000, 48000, enable_noise_blanker=True, enable_noise_reduction=True)
    np.savez_compressed(os.path.join(OUT, "audiofx.npz"), **out)


def am_blanker_input():
    """48 kS/s AM-ish signal with impulse noise (the stable regime of the AM/SSB filters)."""
    rng = np.random.default_rng(31)
    n = 9600
    t = np.arange(n) / 48000.0
    x = ((0.3 * (1 + 0.5 * np.sin(2 * np.pi * 700 * t))) * np.exp(2j * np.pi * 1200 * t)
         + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    x[[50, 51, 4000, 9599]] += np.array([3, -4j, 2.5 + 2j, 5], dtype=np.complex64)
    return x


def gen_am_blanker():
    """am_demod / ssb_demod with enable_noise_blanker=True (dsp/am.py:100-101, 213-215)."""
    from wavecapsdr.dsp import am as ram

    x = am_blanker_input()
    out = {"am_nb": ram.am_demod(x, 48000, 16000, enable_noise_blanker=True, noise_blanker_threshold_db=8.0),
           "am_nb_noagc": ram.am_demod(x, 48000, 48000, enable_agc=False, enable_noise_blanker=True),
           "ssb_nb": ram.ssb_demod(x, 48000, 16000, mode="lsb", enable_noise_blanker=True, noise_blanker_threshold_db=6.0)}
    np.savez_compressed(os.path.join(OUT, "am_blanker.npz"), **out)


def sam_input(n=14400, fs=48000, carrier_hz=7.0, seed=41):
    """AM broadcast-like signal: carrier `carrier_hz` off centre with an initial phase, two audio tones, noise and one impulse."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / float(fs)
    env = 0.4 * (1 + 0.45 * np.sin(2 * np.pi * 620 * t) + 0.25 * np.sin(2 * np.pi * 1870 * t + 0.4))
    x = env * np.exp(1j * (2 * np.pi * carrier_hz * t + 0.9)) + 0.004 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x[n // 3] += 2.0 - 1.0j
    return x.astype(np.complex64)


def gen_sam():
    """dsp/sam.py: CarrierRecoveryPLL.process over two consecutive calls (carried state), sam_demod in its three sideband
    settings, sam_demod_simple, and capture._process_channel_dsp_stateless(mode='sam') with an offset."""
    import wavecapsdr.capture as rc
    from wavecapsdr.dsp import sam as rs

    x = sam_input()
    out = {}
    pll = rs.CarrierRecoveryPLL(sample_rate=48000.0, loop_bandwidth=50.0)
    for k, part in enumerate((x[:5001], x[5001:9000])):
        ci, cq, f = pll.process(part)
        out[f"pll_i{k}"], out[f"pll_q{k}"] = ci, cq
        out[f"pll_state{k}"] = np.array([pll._phase, pll._frequency, pll._integrator, f], dtype=np.float64)
    a, f, st = rs.sam_demod(x, 48000, 48000)
    out["dsb"], out["dsb_f"] = a, np.float64(f)
    a2, f2, _ = rs.sam_demod(x[:4000], 48000, 48000, pll_state=st)          # carried PLL
    out["dsb_cont"], out["dsb_cont_f"] = a2, np.float64(f2)
    out["usb_noagc"] = rs.sam_demod(x, 48000, 16000, sideband="usb", pll_bandwidth=30.0, enable_agc=False, lowpass_hz=3000.0)[0]
    out["lsb_nb_notch"] = rs.sam_demod(x, 48000, 16000, sideband="LSB", pll_bandwidth=100.0, pll_damping=1.0,
                                       enable_noise_blanker=True, noise_blanker_threshold_db=8.0,
                                       notch_frequencies=[1870.0, 30000.0])[0]
    out["simple"] = rs.sam_demod_simple(x, 48000, 24000, sideband="dsb", enable_highpass=False)
    # the capture path, channel SAM_OFFSET_HZ off centre: 48 kS/s (the regime where the tf-form 100 Hz high-pass is well enough
    # conditioned for a 1e-4 comparison) and one 240 kS/s case (kept to document the reference's own floor there)
    for fs, tag in ((48000, "st"), (240000, "st240")):
        xs = sam_stateless_input(fs)
        for name, kw in sam_stateless_cases():
            if fs == 240000 and name != "default":
                continue
            cfg = rc.ChannelConfig(id="s", capture_id="c", mode="sam", offset_hz=SAM_OFFSET_HZ)
            for k, v in kw.items():
                setattr(cfg, k, v)
            audio, m = rc._process_channel_dsp_stateless(xs, fs, cfg)
            out[f"{tag}_{name}"] = audio
            out[f"{tag}_{name}_m"] = np.array([m["rssi_db"], m["signal_power_db"]], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "sam.npz"), **out)


SAM_OFFSET_HZ = 6000.0


def sam_stateless_input(fs=48000):
    n = fs // 4
    x = sam_input(n=n, fs=fs, carrier_hz=-11.0, seed=43)
    t = np.arange(n) / float(fs)
    return (x * np.exp(2j * np.pi * SAM_OFFSET_HZ * t)).astype(np.complex64)


def sam_stateless_cases():
    return [("default", {}), ("usb_agc", {"sam_sideband": "usb", "enable_agc": True, "sam_pll_bandwidth_hz": 80.0}),
            ("lsb_16k", {"sam_sideband": "lsb", "audio_rate": 16000, "enable_am_highpass": False})]


def oa_shift(iq):
    """the C1 carrier sits at +200 kHz: bring it to baseband like capture.freq_shift does"""
    import wavecapsdr.capture as rc

    return rc.freq_shift(iq, 200000.0, 2_400_000)



def framer_streams():
    """(name, dibits uint8, soft float32, chunk) — symbol streams with valid NIDs (oracle.bch.bch_encode), shared with the
    tests: contiguous 3-TSBK TSDUs, a voice call, single TSBKs with gaps, mixed/unknown DUIDs with uncorrectable NIDs,
    noise. Soft symbols = ideal levels + seeded Gaussian noise."""
    from oracle import p25_framer as of

    out = []
    rng = np.random.default_rng(11)
    s = list(rng.integers(0, 4, 70))
    for k in range(8):
        s += of.frame_dibits(rng, 0x293, 0x7, 588, nid_errors=k % 5)
    s += list(rng.integers(0, 4, 100))
    out.append(("tsdu3", s, 720))
    s = list(rng.integers(0, 4, 40))
    for duid, nb in [(0x0, 648), (0x5, 1568), (0xA, 1568), (0x5, 1568), (0xA, 1568), (0xF, 168), (0x3, 28), (0x3, 28)]:
        s += of.frame_dibits(rng, 0x4A1, duid, nb, nid_errors=int(rng.integers(0, 8)))
        while len(s) % 36 != 40 % 36:
            s.append(int(rng.integers(0, 4)))
    s += list(rng.integers(0, 4, 120))
    out.append(("voice", s, 500))
    s = list(rng.integers(0, 4, 60))
    for k in range(6):
        s += of.frame_dibits(rng, 0x293, 0x7, 196)
        s += list(rng.integers(0, 4, int(rng.integers(0, 30))))
    s += list(rng.integers(0, 4, 80))
    out.append(("tsbk1_gaps", s, 300))
    s = list(rng.integers(0, 4, 30))
    for k in range(14):
        duid = [0xC, 0x9, 0x7, 0x3, 0xE, 0xD, 0x5][k % 7]
        nac = [0x293, 0x293, 0x111, 0x293][k % 4]
        nb = {0xC: 196 * 3, 0x9: 300, 0x7: 588, 0x3: 28, 0xE: 100, 0xD: 400, 0x5: 1568}[duid]
        s += of.frame_dibits(rng, nac, duid, nb, nid_errors=[0, 3, 12, 11, 13][k % 5])
        s += list(rng.integers(0, 4, int(rng.integers(0, 50))))
    out.append(("mixed", s, 777))
    out.append(("noise", list(rng.integers(0, 4, 6000)), 1000))
    res = []
    for name, s, chunk in out:
        dib = np.array(s, dtype=np.uint8)
        soft = (of.dibits_to_soft(dib) + np.random.default_rng(5).normal(0, 0.3, len(dib))).astype(np.float32)
        res.append((name, dib, soft, chunk))
    return res


def framer_e2e_signal():
    """48 kS/s C4FM signal carrying a train of valid 3-TSBK TSDUs (the SURVEY §8d C4 control-channel recipe)."""
    from oracle import c4fm as oc, p25_framer as of

    rng = np.random.default_rng(77)
    s = list(rng.integers(0, 4, 90))
    for k in range(10):
        s += of.frame_dibits(rng, 0x293, 0x7, 588, nid_errors=k % 3)
    s += list(rng.integers(0, 4, 150))
    dib = np.array(s, dtype=np.uint8)
    return dib, oc.modulate_c4fm(dib, 48000, snr_db=26.0, cfo_hz=40.0, timing=0.35, seed=9)


def _pack_msgs(msgs):
    meta = np.array([[int(m.duid), int(m.nac), int(m.timestamp), int(m.corrected_bit_count), len(m.bits)] for m in msgs],
                    dtype=np.int64).reshape(-1, 5)
    bits = np.concatenate([np.asarray(m.bits, dtype=np.uint8) for m in msgs]) if msgs else np.zeros(0, np.uint8)
    return meta, bits


def gen_p25_framer():
    """bch_decode and P25P1MessageFramer of the live reference: (a) 1200 codewords with 0-15 flipped bits / random words,
    with and without a tracked NAC; (b) the framer_streams() through process_batch (chunked; AssertionErrors recorded where
    the reference raises them) and through process_with_soft_sync symbol by symbol; (c) C4FMDemodulator.demodulate ->
    process_batch / process_with_soft_sync on a C4FM signal with valid TSDUs."""
    import json

    from wavecapsdr.decoders.p25_framer import P25P1MessageFramer
    from wavecapsdr.dsp.fec.bch import bch_decode
    from wavecapsdr.dsp.p25.c4fm import C4FMDemodulator
    from oracle import bch as ob

    out = {}
    rng = np.random.default_rng(21)
    cws, trs, ds, es = [], [], [], []
    for t in range(1200):
        if t % 4 == 3:
            c = rng.integers(0, 2, 63).astype(np.uint8)
            d = 0
        else:
            d = int(rng.integers(0, 65536))
            c = ob.bch_encode(d).copy()
            c[rng.choice(63, int(rng.integers(0, 16)), replace=False)] ^= 1
        tr = [0, 0x293, d >> 4][t % 3]
        a = bch_decode(c, tr if tr else None)
        cws.append(c); trs.append(tr); ds.append(int(a[0])); es.append(int(a[1]))
    out["bch_cw"] = np.array(cws, dtype=np.uint8)
    out["bch_tracked"] = np.array(trs, dtype=np.int32)
    out["bch_data"] = np.array(ds, dtype=np.int32)
    out["bch_errors"] = np.array(es, dtype=np.int32)

    TS = 1_700_000_000_000

    def run_batch(chunks):
        fr = P25P1MessageFramer(); msgs = []
        fr.set_listener(msgs.append); fr.start(); fr.set_timestamp(TS)
        log = []
        for soft, dib in chunks:
            try:
                log.append(int(fr.process_batch(soft, dib)))
            except AssertionError as e:
                log.append(str(e))
        return msgs, log

    def run_stream(soft, dib, max_errors=40):
        fr = P25P1MessageFramer(); msgs = []
        fr.set_listener(msgs.append); fr.start(); fr.set_timestamp(TS)
        log = []
        for i in range(len(dib)):
            try:
                if fr.process_with_soft_sync(float(soft[i]), int(dib[i])):
                    log.append(i)
            except AssertionError as e:
                log.append([i, str(e)])
                if len(log) > max_errors:
                    break
        return msgs, log

    names = []
    for name, dib, soft, chunk in framer_streams():
        names.append(name)
        out[f"{name}_dibits"] = dib
        out[f"{name}_soft"] = soft
        out[f"{name}_chunk"] = np.int32(chunk)
        m, log = run_batch([(soft[s:s + chunk], dib[s:s + chunk]) for s in range(0, len(dib), chunk)])
        out[f"{name}_batch_meta"], out[f"{name}_batch_bits"] = _pack_msgs(m)
        out[f"{name}_batch_log"] = np.array(json.dumps(log))
        m, log = run_stream(soft, dib)
        out[f"{name}_stream_meta"], out[f"{name}_stream_bits"] = _pack_msgs(m)
        out[f"{name}_stream_log"] = np.array(json.dumps(log))
    out["stream_names"] = np.array(json.dumps(names))

    tx, x = framer_e2e_signal()
    d = C4FMDemodulator(sample_rate=48000)
    chunks = []
    for s in range(0, len(x), 2400):
        a, b = d.demodulate(x[s:s + 2400])
        chunks.append((np.asarray(b, dtype=np.float32), np.asarray(a, dtype=np.uint8)))
    out["e2e_x"] = x
    out["e2e_tx"] = tx
    out["e2e_counts"] = np.array([len(c[1]) for c in chunks], dtype=np.int32)
    out["e2e_dibits"] = np.concatenate([c[1] for c in chunks])
    out["e2e_soft"] = np.concatenate([c[0] for c in chunks])
    m, log = run_batch(chunks)
    out["e2e_batch_meta"], out["e2e_batch_bits"] = _pack_msgs(m)
    out["e2e_batch_log"] = np.array(json.dumps(log))
    m, log = run_stream(out["e2e_soft"], out["e2e_dibits"])
    out["e2e_stream_meta"], out["e2e_stream_bits"] = _pack_msgs(m)
    out["e2e_stream_log"] = np.array(json.dumps(log))
    np.savez_compressed(os.path.join(OUT, "p25_framer.npz"), **out)



def c4fm_disc_cases():
    """(name, sample_rate, chunk, seed, audio dtype) — discriminator-audio entry, shared with the tests."""
    return [("disc_48k_2400", 48000, 2400, 51, "float64"), ("disc_48k_7000", 48000, 7000, 52, "float32"),
            ("disc_50k_2500", 50000, 2500, 53, "float64"), ("disc_48k_ragged", 48000, 1777, 54, "float64")]


def gen_p25_c4fm_disc():
    """C4FMDemodulator.demodulate_discriminator of the reference (c4fm.py:2817-2992) on np.diff(np.unwrap(np.angle(iq)))
    of seeded C4FM signals, replaying fixed chunk sequences; the last case also calls reset() half way (which leaves the
    discriminator RRC state alone in the reference)."""
    from wavecapsdr.dsp.p25.c4fm import C4FMDemodulator
    from oracle import c4fm as oc

    out = {}
    for name, fs, chunk, seed, dt in c4fm_disc_cases():
        rng = np.random.default_rng(seed)
        dib = oc.random_frames(rng, n_frames=10, payload=150, gap=40)
        x = oc.modulate_c4fm(dib, fs, snr_db=25.0, cfo_hz=50.0, timing=0.4, seed=seed)
        au = oc.discriminator_audio(x).astype(dt)
        d = C4FMDemodulator(sample_rate=fs)
        ds, ss, cnt = [], [], []
        starts = list(range(0, len(au), chunk))
        for j, s0 in enumerate(starts):
            if name.endswith("ragged") and j == len(starts) // 2:
                d.reset()
            a, b = d.demodulate_discriminator(au[s0:s0 + chunk])
            ds.append(a); ss.append(b); cnt.append(len(a))
        out[name + "_audio"] = au
        out[name + "_dibits"] = np.concatenate(ds).astype(np.uint8)
        out[name + "_soft"] = np.concatenate(ss).astype(np.float32)
        out[name + "_counts"] = np.array(cnt, dtype=np.int32)
        out[name + "_state"] = np.array([float(d._fine_sync), d._sample_point, d._equalizer.gain, d._equalizer.pll])
    np.savez_compressed(os.path.join(OUT, "p25_c4fm_disc.npz"), **out)



def discriminator_cases():
    """(name, chunk, seed, snr_db, cfo_hz) — 48 kS/s voice-channel IQ -> discriminator audio -> dibits; shared with the tests."""
    return [("disc_2400", 2400, 61, 24.0, 80.0), ("disc_7000", 7000, 62, 20.0, -120.0), ("disc_999", 999, 63, 26.0, 30.0),
            ("disc_60", 60, 64, 24.0, 0.0)]


def gen_p25_discriminator():
    """VoiceRecorder's discriminator (trunking/system.py:708-717, restated inline exactly as the reference writes it) and
    decoders.p25.DiscriminatorDemodulator.demodulate of the live reference, replaying fixed chunk sequences; disc_999 also
    calls reset() half way."""
    from wavecapsdr.decoders.p25 import DiscriminatorDemodulator
    from oracle import c4fm as oc

    out = {}
    for name, chunk, seed, snr, cfo in discriminator_cases():
        rng = np.random.default_rng(seed)
        dib = oc.random_frames(rng, n_frames=12, payload=150, gap=40)
        x = oc.modulate_c4fm(dib, 48000, snr_db=snr, cfo_hz=cfo, timing=0.3, seed=seed)
        d = DiscriminatorDemodulator(sample_rate=48000)
        last = 0.0
        aud, ds, cnt = [], [], []
        starts = list(range(0, len(x), chunk))
        for j, s0 in enumerate(starts):
            iq = x[s0:s0 + chunk]
            phase = np.angle(iq)
            up = np.unwrap(np.concatenate([[last], phase]))
            last = up[-1] if len(up) > 1 else last
            au = np.diff(up)
            if name == "disc_999" and j == len(starts) // 2:
                d.reset()
            a = d.demodulate(au.astype(np.float32))
            aud.append(au); ds.append(a); cnt.append(len(a))
        out[name + "_x"] = x
        out[name + "_audio"] = np.concatenate(aud)
        out[name + "_dibits"] = np.concatenate(ds).astype(np.uint8)
        out[name + "_counts"] = np.array(cnt, dtype=np.int32)
        out[name + "_state"] = np.array([float(d._input_gain), float(d._dc_estimate), float(d._symbol_clock),
                                         float(d._symbol_spread), float(d._fine_freq_correction)])
    np.savez_compressed(os.path.join(OUT, "p25_discriminator.npz"), **out)



def gen_p25_trellis():
    """dsp.fec.trellis.trellis_decode and the TSBK block decode (P25Decoder._deinterleave_data + P25TrellisDecoder.decode)
    of the live reference on encoded blocks with injected dibit errors: (a) 300 blocks of random length 1-60 input dibits,
    every third with soft values, every fifth with an odd received length; (b) 200 interleaved TSBK blocks."""
    from wavecapsdr.decoders.p25 import P25Decoder, P25TrellisDecoder
    from wavecapsdr.dsp.fec.trellis import trellis_decode
    from oracle import trellis as ot

    rng = np.random.default_rng(31)
    rx_all = np.zeros((300, 120), dtype=np.uint8)
    soft_all = np.zeros((300, 120), dtype=np.float64)
    has_soft = np.zeros(300, dtype=np.uint8)
    lens = np.zeros(300, dtype=np.int32)
    dec_all = np.zeros((300, 60), dtype=np.uint8)
    dec_len = np.zeros(300, dtype=np.int32)
    met = np.zeros(300, dtype=np.int32)
    for t in range(300):
        n = int(rng.integers(1, 61))
        rx = ot.encode(rng.integers(0, 4, n)).copy()
        for p in rng.integers(0, len(rx), int(rng.integers(0, 8))):
            rx[p] = rng.integers(0, 4)
        if t % 5 == 0 and len(rx) > 1:
            rx = rx[:-1]
        soft = None
        if t % 3 == 0:
            soft = np.array(ot.LEVEL)[rx] + rng.normal(0, 0.8, len(rx))
            has_soft[t] = 1
            soft_all[t, :len(rx)] = soft
        d, m = trellis_decode(rx, soft)
        rx_all[t, :len(rx)] = rx
        lens[t] = len(rx)
        dec_all[t, :len(d)] = d
        dec_len[t] = len(d)
        met[t] = m
    out = dict(rx=rx_all, soft=soft_all, has_soft=has_soft, lens=lens, dec=dec_all, dec_len=dec_len, metric=met)
    bits = np.zeros((200, 196), dtype=np.uint8)
    dec96 = np.zeros((200, 96), dtype=np.uint8)
    tmet = np.zeros(200, dtype=np.int32)
    tx96 = np.zeros((200, 96), dtype=np.uint8)
    deint = np.array(P25Decoder.DATA_DEINTERLEAVE)
    for t in range(200):
        msg = np.concatenate([rng.integers(0, 4, 48), [0]])
        blk = ot.interleave(ot.encode(msg))
        for p in rng.integers(0, 98, int(rng.integers(0, 9))):
            blk[p] = rng.integers(0, 4)
        bits[t, 0::2] = blk >> 1
        bits[t, 1::2] = blk & 1
        d, m = P25TrellisDecoder().decode(blk[deint].astype(np.uint8))
        dec96[t, 0::2] = (d[:48] >> 1) & 1
        dec96[t, 1::2] = d[:48] & 1
        tmet[t] = m
        tx96[t, 0::2] = (msg[:48] >> 1) & 1
        tx96[t, 1::2] = msg[:48] & 1
    out.update(tsbk_bits=bits, tsbk_dec96=dec96, tsbk_metric=tmet, tsbk_tx96=tx96)
    np.savez_compressed(os.path.join(OUT, "p25_trellis.npz"), **out)



def gen_cc_scanner():
    """ControlChannelScanner.scan_all of the live reference on oracle.cc_scanner.synth_band() (regenerated from its seed in
    the tests: the comparison is at float tolerance, not bit level)."""
    from wavecapsdr.trunking.cc_scanner import ControlChannelScanner
    from oracle import cc_scanner as oc

    x, center, freqs = oc.synth_band()
    sc = ControlChannelScanner(center_hz=center, sample_rate=1_200_000, control_channels=freqs)
    m = sc.scan_all(x)
    rows = []
    for f in freqs:
        if f in m:
            r = m[f]
            rows.append([f, r.power_db, r.peak_power_db, r.noise_floor_db, r.snr_db, float(r.sync_detected), r.sample_count])
    best = sc.get_best_channel()
    np.savez_compressed(os.path.join(OUT, "cc_scanner.npz"), rows=np.array(rows, dtype=np.float64), best=np.float64(best[0]),
                        ranking=np.array([f for f, _ in sc.get_channel_ranking()], dtype=np.float64),
                        x_checksum=np.float64(np.sum(np.abs(x.astype(np.complex128)) ** 2)))


GENERATORS = {"sam": gen_sam, "channel_calc": gen_channel_calc, "am_blanker": gen_am_blanker, "cc_scanner": gen_cc_scanner, "p25_trellis": gen_p25_trellis, "p25_discriminator": gen_p25_discriminator, "p25_c4fm_disc": gen_p25_c4fm_disc, "p25_framer": gen_p25_framer, "audiofx": gen_audiofx, "ddc": gen_ddc, "p25_c4fm": gen_p25_c4fm, "p25_cqpsk": gen_p25_cqpsk, "channelizer_c5": gen_channelizer_c5, "analog": gen_analog, "spectrum": gen_spectrum}


def main(argv):
    refenv.load()
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(GENERATORS)
    for name in names:
        GENERATORS[name]()
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1:])
