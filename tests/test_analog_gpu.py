"""GPU parity: csrc/analog.cu + host chain vs the oracle restatement of the reference's analog path.
Tolerance (BASELINE.json north_star): audio within 1e-4 relative RMS; dB metrics within 1e-3 dB."""
import numpy as np
import pytest

from conftest import rel_rms, golden_path, wrap_rel_rms
from oracle import analog as oa

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def g():
    return np.load(golden_path("analog.npz"))


def _cfg(native, **kw):
    from wavecap_sdr_b200.capture import ChannelConfig

    return ChannelConfig(id="t", capture_id="c", **kw)


def test_stage_functions_against_golden(native, g):
    from wavecap_sdr_b200 import capture as C
    from wavecap_sdr_b200.dsp import agc, am, fm

    xa, fs = g["am_x"], 48000
    f = np.real(xa).astype(np.float32)
    assert rel_rms(fm.quadrature_demod(xa, fs), g["quad"]) < TOL
    assert rel_rms(fm.deemphasis_filter(f, fs), g["deemph"]) < TOL
    assert rel_rms(fm.lpf_audio(f, fs, 5000), g["lpf"]) < TOL
    assert rel_rms(fm.resample_poly(f, 48000, 14400), g["resamp_3_10"]) < TOL
    assert rel_rms(fm.resample_poly(f[:1000], 8000, 48000), g["resamp_up"]) < TOL
    assert rel_rms(agc.apply_agc(f, fs), g["agc"]) < TOL
    assert rel_rms(C.freq_shift(xa, 1234.4, fs), g["fshift"]) < TOL
    assert rel_rms(am.freq_shift(xa, 1500.0, fs), g["am_fshift"]) < TOL
    assert rel_rms(am.am_demod(xa, fs, 16000), g["am_audio"]) < TOL
    assert rel_rms(am.am_demod(xa, fs, 16000, enable_agc=False), g["am_audio_noagc"]) < TOL
    assert rel_rms(am.ssb_demod(xa, fs, 16000), g["ssb_audio"]) < TOL
    assert rel_rms(am.ssb_demod(xa, fs, 16000, mode="lsb", enable_agc=False), g["ssb_audio_lsb"]) < TOL
    assert rel_rms(fm.soft_clip(f * 3), oa.soft_clip_fm(f * 3)) < 1e-6
    assert rel_rms(agc.soft_clip(f * 3), oa.soft_clip_agc(f * 3)) < 1e-6
    assert rel_rms(fm.rms_normalize(f), oa.rms_normalize(f)) < 1e-6
    assert fm.quadrature_demod(np.zeros(0, np.complex64), fs).size == 0
    assert C.freq_shift(xa, 0.0, fs) is xa


def test_c1_wbfm_golden_and_oracle(native, g):
    from wavecap_sdr_b200.capture import _process_channel_dsp_stateless

    cfg = _cfg(native, mode="wbfm", offset_hz=200000.0)
    for i in range(2):
        x = oa.synth_c1(seed=1, n=120_000, t0=i * 120_000)
        a, m = _process_channel_dsp_stateless(x, 2_400_000, cfg)
        assert a.dtype == np.float32 and a.shape == (2400,)
        assert rel_rms(a, g[f"c1_audio{i}"]) < TOL
        assert abs(m["rssi_db"] - g[f"c1_metrics{i}"][0]) < 1e-3
        assert abs(m["signal_power_db"] - g[f"c1_metrics{i}"][1]) < 1e-3


def test_c1_batched_chunks_equal_per_chunk_calls(native):
    from wavecap_sdr_b200.capture import process_channels_batch

    cfg = _cfg(native, mode="wbfm", offset_hz=200000.0)
    xs = [oa.synth_c1(seed=5, n=120_000, t0=i * 120_000) for i in range(3)]
    res = process_channels_batch(np.concatenate(xs), 2_400_000, [cfg], n_chunks=3)
    for i in range(3):
        exp, m = oa.process_channel_dsp_stateless(xs[i], 2_400_000, oa.OracleChannelConfig(mode="wbfm", offset_hz=200000.0))
        assert rel_rms(res[i][0][0], exp) < TOL
        assert abs(res[i][0][1]["rssi_db"] - m["rssi_db"]) < 1e-3


def test_c2_16_nbfm_int16_with_squelch(native, g):
    from wavecap_sdr_b200.capture import apply_mode_defaults, process_channels_batch

    q, offs = oa.synth_c2(seed=2, n=500_000, keyed_off=(3, 12))
    cfgs = [apply_mode_defaults("nbfm", _cfg(native, mode="nbfm", offset_hz=float(o), squelch_db=-45.0)) for o in offs]
    res = process_channels_batch(q[None], 10_000_000, cfgs, n_chunks=1, in_fmt="cs16")[0]
    for k in range(16):
        a, m = res[k]
        assert rel_rms(a, g["c2_audio"][k]) < TOL, k
        assert abs(m["rssi_db"] - g["c2_metrics"][k][0]) < 1e-3
        assert abs(m["signal_power_db"] - g["c2_metrics"][k][1]) < 1e-3
    # squelch uses total capture power (capture.py:331-334, 2918-2921): one threshold either side
    hi = [apply_mode_defaults("nbfm", _cfg(native, mode="nbfm", offset_hz=float(o), squelch_db=-10.0)) for o in offs]
    res_sq = process_channels_batch(q[None], 10_000_000, hi, n_chunks=1, in_fmt="cs16", apply_squelch=True)[0]
    assert all(not r[0].any() for r in res_sq)
    res_open = process_channels_batch(q[None], 10_000_000, cfgs, n_chunks=1, in_fmt="cs16", apply_squelch=True)[0]
    assert all(r[0].any() for r in res_open)


def test_mixed_modes_one_capture(native):
    from wavecap_sdr_b200.capture import process_channels_batch

    fs, n = 48000, 9600
    rng = np.random.default_rng(11)
    t = np.arange(n) / fs
    x = ((1 + 0.4 * np.sin(2 * np.pi * 500 * t)) * 0.2 * np.exp(2j * np.pi * 3000 * t)
         + 0.1 * np.exp(1j * (2 * np.pi * -6000 * t + 2.0 * np.sin(2 * np.pi * 800 * t)))
         + 0.005 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    specs = [dict(mode="am", offset_hz=3000.0, enable_agc=True, audio_rate=16000),
             dict(mode="nbfm", offset_hz=-6000.0, enable_deemphasis=False, audio_rate=16000),
             dict(mode="ssb", offset_hz=3000.0, enable_agc=True, audio_rate=16000),
             dict(mode="nbfm", offset_hz=-6000.0, enable_deemphasis=True, enable_fm_lowpass=True,
                  enable_fm_highpass=True, fm_highpass_hz=300, notch_frequencies=[1000.0], audio_rate=8000),
             dict(mode="raw", offset_hz=1000.0), dict(mode="p25", offset_hz=0.0),
             dict(mode="am", offset_hz=0.0, enable_agc=False, enable_am_highpass=False, audio_rate=48000)]
    res = process_channels_batch(x, fs, [_cfg(native, **s) for s in specs])[0]
    for s, (a, m) in zip(specs, res):
        exp, me = oa.process_channel_dsp_stateless(x, fs, oa.OracleChannelConfig(**s))
        assert (a is None) == (exp is None), s
        for key in me:
            assert abs(m[key] - me[key]) < 1e-3, (s, key)
        if exp is not None:
            assert rel_rms(a, exp) < TOL, s


def test_nonfinite_chunk_is_dropped_and_empty_input(native):
    from wavecap_sdr_b200.capture import _process_channel_dsp_stateless, process_channels_batch

    cfg = _cfg(native, mode="nbfm", offset_hz=1000.0, enable_deemphasis=False)
    assert _process_channel_dsp_stateless(np.zeros(0, np.complex64), 48000, cfg) == (None, {})
    x = oa.synth_c1(seed=3, n=9600 * 2, fs=48000, offset=1000.0, dev=3000.0)
    x[9600 + 17] = np.nan
    res = process_channels_batch(x, 48000, [cfg], n_chunks=2)
    assert res[1][0] == (None, {})
    exp, _ = oa.process_channel_dsp_stateless(x[:9600], 48000, oa.OracleChannelConfig(mode="nbfm", offset_hz=1000.0, enable_deemphasis=False))
    assert rel_rms(res[0][0][0], exp) < TOL


def test_long_iir_matches_sequential_recursion(native):
    """Block-scan lfilter vs scipy on a 500 000-sample sequence (several tiles, order 1/2/5/10)."""
    from scipy import signal
    from wavecap_sdr_b200.dsp import filters as F

    rng = np.random.default_rng(4)
    x = rng.standard_normal(500_000).astype(np.float32)
    fs = 2_400_000
    assert rel_rms(F.lowpass_filter(x, fs, 15000), oa.lowpass_filter(x, fs, 15000)) < TOL
    assert rel_rms(F.highpass_filter(x, 48000, 300), oa.highpass_filter(x, 48000, 300)) < TOL
    # order-10 tf-form band-pass (ssb_demod, dsp/filters.py:177-217): too ill-conditioned for the block scan, so the handle
    # replays lfilter's recursion operation for operation — bit-equal to scipy, not just inside the 1e-4 budget
    got_bp, exp_bp = F.bandpass_filter(x, 48000, 300, 3000), oa.bandpass_filter(x, 48000, 300, 3000)
    assert rel_rms(got_bp, exp_bp) < TOL
    assert np.array_equal(got_bp, exp_bp)
    assert rel_rms(F.notch_filter(x, 48000, 1000.0), oa.notch_filter(x, 48000, 1000.0)) < TOL
    # invalid cut-offs return the input unchanged
    assert np.array_equal(F.lowpass_filter(x[:100], 48000, 30000), x[:100])


def test_ill_conditioned_iirs_replay_lfilter_bit_for_bit(native):
    """where the tf-form recursion is ill-conditioned (3 kHz low-pass / 300 Hz high-pass at
    multi-MS/s rates, the AM chain's 100 Hz high-pass) only an exact replay of the sequential float64 recursion matches
    the reference. wc_iir_create picks the replay kernel for those filters and keeps the block scan for the rest."""
    from wavecap_sdr_b200 import _native as N
    from wavecap_sdr_b200.dsp import _stages as S
    from wavecap_sdr_b200.dsp import filters as F

    def is_seq(coeffs):
        b, a = coeffs
        return bool(N.lib().wc_iir_is_sequential(S.iir_handle(tuple(map(float, b)), tuple(map(float, a))).h))

    assert is_seq(F.bandpass_coeffs(48000, 300, 3000)) and is_seq(F.lowpass_coeffs(10_000_000, 3000))
    assert is_seq(F.highpass_coeffs(976_560, 300)) and is_seq(F.highpass_coeffs(48000, 100))
    assert not is_seq(F.lowpass_coeffs(2_400_000, 15000)) and not is_seq(F.highpass_coeffs(48000, 300))
    assert not is_seq(F.notch_coeffs(48000, 1000.0)) and not is_seq(F.lowpass_coeffs(48000, 3000))

    def kind(coeffs):
        b, a = coeffs
        return int(N.lib().wc_iir_kind(S.iir_handle(tuple(map(float, b)), tuple(map(float, a))).h))

    from wavecap_sdr_b200.dsp import fm as FMm

    # one-pole filters (de-emphasis, AGC envelopes) chain their scan in plain float64; the order-5 MPX low-pass (|A^64| ~ 7e5)
    # needs the double-double chain; the order-10 band-pass is replayed sequentially
    assert kind(FMm.deemphasis_coeffs(2_400_000)) == 1 and kind(FMm.mpx_coeffs(2_400_000)) == 0
    assert kind(F.bandpass_coeffs(48000, 300, 3000)) == 2
    y = rng_check = np.random.default_rng(42).standard_normal(300_000).astype(np.float32)
    assert rel_rms(FMm.lpf_audio(y, 2_400_000), oa.lpf_audio(y, 2_400_000)) < 1e-6
    assert rel_rms(FMm.deemphasis_filter(y, 2_400_000), oa.deemphasis_filter(y, 2_400_000)) < 1e-6
    rng = np.random.default_rng(41)
    x = rng.standard_normal((5, 30_011)).astype(np.float32)       # 5 sequences, ragged against the 64-sample staging tile
    for fs, fn, ofn, args in ((10_000_000, F.lowpass_filter, oa.lowpass_filter, (3000,)),
                              (976_560, F.highpass_filter, oa.highpass_filter, (300,)),
                              (48000, F.highpass_filter, oa.highpass_filter, (100,)),
                              (48000, F.bandpass_filter, oa.bandpass_filter, (300, 3000))):
        got = fn(x, fs, *args)
        exp = np.stack([ofn(r, fs, *args) for r in x])
        assert got.shape == exp.shape and np.array_equal(got, exp), (fs, args)


def test_shared_iir_handle_from_two_threads(native):
    """dsp/_stages.iir_handle caches one handle per coefficient set and the reference calls the stateless chain from a
    3-worker pool (capture.py:1906-1925): concurrent calls on one handle, each on its own stream, must not share scratch."""
    import threading

    import torch

    from wavecap_sdr_b200.dsp import filters as F

    rng = np.random.default_rng(43)
    xs = [rng.standard_normal(n).astype(np.float32) for n in (400_000, 90_000, 250_000)]
    exp = [oa.lowpass_filter(x, 2_400_000, 15000) for x in xs]
    errs = []

    def worker(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(6):
                    got = F.lowpass_filter(torch.from_numpy(xs[i]).cuda(), 2_400_000, 15000)
                    torch.cuda.current_stream().synchronize()
                    if rel_rms(got.cpu().numpy(), exp[i]) >= TOL:
                        errs.append(i)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_decimate_iq_for_p25(native):
    from wavecap_sdr_b200.capture import decimate_iq_for_p25
    from scipy import signal

    x = oa.synth_c1(seed=8, n=24000, fs=240_000, offset=2000.0, dev=2000.0)
    y, r = decimate_iq_for_p25(x, 240_000)            # down = 5 -> resample_poly on I and Q
    exp = (signal.resample_poly(x.real, 1, 5) + 1j * signal.resample_poly(x.imag, 1, 5)).astype(np.complex64)
    assert r == 48000 and rel_rms(y, exp) < TOL
    y2, r2 = decimate_iq_for_p25(x, 4_800_000)         # down = 100 > 50 -> plain subsample
    assert r2 == 48000 and np.array_equal(y2, x[::100])
    y3, r3 = decimate_iq_for_p25(x, 48000)
    assert y3 is x and r3 == 48000


def test_rds_path_input_is_the_front_end_output(native):
    """capture.py:2869-2884: the RDS decoder is fed quadrature_demod(freq_shift(iq)); process_channels_batch hands it out."""
    from oracle import analog as oa
    from wavecap_sdr_b200.capture import ChannelConfig, apply_mode_defaults, process_channels_batch

    fs, n = 2_400_000, 24_000
    rng = np.random.default_rng(12)
    x = ((rng.standard_normal(2 * n) + 1j * rng.standard_normal(2 * n)) * 0.2).astype(np.complex64)
    cfg = apply_mode_defaults("wbfm", ChannelConfig(id="a", capture_id="c", mode="wbfm", offset_hz=150000.0))
    res = process_channels_batch(x, fs, [cfg], n_chunks=2, want_fm_baseband=True)
    for b in range(2):
        ref = oa.quadrature_demod(oa.freq_shift(x[b * n:(b + 1) * n], 150000.0, fs), fs)
        got = res[b][0][1]["fm_baseband"]
        period = 2 * np.pi * float(np.float32(fs / (2.0 * np.pi * 75000.0)))
        assert wrap_rel_rms(got, ref, period) < 1e-5
    cfg.enable_rds = False
    assert "fm_baseband" not in process_channels_batch(x, fs, [cfg], n_chunks=2, want_fm_baseband=True)[0][0][1]


@pytest.mark.parametrize("in_rate,out_rate", [(2_400_000, 48_000), (10_000_000, 48_000), (1_800_000, 48_000),
                                              (2_048_000, 48_000), (1_536_000, 48_000), (6_000_000, 48_000)])
def test_resampler_residue_form_edges(native, in_rate, out_rate):
    """The residue-form resampler (1/50, 3/625, 2/75, 3/128, 1/32, 1/125) against scipy's float64 resample_poly at lengths
    around its block and zero-extension edges: shorter than one decimation period, shorter than the filter, ragged ends."""
    from math import gcd

    from scipy import signal
    from wavecap_sdr_b200.dsp import fm

    g = gcd(in_rate, out_rate)
    up, down = out_rate // g, in_rate // g
    rng = np.random.default_rng(down)
    for n in (1, down - 1, down + 1, 8 * down - 3, 10 * down + 1, 20 * down + 7, 67 * down + down // 2):
        x = rng.standard_normal(n).astype(np.float32)
        got = fm.resample_poly(x, in_rate, out_rate)
        exp = signal.resample_poly(x.astype(np.float64), up, down).astype(np.float32)
        assert got.shape == exp.shape and got.dtype == np.float32
        scale = max(float(np.sqrt(np.mean(exp.astype(np.float64) ** 2))), 1e-3)
        assert float(np.sqrt(np.mean((got.astype(np.float64) - exp) ** 2))) / scale < 2e-6, (n, up, down)


def test_front_end_sum_of_squares_matches_the_separate_pass(native):
    """wc_front_run_ex's fused sum(out**2) (rms_normalize input of the default NBFM chain) against wc_sumsq of the same output."""
    import torch

    from wavecap_sdr_b200.dsp import _stages as S

    fs, n, n_chunks = 1_000_000, 50_001, 3
    rng = np.random.default_rng(21)
    x = torch.from_numpy(((rng.standard_normal(n * n_chunks) + 1j * rng.standard_normal(n * n_chunks)) * 0.1).astype(np.complex64)).cuda()
    modes = [S.MODE_NBFM, S.MODE_AM, S.MODE_SSB, S.MODE_WBFM]
    out, _, power, _, ss = S.front(x, S.FMT_CF32, n, n_chunks, modes, [12_000.0, 0.0, -40_000.0, 250_000.0], [0.0, 0.0, 1500.0, 0.0],
                                   fs, want_sumsq=True)
    ref = S.sumsq(out.reshape(len(modes) * n_chunks, n)).reshape(len(modes), n_chunks)
    torch.cuda.synchronize()
    assert torch.allclose(ss, ref, rtol=1e-6, atol=0.0)
    out2, _, power2, _ = S.front(x, S.FMT_CF32, n, n_chunks, modes, [12_000.0, 0.0, -40_000.0, 250_000.0], [0.0, 0.0, 1500.0, 0.0], fs)
    assert torch.equal(out, out2) and torch.allclose(power, power2, rtol=1e-12)


def test_one_call_plan_equals_the_stage_path_and_replays_its_graph(native):
    """wc_analog_run (one C call per batch, captured CUDA graph on repeats) against the stage-by-stage path it replaces:
    mixed chains (FM with and without IIR stages, AM with AGC, SSB without, a digital and an unknown mode), squelch, a
    chunk with non-finite IQ, int16 and complex64 input, numpy and CUDA-tensor input, four calls in a row."""
    import torch

    from wavecap_sdr_b200.capture import process_channels_batch

    fs, n, B = 48000, 4801, 3
    rng = np.random.default_rng(51)
    t = np.arange(n * B) / fs
    x = (0.3 * np.exp(1j * (2 * np.pi * 3000 * t + 2.5 * np.sin(2 * np.pi * 600 * t))) * (1 + 0.5 * np.sin(2 * np.pi * 300 * t))
         + 0.01 * (rng.standard_normal(n * B) + 1j * rng.standard_normal(n * B))).astype(np.complex64)
    x[n + 5] = np.inf                                            # chunk 1 is dropped
    specs = [dict(mode="nbfm", offset_hz=3000.0, enable_deemphasis=False, audio_rate=16000, squelch_db=-3.0),
             dict(mode="nbfm", offset_hz=3000.0, enable_deemphasis=False, audio_rate=16000, squelch_db=-60.0),
             dict(mode="wbfm", offset_hz=3000.0, audio_rate=48000),
             dict(mode="am", offset_hz=3000.0, enable_agc=True, audio_rate=16000),
             dict(mode="ssb", offset_hz=3000.0, enable_agc=False, audio_rate=8000),
             dict(mode="p25", offset_hz=0.0), dict(mode="bogus", offset_hz=10.0)]
    cfgs = [_cfg(native, **s) for s in specs]

    def same(a, b):
        for ra, rb in zip(a, b):
            for (aa, ma), (ab, mb) in zip(ra, rb):
                assert (aa is None) == (ab is None) and set(ma) == set(mb)
                for k in ma:
                    assert abs(ma[k] - mb[k]) < 1e-3, k
                if aa is not None:
                    assert aa.shape == ab.shape and rel_rms(aa, ab) < 1e-6

    ref = process_channels_batch(x, fs, cfgs, n_chunks=B, apply_squelch=True, use_plan=False)
    assert ref[1][0] == (None, {}) and ref[0][5][0] is None and "signal_power_db" in ref[0][5][1] and ref[0][6] == (None, {"rssi_db": ref[0][6][1]["rssi_db"]})
    assert not ref[0][0][0].any() and ref[0][1][0].any()         # squelch closed / open
    xd = torch.from_numpy(x).cuda()
    for rep in range(4):                                         # eager, capture, replay, replay
        same(process_channels_batch(x, fs, cfgs, n_chunks=B, apply_squelch=True), ref)
        got = process_channels_batch(xd, fs, cfgs, n_chunks=B, apply_squelch=True, return_device=True)
        same([[(a.cpu().numpy() if a is not None else None, m) for a, m in row] for row in got], ref)
    # on a non-default stream the plan captures its launch sequence on the second call with the same buffers and replays
    # the graph afterwards (the legacy default stream cannot be captured, so the calls above all ran eagerly)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        xs = torch.from_numpy(x).cuda()
        for rep in range(5):                                     # eager, capture + launch, replay x3
            got = process_channels_batch(xs, fs, cfgs, n_chunks=B, apply_squelch=True, return_device=True)
            side.synchronize()
            same([[(a.cpu().numpy() if a is not None else None, m) for a, m in row] for row in got], ref)
        # a single small chunk through the same plan object family (different n_chunks -> buffers re-reserved, graphs rebuilt)
        one = process_channels_batch(xs[:n], fs, cfgs, n_chunks=1, apply_squelch=True)
        same(one, ref[:1])
        for rep in range(3):
            same(process_channels_batch(xs[:n], fs, cfgs, n_chunks=1, apply_squelch=True), ref[:1])
    # int16 input
    q = np.stack([np.clip(x.real, -1, 1), np.clip(x.imag, -1, 1)], axis=1)
    q = (np.nan_to_num(q, posinf=0.0) * 20000).astype(np.int16)
    r0 = process_channels_batch(q, fs, cfgs[:5], n_chunks=B, in_fmt="cs16", use_plan=False)
    for rep in range(3):
        same(process_channels_batch(q, fs, cfgs[:5], n_chunks=B, in_fmt="cs16"), r0)
