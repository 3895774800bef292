"""CPU: oracle/analog.py (restatement of capture.py / dsp/fm.py / dsp/am.py / dsp/agc.py /
dsp/filters.py) is pinned bit-exact to outputs of the reference itself (tests/golden/analog.npz)."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import analog as oa
from oracle import refenv


@pytest.fixture(scope="module")
def g():
    return np.load(golden_path("analog.npz"))


def test_c1_wbfm_chunks_bit_exact(g):
    for i in range(2):
        x = oa.synth_c1(seed=1, n=120_000, t0=i * 120_000)
        a, m = oa.process_channel_dsp_stateless(x, 2_400_000, oa.OracleChannelConfig(mode="wbfm", offset_hz=200000.0))
        assert np.array_equal(a, g[f"c1_audio{i}"])
        assert np.allclose([m["rssi_db"], m["signal_power_db"]], g[f"c1_metrics{i}"], rtol=0, atol=1e-6)


def test_c2_nbfm_16_channels_bit_exact(g):
    q, offs = oa.synth_c2(seed=2, n=500_000, keyed_off=(3, 12))
    xc = oa.cs16_to_cf32(q)
    assert np.array_equal(np.array(offs, dtype=np.float64), g["c2_offsets"])
    for k in (0, 3, 7, 12, 15):
        cfg = oa.OracleChannelConfig(mode="nbfm", offset_hz=float(offs[k]), enable_deemphasis=False,
                                     enable_mpx_filter=False)
        a, m = oa.process_channel_dsp_stateless(xc, 10_000_000, cfg)
        assert np.array_equal(a, g["c2_audio"][k])
        assert np.allclose([m["rssi_db"], m["signal_power_db"]], g["c2_metrics"][k], rtol=0, atol=1e-6)


def test_stage_functions_bit_exact(g):
    xa, fs = g["am_x"], 48000
    f = np.real(xa).astype(np.float32)
    assert np.array_equal(oa.am_demod(xa, fs, 16000), g["am_audio"])
    assert np.array_equal(oa.am_demod(xa, fs, 16000, enable_agc=False), g["am_audio_noagc"])
    assert np.array_equal(oa.ssb_demod(xa, fs, 16000), g["ssb_audio"])
    assert np.array_equal(oa.ssb_demod(xa, fs, 16000, mode="lsb", enable_agc=False), g["ssb_audio_lsb"])
    assert np.array_equal(oa.apply_agc(f, fs), g["agc"])
    assert np.array_equal(oa.deemphasis_filter(f, fs), g["deemph"])
    assert np.array_equal(oa.lpf_audio(f, fs, 5000), g["lpf"])
    assert np.array_equal(oa.resample_poly(f, 48000, 14400), g["resamp_3_10"])
    assert np.array_equal(oa.resample_poly(f[:1000], 8000, 48000), g["resamp_up"])
    assert np.array_equal(oa.quadrature_demod(xa, fs), g["quad"])
    assert np.array_equal(oa.freq_shift(xa, 1234.4, fs), g["fshift"])
    assert np.array_equal(oa.am_freq_shift(xa, 1500.0, fs), g["am_fshift"])


def test_reference_property_tests_hold_for_the_oracle():
    # tests/unit/test_dsp_core.py:70-77 (rms target), :107-128 (discriminator of DC / tone),
    # :152-162 (resample lengths); tests/unit/test_fm_demod.py:27-47 (|audio| <= 1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4000).astype(np.float32)
    y = oa.rms_normalize(x, 0.18)
    assert abs(float(np.sqrt(np.mean(y ** 2))) - 0.18) < 1e-3
    dc = np.ones(1000, dtype=np.complex64)
    assert np.allclose(oa.quadrature_demod(dc, 48000), 0.0, atol=1e-6)
    tone = np.exp(2j * np.pi * 1000 / 48000 * np.arange(1000)).astype(np.complex64)
    d = oa.quadrature_demod(tone, 48000)
    assert np.allclose(d[1:], d[1], atol=1e-5)
    assert oa.resample_poly(np.zeros(1000, np.float32), 48000, 24000).size == 500
    assert oa.resample_poly(np.zeros(1000, np.float32), 24000, 48000).size == 2000
    iq = oa.synth_c1(n=24000)
    assert np.max(np.abs(oa.wbfm_demod(iq, 2_400_000))) <= 1.0


def test_nonfinite_and_empty_inputs():
    cfg = oa.OracleChannelConfig(mode="nbfm")
    assert oa.process_channel_dsp_stateless(np.zeros(0, np.complex64), 48000, cfg) == (None, {})
    bad = np.ones(100, np.complex64)
    bad[7] = np.nan
    assert oa.process_channel_dsp_stateless(bad, 48000, cfg) == (None, {})


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_oracle_matches_live_reference_on_fresh_input():
    refenv.load()
    import wavecapsdr.capture as rc

    x = oa.synth_c1(seed=77, n=24_000)
    for mode, kw in (("wbfm", {}), ("nbfm", {"enable_deemphasis": False}), ("raw", {}), ("p25", {})):
        cfg = rc.ChannelConfig(id="a", capture_id="c", mode=mode, offset_hz=-150000.0, **kw)
        a_r, m_r = rc._process_channel_dsp_stateless(x, 2_400_000, cfg)
        a_o, m_o = oa.process_channel_dsp_stateless(x, 2_400_000, oa.OracleChannelConfig(mode=mode, offset_hz=-150000.0, **kw))
        assert m_r == m_o
        assert (a_r is None) == (a_o is None)
        if a_r is not None:
            assert np.array_equal(a_r, a_o)


def test_lfilter_is_the_separately_rounded_df2t_recursion():
    """csrc/analog.cu:iir_seq_kernel replays scipy's lfilter with every product and sum rounded on its own (no FMA).
    That is what the scipy build in this image executes: a plain-Python float64 replay is bit-equal to lfilter even for
    the ill-conditioned order-10 band-pass of ssb_demod (dsp/filters.py:177-217)."""
    from scipy import signal

    rng = np.random.default_rng(4)
    x = rng.standard_normal(6000).astype(np.float32).astype(np.float64)
    for b, a in (signal.butter(5, [300 / 24000, 3000 / 24000], btype="band"), signal.butter(5, 3000 / 5e6),
                 signal.butter(5, 100 / 24000, btype="high")):
        ref = signal.lfilter(b, a, x)
        K = len(a) - 1
        bl, al = [float(v / a[0]) for v in b], [float(v / a[0]) for v in a]
        z = [0.0] * K
        out = np.empty_like(x)
        for n, xn in enumerate(x.tolist()):
            yn = z[0] + bl[0] * xn
            for i in range(K - 1):
                z[i] = (z[i + 1] + xn * bl[i + 1]) - yn * al[i + 1]
            z[K - 1] = xn * bl[K] - yn * al[K]
            out[n] = yn
        assert np.array_equal(out, ref)
