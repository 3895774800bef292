"""CPU: ChannelCalculator (dsp/channelizer.py:161-231) — the product's host arithmetic and the oracle restatement against
outputs of the reference itself (tests/golden/channel_calc.npz, generator oracle/make_golden.py:gen_channel_calc) —
and the oracle's channelize_samples (:234-268) against the reference's."""
import numpy as np

from conftest import golden_path, rel_rms
from oracle.channelizer import ChannelCalculatorOracle, channelize_samples
from oracle.make_golden import channel_calc_cases, channelize_samples_input


def test_channel_calculator_matches_reference():
    from wavecap_sdr_b200.dsp.channelizer import ChannelCalculator

    g = np.load(golden_path("channel_calc.npz"))
    for i, (center, fs, bw, targets) in enumerate(channel_calc_cases()):
        for cls in (ChannelCalculator, ChannelCalculatorOracle):
            calc = cls(center, fs, bw)
            assert calc.channel_count == int(g[f"count{i}"]) and calc.channel_count % 2 == 0
            idx = np.array([calc.get_channel_index(float(f)) for f in targets], dtype=np.int64)
            assert np.array_equal(idx, g[f"index{i}"]), cls.__name__
            cen = np.array([calc.get_channel_center_frequency(k) for k in range(calc.channel_count)])
            assert np.array_equal(cen, g[f"center{i}"]), cls.__name__
            # round trip on the bins that have a centre frequency of their own
            for k in range(calc.channel_count):
                if k != calc.channel_count // 2:
                    assert calc.get_channel_index(calc.get_channel_center_frequency(k)) == k


def test_oracle_channelize_samples_matches_reference():
    g = np.load(golden_path("channel_calc.npz"))
    x, fs, bw, center = channelize_samples_input()
    for j, target in enumerate((center - 75000.0, center + 200000.0, center)):
        y, rate = channelize_samples(x, fs, target, center, bw)
        assert rate == float(g[f"rate{j}"]) and y.dtype == np.complex64
        assert y.shape == g[f"chan{j}"].shape and rel_rms(y, g[f"chan{j}"]) < 1e-6
