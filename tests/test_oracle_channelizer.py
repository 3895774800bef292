"""CPU: the oracle restatement of dsp/channelizer.py is pinned to outputs of the reference itself
(tests/golden/channelizer_c5.npz, made by oracle/make_golden.py) and, when /root/reference is
present, to the live reference."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import refenv
from oracle.channelizer import ChannelizerOracle, channelize_fm, design_arms


def test_oracle_matches_golden_bit_exact():
    g = np.load(golden_path("channelizer_c5.npz"))
    for variant in ("process", "process_vectorized"):
        o = ChannelizerOracle(float(g["fs"]), int(g["bw"]))
        cut = int(g["cut"])
        y1 = getattr(o, variant)(g["x"][:cut])
        y2 = getattr(o, variant)(g["x"][cut:])
        assert np.array_equal(y1, g["frames1"]) and np.array_equal(y2, g["frames2"])
        assert np.array_equal(o.arm_history, g["arm_history"])
        assert np.array_equal(o.arms, g["arms"])
    assert np.array_equal(channelize_fm(g["frames2"], int(g["demod_rate"])), g["fm2"])


def test_frame_count_rule_and_dropped_tail():
    o = ChannelizerOracle(125_000_000, 488281)
    x = np.zeros(10240, np.complex64)
    assert o.process(x).shape[0] == 79
    o.reset()
    assert o.process(x[:5120]).shape[0] + o.process(x[5120:]).shape[0] == 78
    assert o.process(x[:255]).shape == (0, 256)


def test_channel_count_made_even():
    m, arms = design_arms(8_000_000, 25000, 9)
    assert m == 320 and arms.shape == (320, 9) and arms[-1, -1] == 0.0
    m, _ = design_arms(1_000_000, 3003, 9)  # int(333.0) -> 332
    assert m == 332


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_oracle_matches_live_reference():
    refenv.load()
    from wavecapsdr.dsp.channelizer import PolyphaseChannelizer

    rng = np.random.default_rng(123)
    for fs, bw, n in ((125_000_000, 488281, 9000), (8_000_000, 25000, 5000)):
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
        ref, o = PolyphaseChannelizer(fs, bw), ChannelizerOracle(fs, bw)
        for a, b in ((0, n // 3), (n // 3, n)):
            r = ref.process(x[a:b])
            y = o.process_vectorized(x[a:b])
            assert len(r) == y.shape[0]
            assert np.array_equal(np.array(r), y)
        assert np.array_equal(ref.arm_history, o.arm_history)
