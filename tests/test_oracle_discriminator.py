"""CPU: oracle/discriminator.py (voice-channel FM discriminator + DiscriminatorDemodulator restatement) against the
outputs of the live reference in tests/golden/p25_discriminator.npz, in both the literal and the portable mode."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import discriminator as od
from oracle.make_golden import discriminator_cases


@pytest.fixture(scope="module")
def gold():
    return np.load(golden_path("p25_discriminator.npz"))


def replay(g, name, chunk, portable):
    x = g[name + "_x"]
    o = od.DiscriminatorOracle(48000, portable=portable)
    last = 0.0
    aud, ds, cnt = [], [], []
    starts = list(range(0, len(x), chunk))
    for j, s0 in enumerate(starts):
        au, last = od.fm_discriminator(x[s0:s0 + chunk], last)
        if name == "disc_999" and j == len(starts) // 2:
            o.reset()
        a = o.demodulate(au.astype(np.float32))
        aud.append(au)
        ds.append(a)
        cnt.append(len(a))
    return np.concatenate(aud), np.concatenate(ds), np.array(cnt, np.int32), o


@pytest.mark.parametrize("case", discriminator_cases(), ids=lambda c: c[0])
def test_portable_oracle_matches_reference_golden(gold, case):
    name, chunk = case[0], case[1]
    au, dib, cnt, o = replay(gold, name, chunk, portable=True)
    assert np.array_equal(au, gold[name + "_audio"])
    assert np.array_equal(cnt, gold[name + "_counts"])
    assert np.array_equal(dib, gold[name + "_dibits"])
    st = gold[name + "_state"]
    assert abs(float(o.input_gain) - st[0]) <= 1e-6 * abs(st[0]) and abs(float(o.spread) - st[3]) <= 1e-5


@pytest.mark.parametrize("case", discriminator_cases()[:2], ids=lambda c: c[0])
def test_literal_oracle_matches_reference_golden(gold, case):
    """The literal mode uses np.convolve in float32 (host-dependent summation order); dibits still have to agree."""
    name, chunk = case[0], case[1]
    _, dib, cnt, _ = replay(gold, name, chunk, portable=False)
    assert np.array_equal(cnt, gold[name + "_counts"]) and np.array_equal(dib, gold[name + "_dibits"])


def test_state_kinds_documented():
    """The dtype flow the CUDA kernel hard-codes: every loop variable is float32 after the first symbol."""
    g = np.load(golden_path("p25_discriminator.npz"))
    o = od.DiscriminatorOracle(48000)
    o.demodulate(g["disc_2400_audio"][:2400].astype(np.float32))
    kinds = o.state_dtypes()
    assert kinds["clock"] == "float32" and kinds["fine"] == "float32" and kinds["dc"] == "float32"
    assert kinds["spread"] in ("float32", "float")  # `float` right after a clamp to the literal 1.6 / 2.4
