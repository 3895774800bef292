"""CPU: oracle/trellis.py (1/2-rate Viterbi + TSBK block decode restatement) against the live reference's outputs in
tests/golden/p25_trellis.npz."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import trellis as ot


@pytest.fixture(scope="module")
def gold():
    return np.load(golden_path("p25_trellis.npz"))


def test_trellis_decode_matches_reference_golden(gold):
    for t in range(len(gold["lens"])):
        n = int(gold["lens"][t])
        soft = gold["soft"][t, :n] if gold["has_soft"][t] else None
        d, m = ot.decode(gold["rx"][t, :n], soft)
        k = int(gold["dec_len"][t])
        assert len(d) == k and np.array_equal(d, gold["dec"][t, :k]) and m == int(gold["metric"][t]), t


def test_tsbk_block_decode_matches_reference_golden(gold):
    clean = 0
    for t in range(len(gold["tsbk_bits"])):
        bits, m = ot.tsbk_decode_bits(gold["tsbk_bits"][t])
        assert np.array_equal(bits, gold["tsbk_dec96"][t]) and m == int(gold["tsbk_metric"][t]), t
        clean += bool(np.array_equal(bits, gold["tsbk_tx96"][t]))
    assert clean > 150  # the injected errors are mostly corrected: real decoding is being compared


def test_encoder_roundtrip_and_interleaver():
    rng = np.random.default_rng(1)
    msg = np.concatenate([rng.integers(0, 4, 48), [0]])
    d, m = ot.decode(ot.encode(msg))
    assert np.array_equal(d, msg) and m == 0
    blk = rng.integers(0, 4, 98).astype(np.uint8)
    assert np.array_equal(ot.interleave(blk)[ot.DEINTERLEAVE], blk)
    assert sorted(ot.DEINTERLEAVE.tolist()) == list(range(98))
    lb, pr, op, mf, data = ot.tsbk_fields([1, 0, 1, 0, 1, 0, 1, 0] + [0] * 7 + [1] + [1] * 8 + [0] * 72)
    assert (lb, pr, op, mf, data[0]) == (1, 0, 0b101010, 1, 0xFF)
