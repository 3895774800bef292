"""GPU parity (through the C ABI): batched 1/2-rate trellis decoding and the TSBK block decode (csrc/p25frame.cu) vs the
reference goldens and the oracle — decoded dibits / bits, lengths and error metrics identical."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import trellis as ot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(golden_path("p25_trellis.npz"))


def test_trellis_batch_matches_reference_golden(native, gold):
    from wavecap_sdr_b200.dsp.fec.trellis import trellis_decode, trellis_decode_batch

    hard = np.nonzero(gold["has_soft"] == 0)[0]
    soft = np.nonzero(gold["has_soft"] == 1)[0]
    for sel, sv in ((hard, None), (soft, gold["soft"][soft])):
        out, n_out, met = trellis_decode_batch(gold["rx"][sel], sv, lengths=gold["lens"][sel])
        assert np.array_equal(n_out, gold["dec_len"][sel])
        assert np.array_equal(met, gold["metric"][sel])
        for j, t in enumerate(sel):
            k = int(n_out[j])
            assert np.array_equal(out[j, :k], gold["dec"][t, :k]), t
    # scalar call surface (trellis.py:312-329), incl. an odd length and an empty block
    for t in (0, 5, 7):
        n = int(gold["lens"][t])
        sv = gold["soft"][t, :n] if gold["has_soft"][t] else None
        d, m = trellis_decode(gold["rx"][t, :n], sv)
        k = int(gold["dec_len"][t])
        assert np.array_equal(d, gold["dec"][t, :k]) and m == int(gold["metric"][t])
    d, m = trellis_decode(np.zeros(0, np.uint8))
    assert d.size == 0


def test_tsbk_blocks_match_reference_golden(native, gold):
    from wavecap_sdr_b200.decoders.p25 import P25TrellisDecoder
    from wavecap_sdr_b200.dsp.fec.trellis import tsbk_decode_batch

    bits96, met, fields, data = tsbk_decode_batch(gold["tsbk_bits"])
    assert np.array_equal(bits96, gold["tsbk_dec96"])
    assert np.array_equal(met, gold["tsbk_metric"])
    for t in range(0, 200, 17):
        lb, pr, op, mf, payload = ot.tsbk_fields(gold["tsbk_dec96"][t])
        assert tuple(fields[t]) == (lb, pr, op, mf) and bytes(data[t]) == payload
    # drop-in class (decoders/p25.py:1348-1393) on the deinterleaved dibits
    blk = ((gold["tsbk_bits"][3, 0::2] << 1) | gold["tsbk_bits"][3, 1::2]).astype(np.uint8)
    d, e = P25TrellisDecoder().decode(blk[ot.DEINTERLEAVE])
    ref = gold["tsbk_dec96"][3]
    assert e == int(gold["tsbk_metric"][3]) and np.array_equal((d >> 1) & 1, ref[0::2]) and np.array_equal(d & 1, ref[1::2])
    assert P25TrellisDecoder().decode(np.zeros(3, np.uint8)) == (None, -1)


def test_long_blocks_and_random_words_vs_oracle(native):
    from wavecap_sdr_b200.dsp.fec.trellis import trellis_decode_batch

    rng = np.random.default_rng(77)
    rx = rng.integers(0, 4, (40, 700)).astype(np.uint8)          # pure noise: every tie-break path gets exercised
    lens = rng.integers(0, 701, 40).astype(np.int32)
    out, n_out, met = trellis_decode_batch(rx, None, lengths=lens)
    for b in range(40):
        d, m = ot.decode(rx[b, :lens[b]])
        assert int(n_out[b]) == len(d) and np.array_equal(out[b, :len(d)], d) and int(met[b]) == m, b
