"""CPU: the BCH(63,16,23) and P25 framer restatements (oracle/bch.py, oracle/p25_framer.py) against the golden
outputs of the live reference (tests/golden/p25_framer.npz, oracle/make_golden.py:gen_p25_framer) — and against the
reference itself when /root/reference is present."""
import json
import os

import numpy as np
import pytest

from oracle import bch as ob
from oracle import p25_framer as of

GOLD = os.path.join(os.path.dirname(__file__), "golden", "p25_framer.npz")
TS = 1_700_000_000_000


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def unpack(meta, bits):
    out, off = [], 0
    for duid, nac, ts, corrected, nbits in meta:
        out.append((int(duid), int(nac), int(ts), bytes(bits[off:off + int(nbits)]), int(corrected)))
        off += int(nbits)
    return out


def oracle_batch(soft, dib, bounds):
    fr = of.FramerOracle()
    fr.ref_ts, fr.ts_base = TS, 0
    log = []
    for a, b in bounds:
        try:
            log.append(fr.process_batch(soft[a:b], dib[a:b]))
        except AssertionError as e:
            log.append(str(e))
    return [(d, n, t, bytes(b), c) for d, n, t, b, c in fr.out], log


def oracle_stream(soft, dib, max_errors=40):
    fr = of.FramerOracle()
    fr.ref_ts, fr.ts_base = TS, 0
    log = []
    for i in range(len(dib)):
        try:
            if fr.process_stream(soft[i:i + 1], dib[i:i + 1], 1):
                log.append(i)
        except AssertionError as e:
            log.append([i, str(e)])
            if len(log) > max_errors:
                break
    return [(d, n, t, bytes(b), c) for d, n, t, b, c in fr.out], log


def test_bch_known_answers(gold):
    # the reference's own known answers (tests/test_p25_bch.py:38-45): the all-zero word decodes to (0, 0)
    assert ob.bch_decode(np.zeros(63, dtype=np.uint8)) == (0, 0)
    for cw, tr, d, e in zip(gold["bch_cw"], gold["bch_tracked"], gold["bch_data"], gold["bch_errors"]):
        assert ob.bch_decode(cw, int(tr) if tr else None) == (int(d), int(e))


def test_bch_encoder_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(50):
        d = int(rng.integers(0, 65536))
        c = ob.bch_encode(d)
        assert not any(ob.syndromes(c))
        c[rng.choice(63, 11, replace=False)] ^= 1
        assert ob.bch_decode(c) == (d, 11)


def test_framer_streams_batch_and_stream(gold):
    for name in json.loads(str(gold["stream_names"])):
        dib, soft, chunk = gold[f"{name}_dibits"], gold[f"{name}_soft"], int(gold[f"{name}_chunk"])
        bounds = [(s, min(s + chunk, len(dib))) for s in range(0, len(dib), chunk)]
        msgs, log = oracle_batch(soft, dib, bounds)
        assert log == json.loads(str(gold[f"{name}_batch_log"])), name
        assert msgs == unpack(gold[f"{name}_batch_meta"], gold[f"{name}_batch_bits"]), name
        msgs, log = oracle_stream(soft, dib)
        assert log == json.loads(str(gold[f"{name}_stream_log"])), name
        assert msgs == unpack(gold[f"{name}_stream_meta"], gold[f"{name}_stream_bits"]), name


def test_framer_on_reference_demodulator_output(gold):
    dib, soft = gold["e2e_dibits"], gold["e2e_soft"]
    edges = np.concatenate([[0], np.cumsum(gold["e2e_counts"])])
    msgs, log = oracle_batch(soft, dib, list(zip(edges[:-1], edges[1:])))
    assert log == json.loads(str(gold["e2e_batch_log"]))
    assert msgs == unpack(gold["e2e_batch_meta"], gold["e2e_batch_bits"])
    msgs, log = oracle_stream(soft, dib)
    assert log == json.loads(str(gold["e2e_stream_log"]))
    assert msgs == unpack(gold["e2e_stream_meta"], gold["e2e_stream_bits"])
    assert len(msgs) >= 8  # decoded frames are really being compared


def test_live_reference_bch_if_present():
    from oracle import refenv

    if not refenv.available():
        pytest.skip("reference not present")
    refenv.load()
    from wavecapsdr.dsp.fec.bch import bch_decode

    rng = np.random.default_rng(99)
    for t in range(200):
        d = int(rng.integers(0, 65536))
        c = ob.bch_encode(d).copy()
        c[rng.choice(63, int(rng.integers(0, 15)), replace=False)] ^= 1
        tr = [None, 0x5A5][t % 2]
        a = bch_decode(c, tr)
        assert (int(a[0]), int(a[1])) == ob.bch_decode(c, tr)
