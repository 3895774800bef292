"""CPU: the oracle restatement of decoders/p25.py CQPSKDemodulator is pinned to outputs of the
reference itself (tests/golden/p25_cqpsk.npz) and, when /root/reference is present, to the live
reference: dibits, symbol counts and the loop state must be identical."""
import warnings

import numpy as np
import pytest

from conftest import golden_path, require_golden_host
from oracle import refenv
from oracle.cqpsk import CQPSKOracle, mmse_table, modulate_cqpsk
from oracle.make_golden import cqpsk_cases

warnings.filterwarnings("ignore")


def replay(demod, x, chunk):
    ds, cnt = [], []
    for s in range(0, len(x), chunk):
        a = demod.demodulate(x[s:s + chunk])
        ds.append(a)
        cnt.append(len(a))
    return np.concatenate(ds), np.array(cnt, dtype=np.int32)


@pytest.mark.parametrize("portable", [False, True], ids=["literal", "portable"])
@pytest.mark.parametrize("case", cqpsk_cases(), ids=lambda c: c[0])
def test_oracle_matches_golden(case, portable):
    name, fs, sr, chunk = case[:4]
    if not portable:
        require_golden_host()
    g = np.load(golden_path("p25_cqpsk.npz"))
    o = CQPSKOracle(sample_rate=fs, symbol_rate=sr, portable=portable)
    d, c = replay(o, g[name + "_x"], chunk)
    assert np.array_equal(c, g[name + "_counts"])
    assert np.array_equal(d, g[name + "_dibits"])
    st = g[name + "_state"]
    # float32 SIMD arctan2 / pairwise mean may differ in the last bit between CPUs: loop state within 1e-5
    assert abs(float(o.freq_offset) - st[0]) < 1e-5 and abs(float(o.clock) - st[2]) < 1e-4
    assert abs(float(o.agc_gain) - st[4]) < 1e-5


def test_golden_signal_is_decodable():
    g = np.load(golden_path("p25_cqpsk.npz"))
    tx, rx = g["cqpsk_48k_72000_tx"], g["cqpsk_48k_72000_dibits"]
    best = 0.0
    for lag in range(0, 20):
        n = min(len(rx) - lag, len(tx))
        best = max(best, float(np.mean(rx[lag:lag + n][500:] == tx[:n][500:])))
    assert best > 0.95


def test_mmse_table_and_empty():
    t = mmse_table()
    assert t.shape == (129, 8) and t.dtype == np.float32 and np.allclose(t.sum(axis=1), 1.0, atol=1e-6)
    assert t[0, 3] == 1.0
    assert CQPSKOracle(48000).demodulate(np.zeros(0, np.complex64)).size == 0


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_oracle_matches_live_reference():
    refenv.load()
    from wavecapsdr.decoders.p25 import CQPSKDemodulator

    rng = np.random.default_rng(8)
    x = modulate_cqpsk(rng.integers(0, 4, 1500), 48000, 4800, snr_db=20, cfo_hz=55, timing=0.45, seed=8)
    r, o = CQPSKDemodulator(sample_rate=48000), CQPSKOracle(sample_rate=48000)
    d1, c1 = replay(r, x, 3100)
    d2, c2 = replay(o, x, 3100)
    assert np.array_equal(c1, c2) and np.array_equal(d1, d2)
    assert float(r._freq_offset) == float(o.freq_offset) and float(r._symbol_clock) == float(o.clock)
