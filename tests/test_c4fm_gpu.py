"""GPU parity (through the C ABI): batched C4FM symbol recovery vs the oracle and the reference goldens.
Bar: dibits, symbol counts and sync-event counts identical; soft symbols within 2e-6 absolute (the
reference's float32 arctan2 is a SIMD approximation that is not correctly rounded; ours is)."""
import numpy as np
import pytest

from conftest import golden_path
from oracle.c4fm import C4FMOracle, modulate_c4fm, random_frames
from oracle.make_golden import c4fm_cases

pytestmark = pytest.mark.gpu


def run_bank(bank, xs, chunk):
    """xs: [C][n]; replay the chunk sequence; returns per-channel (dibits, soft, counts)."""
    C = xs.shape[0]
    ds, ss, cs = [[] for _ in range(C)], [[] for _ in range(C)], [[] for _ in range(C)]
    for s in range(0, xs.shape[1], chunk):
        d, so, cnt = bank.demodulate(xs[:, s:s + chunk])
        for c in range(C):
            n = int(cnt[c])
            ds[c].append(d[c, :n].copy())
            ss[c].append(so[c, :n].copy())
            cs[c].append(n)
    return [np.concatenate(v) for v in ds], [np.concatenate(v) for v in ss], [np.array(v, np.int32) for v in cs]


def run_oracle(fs, x, chunk):
    o = C4FMOracle(sample_rate=fs, portable=True)
    ds, ss, cs = [], [], []
    for s in range(0, len(x), chunk):
        d, so = o.demodulate(x[s:s + chunk])
        ds.append(d)
        ss.append(so)
        cs.append(len(d))
    return np.concatenate(ds), np.concatenate(ss), np.array(cs, np.int32), o


@pytest.mark.parametrize("case", c4fm_cases(), ids=lambda c: c[0])
def test_matches_reference_golden(native, case):
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator

    name, fs, chunk = case[0], case[1], case[2]
    g = np.load(golden_path("p25_c4fm.npz"))
    x = g[name + "_x"]
    dm = C4FMDemodulator(sample_rate=fs)
    ds, ss, cs = [], [], []
    for s in range(0, len(x), chunk):
        d, so = dm.demodulate(x[s:s + chunk])
        assert d.dtype == np.uint8 and so.dtype == np.float32
        ds.append(d)
        ss.append(so)
        cs.append(len(d))
    assert np.array_equal(np.array(cs, np.int32), g[name + "_counts"])
    d, so = np.concatenate(ds), np.concatenate(ss)
    assert np.array_equal(d, g[name + "_dibits"]), f"{int((d != g[name + '_dibits']).sum())} dibit mismatches"
    assert np.max(np.abs(so - g[name + "_soft"])) <= 2e-6
    assert dm._sync_count == int(g[name + "_sync_count"])


@pytest.mark.parametrize("fs,chunk,C", [(48000, 2400, 8), (50000, 2500, 8), (48000, 72000, 8),
                                        (48000, 2400, 64), (50000, 2500, 64), (48000, 72000, 64), (50000, 75000, 64)],
                         ids=lambda v: str(v))
def test_bank_matches_oracle(native, fs, chunk, C):
    """C channels with different CFO / timing / SNR / content advanced together; C = 64 is BASELINE.json configs[3]
    (generic path: 50 ms chunks; control-channel path: 1.5 s chunks, longer than the 65 536-sample history buffer, so the
    half-shift happens inside a call and the run spans several calls)."""
    from conftest import parity_note
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMBank

    nfr = 5 if chunk < 10000 else (30 if C <= 8 else 50)
    xs = []
    for c in range(C):
        rng = np.random.default_rng(400 + c)
        dib = random_frames(rng, n_frames=nfr, payload=150, gap=40)
        xs.append(modulate_c4fm(dib, fs, snr_db=20.0 + (10.0 * c) / max(C - 1, 1), cfo_hz=-200.0 + (400.0 * c) / max(C - 1, 1),
                                timing=(0.11 * c) % 1.0, seed=400 + c))
    n = min(len(x) for x in xs)
    xs = np.stack([x[:n] for x in xs])
    bank = C4FMBank(C, sample_rate=fs)
    gd, gs, gc = run_bank(bank, xs, chunk)
    worst, total, syncs = 0.0, 0, 0
    for c in range(C):
        d, so, cnt, o = run_oracle(fs, xs[c], chunk)
        assert np.array_equal(gc[c], cnt), f"channel {c}: symbol counts differ"
        assert np.array_equal(gd[c], d), f"channel {c}: {int((gd[c] != d).sum())} dibit mismatches of {len(d)}"
        worst = max(worst, float(np.max(np.abs(gs[c] - so))))
        total += len(d)
        syncs += o.sync_count
        st = bank.state(c)
        assert st["sync_count"] == o.sync_count and st["fine_sync"] == o.fine
        assert abs(st["pll"] - o.pll) < 1e-6 and abs(st["gain"] - o.gain) < 1e-6
        assert abs(st["sample_point"] - o.sample_point) < 1e-6 and st["buffer_pointer"] == o.buf_ptr
    assert worst <= 2e-6, worst
    parity_note(f"c4fm bank C={C} fs={fs} chunk={chunk}: {total} dibits identical to the oracle, {syncs} sync events, "
                f"{-(-n // chunk)} calls of <= {chunk} samples, max |soft - oracle| = {worst:.1e}")


def test_reset_and_empty(native):
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator

    rng = np.random.default_rng(9)
    x = modulate_c4fm(random_frames(rng, n_frames=3), 48000, seed=9)
    dm = C4FMDemodulator(sample_rate=48000)
    d0, s0 = dm.demodulate(x)
    e, es = dm.demodulate(np.zeros(0, np.complex64))
    assert e.size == 0 and es.size == 0 and e.dtype == np.uint8 and es.dtype == np.float32
    dm.reset()
    d1, s1 = dm.demodulate(x)
    assert np.array_equal(d0, d1) and np.array_equal(s0, s1)
    o = C4FMOracle(sample_rate=48000, portable=True)
    d2, _ = o.demodulate(x)
    assert np.array_equal(d0, d2)


def test_noise_only_and_short_chunks(native):
    """pure noise (no sync) and chunks shorter than the filter history / one symbol."""
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator

    rng = np.random.default_rng(11)
    x = ((rng.standard_normal(6000) + 1j * rng.standard_normal(6000)) * 0.1).astype(np.complex64)
    dm, o = C4FMDemodulator(sample_rate=48000), C4FMOracle(sample_rate=48000, portable=True)
    for chunk in (7, 130, 3, 500, 1, 2000, 3359):
        if len(x) < chunk:
            break
        c, x = x[:chunk], x[chunk:]
        d, s = dm.demodulate(c)
        do, so = o.demodulate(c)
        assert np.array_equal(d, do) and (len(s) == 0 or np.max(np.abs(s - so)) <= 2e-6)


def test_benchmark_helper_classes(native):
    """_FMDemodulator / _Interpolator / _SoftSyncDetector (the names backend/benchmark_dsp.py:17-114 times) vs the oracle."""
    from oracle.c4fm import DiffDemodOracle, SoftSyncOracle, TAPS, tap_row
    from wavecap_sdr_b200.dsp.p25.c4fm import _FMDemodulator, _Interpolator, _SoftSyncDetector

    rng = np.random.default_rng(21)
    i = (rng.standard_normal(5000) * 0.5).astype(np.float32)
    q = (rng.standard_normal(5000) * 0.5).astype(np.float32)
    for kw in ({"symbol_delay": 10}, {"samples_per_symbol": 50000 / 4800}):
        d = _FMDemodulator(**kw)
        o = DiffDemodOracle(list(kw.values())[0], portable=True)
        for a, b in ((0, 1000), (1000, 1007), (1007, 5000)):
            got, exp = d.demodulate(i[a:b], q[a:b]), o.demodulate(i[a:b], q[a:b])
            assert got.dtype == np.float32 and np.max(np.abs(got - exp)) <= 5e-7
        d.reset()
        o.reset()
        assert np.max(np.abs(d.demodulate(i[:100], q[:100]) - o.demodulate(i[:100], q[:100]))) <= 5e-7
    s = rng.standard_normal(600).astype(np.float32)
    it = _Interpolator()
    offs = rng.integers(-3, 598, 200)
    mus = rng.random(200)
    got = it.filter_batch(s, offs, mus)
    for k in range(200):
        row = tap_row(float(mus[k]))
        exp = 0.0
        for t in range(8):
            j = int(offs[k]) + t
            if 0 <= j < len(s):
                exp += float(np.float32(s[j] * TAPS[row][t]))
        assert abs(got[k] - exp) < 1e-12
    assert abs(it.filter(s, 5, 0.0) - float(s[8])) < 1e-6 and abs(it.filter(s, 5, 1.0) - float(s[9])) < 1e-6
    det, od = _SoftSyncDetector(), SoftSyncOracle()
    soft = (rng.standard_normal(300) * 3).astype(np.float32)
    got = np.concatenate([det.process_block(soft[:7]), det.process_block(soft[7:250]), [det.process(float(v)) for v in soft[250:]]])
    exp = np.array([od.process(v) for v in soft])
    assert np.max(np.abs(got - exp)) < 1e-9


def test_discriminator_entry_matches_reference_golden(native):
    """C4FMDemodulator.demodulate_discriminator (c4fm.py:2817-2992) through wc_c4fm_demod_disc: dibits and counts
    identical to the live reference's outputs, soft symbols within 2e-6, final timing state equal."""
    from oracle.make_golden import c4fm_disc_cases
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator

    g = np.load(golden_path("p25_c4fm_disc.npz"))
    for name, fs, chunk, seed, dt in c4fm_disc_cases():
        au = g[name + "_audio"]
        d = C4FMDemodulator(sample_rate=fs)
        ds, ss, cnt = [], [], []
        starts = list(range(0, len(au), chunk))
        for j, s0 in enumerate(starts):
            if name.endswith("ragged") and j == len(starts) // 2:
                d.reset()
            a, b = d.demodulate_discriminator(au[s0:s0 + chunk])
            ds.append(a)
            ss.append(b)
            cnt.append(len(a))
        assert np.array_equal(np.array(cnt, np.int32), g[name + "_counts"]), name
        assert np.array_equal(np.concatenate(ds), g[name + "_dibits"]), name
        assert np.max(np.abs(np.concatenate(ss) - g[name + "_soft"])) <= 2e-6, name
        st = g[name + "_state"]
        s = d._bank.state(0)
        assert bool(st[0]) == s["fine_sync"] and abs(st[1] - s["sample_point"]) < 1e-9 and st[2] == s["gain"], name
    assert d.demodulate_discriminator(np.zeros(0))[0].size == 0


def test_discriminator_bank_vs_oracle(native):
    """8 channels with different signals through the batched discriminator entry vs one oracle object per channel."""
    from oracle.c4fm import discriminator_audio
    from wavecap_sdr_b200.dsp.p25.c4fm import C4FMBank

    fs, C, chunk = 48000, 8, 3000
    aus = []
    for c in range(C):
        rng = np.random.default_rng(300 + c)
        x = modulate_c4fm(random_frames(rng, n_frames=6, payload=150, gap=40), fs, snr_db=22.0 + c, cfo_hz=30.0 * (c - 4),
                          timing=0.1 * c, seed=300 + c)
        aus.append(discriminator_audio(x))
    n = min(len(a) for a in aus)
    A = np.array([a[:n] for a in aus])
    bank = C4FMBank(C, fs)
    oracles = [C4FMOracle(sample_rate=fs) for _ in range(C)]
    for s0 in range(0, n, chunk):
        dib, soft, cnt = bank.demodulate_discriminator(A[:, s0:s0 + chunk])
        for c in range(C):
            od, os_ = oracles[c].demodulate_discriminator(A[c, s0:s0 + chunk])
            k = int(cnt[c])
            assert k == len(od), (c, s0)
            assert np.array_equal(dib[c, :k], od), (c, s0)
            assert np.max(np.abs(soft[c, :k] - os_)) <= 2e-6 if k else True
