"""GPU parity (through the C ABI): streaming FIR / FIR-decimate and the trunking fan-out vs the oracle and the
reference goldens. Float path: <= 1e-4 relative RMS (north star); measured ~1e-7."""
import numpy as np
import pytest
from scipy import signal

from conftest import golden_path, rel_rms
from oracle import ddc

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_fir_decimate_matches_reference_golden(native):
    from wavecap_sdr_b200.dsp.filters import fir_decimate, fir_filter_complex

    g = np.load(golden_path("ddc.npz"))
    x, taps = g["x"], g["taps"]
    zi = signal.lfilter_zi(taps, 1.0).astype(np.complex128) * x[0]
    ys, cuts = [], [0, 12000, 12077, 30000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        y, zi = fir_decimate(x[a:b], taps, 30, zi=zi)
        assert y.dtype == np.complex64 and zi.dtype == np.complex128
        ys.append(y)
    assert [len(y) for y in ys] == g["dec30_counts"].tolist()
    assert rel_rms(np.concatenate(ys), g["dec30"]) < TOL
    assert np.array_equal(zi, g["dec30_zi"])
    y, z = fir_filter_complex(x[:5000], taps[:73].copy(), None)
    assert rel_rms(y, g["filt73"]) < TOL and np.array_equal(z, g["filt73_zi"])
    e, ez = fir_filter_complex(np.zeros(0, np.complex64), taps)
    assert e.size == 0 and e.dtype == np.complex64 and ez.shape == (156,)


@pytest.mark.parametrize("flavor", ["control", "voice"])
def test_bank_matches_oracle(native, flavor):
    """12 channels from one 6 MS/s capture, three ragged calls (state, NCO phase and per-call decimation carry)."""
    from wavecap_sdr_b200.trunking import DDCBank

    fs, d1, d2 = 6_000_000, 30, 4
    offs = [412_500.0, -1_200_000.0, 0.0, 2_512_500.0, -37_500.0, 850_000.0, -2_900_000.0, 12_500.0, 1_000_000.0,
            -650_000.0, 2_000_000.0, -1_987_500.0]
    K = len(offs)
    x = ddc.synth_wideband(7, 300_000 + 123_457 + 90, fs, offs[:4])
    cuts = [0, 300_000, 300_000 + 123_457, len(x)]
    bank = DDCBank(K, fs, d1, d2, flavor=flavor)
    bank.set_offsets(offs)
    t1, t2 = bank.taps()
    o1, o2 = ddc.design(d1, d2)
    assert np.max(np.abs(t1 - o1)) < 1e-15 and np.max(np.abs(t2 - o2)) < 1e-15   # in-library Kaiser design == scipy
    got = [bank.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    cls = ddc.ControlChannelDDC if flavor == "control" else ddc.VoiceDDC
    for k in range(K):
        o = cls(fs, d1, d2, offs[k])
        for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            exp = o.process(x[a:b])
            assert got[i].shape[1] == len(exp)
            assert got[i].dtype == (np.complex64 if flavor == "control" else np.complex128)
            err = rel_rms(got[i][k], exp)
            assert err < TOL, f"{flavor} channel {k} call {i}: rel-RMS {err}"


def test_offset_change_restarts_phase_and_single_stage(native):
    from wavecap_sdr_b200.trunking import DDCBank

    fs = 2_400_000
    x = ddc.synth_wideband(9, 60_000, fs, [300_000.0])
    bank = DDCBank(2, fs, 25, 1, flavor="control")
    o = [ddc.ControlChannelDDC(fs, 25, 1, 300_000.0), ddc.ControlChannelDDC(fs, 25, 1, -100_000.0)]
    bank.set_offsets([300_000.0, -100_000.0])
    y = bank.process(x[:30_000])
    for k in range(2):
        assert rel_rms(y[k], o[k].process(x[:30_000])) < TOL
    bank.set_offsets([300_000.0, 450_000.0])          # channel 1 retunes: its NCO phase restarts, filter state stays
    o[1].offset = 450_000.0
    y = bank.process(x[30_000:])
    for k in range(2):
        assert rel_rms(y[k], o[k].process(x[30_000:])) < TOL
