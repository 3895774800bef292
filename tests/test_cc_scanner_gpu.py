"""GPU parity (through the C ABI): control-channel scanner vs the reference golden and the oracle.
Float outputs: dB values within 1e-4 dB (the float32 NCO product is the only float32 step); sync decisions, sample counts,
best channel and ranking identical."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import cc_scanner as oc

pytestmark = pytest.mark.gpu


def test_scan_all_matches_reference_golden(native):
    from wavecap_sdr_b200.cc_scanner import ControlChannelScanner

    g = np.load(golden_path("cc_scanner.npz"))
    x, center, freqs = oc.synth_band()
    sc = ControlChannelScanner(center_hz=center, sample_rate=1_200_000, control_channels=freqs)
    m = sc.scan_all(x)
    assert sorted(m) == sorted(g["rows"][:, 0].tolist())
    for row in g["rows"]:
        r = m[row[0]]
        assert abs(r.power_db - row[1]) < 1e-4 and abs(r.peak_power_db - row[2]) < 2e-4
        assert abs(r.noise_floor_db - row[3]) < 1e-4 and abs(r.snr_db - row[4]) < 2e-4
        assert r.sync_detected == bool(row[5]) and r.sample_count == int(row[6])
    assert sc.get_best_channel()[0] == float(g["best"])
    assert [f for f, _ in sc.get_channel_ranking()] == g["ranking"].tolist()
    assert sc.should_roam(freqs[3]) == float(g["best"]) and sc.should_roam(float(g["best"])) is None


def test_correlation_and_edges_vs_oracle(native):
    from wavecap_sdr_b200.cc_scanner import ControlChannelScanner, measure_offsets

    x, center, freqs = oc.synth_band(seed=72)
    offs = [f - center for f in freqs[:6]] + [0.0]
    mean, peak, corr, m = measure_offsets(x, 1_200_000, offs)
    taps = oc.scanner_taps(25)
    for i, off in enumerate(offs):
        y = oc.shift_decimate(x, off, 1_200_000, 25, taps)
        assert m == len(y)
        assert abs(mean[i] - np.mean(np.abs(y) ** 2)) <= 2e-6 * np.mean(np.abs(y) ** 2)
        assert abs(corr[i] - oc.sync_correlation(y)) < 1e-4
    # a block too short for the sync check, and an empty scan list
    sc = ControlChannelScanner(center_hz=center, sample_rate=1_200_000, control_channels=freqs[:2])
    short = sc.scan_all(x[:4000])
    assert all(not r.sync_detected and r.sample_count == 160 for r in short.values())
    assert ControlChannelScanner(center_hz=center, sample_rate=1_200_000, control_channels=[center + 9e6]).scan_all(x) == {}
