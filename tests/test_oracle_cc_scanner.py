"""CPU: oracle/cc_scanner.py vs ControlChannelScanner.scan_all of the live reference (tests/golden/cc_scanner.npz)."""
import numpy as np

from conftest import golden_path
from oracle import cc_scanner as oc


def test_measure_matches_reference_golden():
    g = np.load(golden_path("cc_scanner.npz"))
    x, center, freqs = oc.synth_band()
    assert abs(np.sum(np.abs(x.astype(np.complex128)) ** 2) - float(g["x_checksum"])) <= 1e-6 * float(g["x_checksum"])
    assert len(g["rows"]) == len(freqs) - 1  # the last candidate is outside the capture bandwidth: skipped by scan_all
    for row in g["rows"]:
        o = oc.measure(x, 1_200_000, center, row[0])
        assert abs(o["power_db"] - row[1]) < 1e-4 and abs(o["peak_power_db"] - row[2]) < 1e-4
        assert abs(o["noise_floor_db"] - row[3]) < 1e-4 and abs(o["snr_db"] - row[4]) < 1e-4
        assert o["sync_detected"] == bool(row[5]) and o["sample_count"] == int(row[6])
