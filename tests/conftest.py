import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


# one-line parity facts the tests want visible in every run's output (e.g. the CQPSK knife-edge mismatch count), printed
# in the terminal summary even under -q so a regression from "1 symbol in 218 000" shows up in the driver's log
PARITY_NOTES: list[str] = []


def parity_note(line: str) -> None:
    PARITY_NOTES.append(line)


def pytest_terminal_summary(terminalreporter):
    if PARITY_NOTES:
        terminalreporter.write_line("parity notes:")
        for line in PARITY_NOTES:
            terminalreporter.write_line("  " + line)


def rel_rms(a, b) -> float:
    """relative RMS error ||a-b|| / ||b|| (b = oracle)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = float(np.sqrt(np.mean(np.abs(b.astype(np.complex128)) ** 2)))
    num = float(np.sqrt(np.mean(np.abs(a.astype(np.complex128) - b.astype(np.complex128)) ** 2)))
    return num / den if den > 0 else num


def wrap_rel_rms(a, b, period) -> float:
    """rel-RMS for phase-like outputs: differences are wrapped into (-period/2, period/2]
    (an angle of -pi+eps and +pi-eps are the same discriminator value up to rounding)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = a - b
    d = d - period * np.round(d / period)
    den = float(np.sqrt(np.mean(b ** 2)))
    return float(np.sqrt(np.mean(d ** 2))) / den if den > 0 else float(np.sqrt(np.mean(d ** 2)))


@pytest.fixture(scope="session")
def native():
    """The ctypes library on an initialised B200; GPU tests fail loudly if it is missing."""
    import wavecap_sdr_b200._native as N

    N.init(0)
    return N


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN, name)


# The reference's P25 paths call three numpy routines whose float32 results depend on the host CPU's SIMD
# path: np.arctan2 (SVML), np.convolve on float32 (OpenBLAS sdot) and np.mean (pairwise). The goldens were
# generated on a host with this fingerprint; the LITERAL oracle mode can only be compared bit for bit with them
# on a host that computes the same bits (the portable mode is host independent and is always checked).
GOLDEN_HOST_FINGERPRINT = "6a1bb319817b8d0e"


def host_simd_fingerprint() -> str:
    import hashlib

    rng = np.random.default_rng(123)
    y = rng.standard_normal(4096).astype(np.float32)
    x = rng.standard_normal(4096).astype(np.float32)
    a = np.arctan2(y, x)
    t = rng.standard_normal(63).astype(np.float32)
    c = np.convolve(x, t, mode="same")
    m = np.float32(np.mean(np.abs(x)))
    return hashlib.sha1(a.tobytes() + c.tobytes() + m.tobytes()).hexdigest()[:16]


def require_golden_host():
    if host_simd_fingerprint() != GOLDEN_HOST_FINGERPRINT:
        pytest.skip("host CPU computes float32 arctan2/convolve/mean with different last bits than the golden host")
