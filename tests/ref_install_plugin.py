"""pytest plugin (test infrastructure): import the packed reference and rebind its DSP entry points to the CUDA
implementation BEFORE the reference's own test modules are imported — they bind names with `from wavecapsdr.dsp.fm import
wbfm_demod` at import time. Used by tests/test_reference_unit_tests_gpu.py in a subprocess:
    python -m pytest -p ref_install_plugin <reference test files>
WC_REF_NO_INSTALL=1 leaves the reference untouched (the baseline run of the same files)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    from oracle import build_ref

    build_ref.load()   # trunking before capture: the reference's import cycle
    config.addinivalue_line("markers", "perf: reference marker")
    config.addinivalue_line("markers", "hardware: reference marker")
    if os.environ.get("WC_REF_NO_INSTALL") != "1":
        import wavecap_sdr_b200.install as b200

        names = b200.install(int(os.environ.get("LOCAL_RANK", "0")))
        config._wc_installed = names


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    names = getattr(config, "_wc_installed", None)
    terminalreporter.write_line(f"wavecap_sdr_b200.install(): {len(names)} reference names rebound" if names
                                else "reference untouched (baseline)")
