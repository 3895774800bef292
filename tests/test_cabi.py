"""CPU: the C-ABI library builds, loads, and exports every symbol include/wcsdr_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os

import pytest


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    import wavecap_sdr_b200._native as N

    return N.lib()


def test_library_exports_every_declared_symbol(lib):
    import wavecap_sdr_b200._native as N

    names = N.exported_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.wc_version()
    assert isinstance(lib.wc_last_error(), bytes)


def test_product_package_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "wavecap-sdr_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f), encoding="utf-8").read()
                for line in src.splitlines():
                    s = line.strip()
                    if (s.startswith("import oracle") or s.startswith("from oracle") or "wavecapsdr" in s and s.startswith(("import ", "from "))
                            and "install" not in f):
                        bad.append((f, s))
    assert not bad, bad


def test_no_gpu_means_loud_failure(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import wavecap_sdr_b200._native as N

    with pytest.raises(N.NativeError):
        N.init(0)
