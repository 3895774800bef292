"""GPU: one fixed-seed pass of the randomised GPU-vs-oracle sweep (tools/fuzz_gpu.py) — ragged call sequences with carried
state, random channel configurations and random chunking, every family against its oracle. The tolerances are the families'
own (1e-4 relative RMS on floats, dibits identical, am / ssb / sam pairs bounded by the reference's measured 1-ulp floor)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import parity_note

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("fuzz_gpu", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                        "tools", "fuzz_gpu.py"))
fuzz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fuzz)


@pytest.mark.parametrize("family", sorted(fuzz.FAMILIES))
def test_randomised_sweep(native, family):
    rng = np.random.default_rng([1, sorted(fuzz.FAMILIES).index(family)])
    parity_note("fuzz " + fuzz.FAMILIES[family](rng, 8))
