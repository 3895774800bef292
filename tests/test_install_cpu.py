"""CPU: the rebinding table of wavecap_sdr_b200.install names things that exist on both sides, with the same
call signature (needs /root/reference, so it only runs in the build container)."""
import importlib
import inspect

import pytest

from oracle import refenv


@pytest.mark.reference
@pytest.mark.skipif(not refenv.available(), reason="/root/reference not present")
def test_rebind_table_resolves_and_signatures_match():
    refenv.load()
    from wavecap_sdr_b200.install import REBIND

    for ref_mod, ref_attr, our_mod, our_attr in REBIND:
        r = getattr(importlib.import_module(ref_mod), ref_attr)
        o = getattr(importlib.import_module(our_mod), our_attr)
        if inspect.isclass(r):
            r, o = r.__init__, o.__init__
        rp = [p for p in inspect.signature(r).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
        op = [p for p in inspect.signature(o).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
        assert [p.name for p in rp] == [p.name for p in op][: len(rp)], (ref_mod, ref_attr)
        for a, b in zip(rp, op):
            if a.default is not inspect.Parameter.empty:
                assert a.default == b.default, (ref_mod, ref_attr, a.name)
