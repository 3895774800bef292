"""dev: SAM stateless branch vs the golden, with intermediate checks and a throughput probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import analog as oa
from oracle.make_golden import sam_stateless_input, sam_stateless_cases
import wavecap_sdr_b200.capture as cap
from wavecap_sdr_b200.dsp import sam as gs, am as AM, _stages as S

# throughput: 64 sequences x 120000 samples
x = (torch.randn(64, 120000, dtype=torch.complex64, device="cuda") * 0.01 + 0.3)
for ex in (True, False):
    for n_seq in (1, 64):
        gs.pll_rows(x[:n_seq], 1e-3, 1e-6, 0, exact=ex)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gs.pll_rows(x[:n_seq], 1e-3, 1e-6, 0, exact=ex)
        torch.cuda.synchronize(); print("exact" if ex else "fast", n_seq, "seq x 120000:", (time.perf_counter() - t0) * 1e3, "ms")
