"""CPU oracle — TEST INFRASTRUCTURE ONLY.

numpy/scipy restatements of the reference algorithms on the hot path, each
function citing the reference file:line it follows (paths relative to /root/reference/backend/).
Nothing in the product package (`wavecap-sdr_b200/`) imports this; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do, and only as
the checker or the timed CPU baseline.

Pinning: the reference repository holds no golden vectors for this path. Every
restatement here is pinned instead against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py`
imports /root/reference in the build container, runs the reference functions on seeded inputs and
commits the results under tests/golden/; tests/test_oracle_*.py check the restatements against
those fixtures (and, when /root/reference is present, against the live reference).

Third-party arithmetic the reference delegates to (not vendored in /root/reference; caret-pinned in
backend/pyproject.toml:13-15 as numpy ^1.26, scipy ^1.11, numba ^0.60; executed here with numpy
2.3.5 / scipy 1.18.1): scipy.signal.{firwin,lfilter,butter,iirnotch,resample_poly,remez,lfilter_zi},
scipy.fft / numpy.fft, numpy.convolve. The oracle calls the same library functions where the
reference does; the CUDA path restates their published algorithms.
"""
