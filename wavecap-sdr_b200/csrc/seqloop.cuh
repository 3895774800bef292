// Helpers for the sequential per-channel loops (symbol clocks of the P25 demodulators). One warp serves up to 32 channels
// and nothing else runs on its SM, so every taken branch and every dependent issue is exposed latency: the helpers below
// are straight-line code.
#pragma once

namespace wc {

// Advance a float32 symbol clock by rounded additions of `st` — exactly the reference's per-sample `clock += symbol_time`
// sequence — until it crosses 1 or `room` (>= 1) samples are consumed, at most eight samples per call. Returns the number
// of samples consumed; `fired` says the last of them took the clock across (GE: clock >= 1, else clock > 1).
// The eight sums are a dependent chain either way; what this removes is the compare-and-branch per sample.
template <bool GE>
__device__ __forceinline__ int clock_run8(float& cf, float st, int room, bool& fired) {
    const float c0 = __fadd_rn(cf, st), c1 = __fadd_rn(c0, st), c2 = __fadd_rn(c1, st), c3 = __fadd_rn(c2, st);
    const float c4 = __fadd_rn(c3, st), c5 = __fadd_rn(c4, st), c6 = __fadd_rn(c5, st), c7 = __fadd_rn(c6, st);
#define WC_X(c) (GE ? ((c) >= 1.0f) : ((c) > 1.0f))
    const unsigned m = (WC_X(c0) ? 1u : 0u) | (WC_X(c1) ? 2u : 0u) | (WC_X(c2) ? 4u : 0u) | (WC_X(c3) ? 8u : 0u) |
                       (WC_X(c4) ? 16u : 0u) | (WC_X(c5) ? 32u : 0u) | (WC_X(c6) ? 64u : 0u) | (WC_X(c7) ? 128u : 0u);
#undef WC_X
    const int j = m ? __ffs(m) : 8;   // samples up to and including the first crossing
    const int take = min(j, room);
    const float lo = take <= 2 ? (take <= 1 ? c0 : c1) : (take == 3 ? c2 : c3);
    const float hi = take <= 6 ? (take == 5 ? c4 : c5) : (take == 7 ? c6 : c7);
    cf = take <= 4 ? lo : hi;
    fired = (m != 0u) && take == j;
    return take;
}

// The same walk with the crossing test only where a crossing is possible: (1 - clock) / st additions are needed to reach
// 1, so all but the last two of them (rounding moves the sum by < 1e-6, st >= 1e-4 here) run unchecked — predicated, lanes
// differ — and one checked block of four follows. Returns the samples consumed (<= 16); the caller loops while there is
// room and `fired` is false. 85 instructions per symbol period of ten samples instead of 110 — and the same time on B200
// (5.52 ms for the 64-channel CQPSK bank either way): the estimate's MUFU.RCP -> F2I chain costs what the tests saved.
template <bool GE>
__device__ __forceinline__ int clock_run(float& cf, float st, int room, bool& fired) {
    const float est = __fdividef(1.0f - cf, st);          // NaN -> 0, huge -> INT_MAX below
    int ns = (st >= 1.0e-4f) ? (int)est - 2 : 0;
    ns = max(0, min(min(ns, 12), room));
#pragma unroll
    for (int q = 0; q < 12; ++q)
        if (q < ns) cf = __fadd_rn(cf, st);
    room -= ns;
    fired = false;
    if (room <= 0) return ns;
    const float c0 = __fadd_rn(cf, st), c1 = __fadd_rn(c0, st), c2 = __fadd_rn(c1, st), c3 = __fadd_rn(c2, st);
#define WC_X(c) (GE ? ((c) >= 1.0f) : ((c) > 1.0f))
    const unsigned m = (WC_X(c0) ? 1u : 0u) | (WC_X(c1) ? 2u : 0u) | (WC_X(c2) ? 4u : 0u) | (WC_X(c3) ? 8u : 0u);
#undef WC_X
    const int j = m ? __ffs(m) : 4;
    const int take = min(j, room);
    cf = take <= 2 ? (take <= 1 ? c0 : c1) : (take == 3 ? c2 : c3);
    fired = (m != 0u) && take == j;
    return ns + take;
}

}  // namespace wc
