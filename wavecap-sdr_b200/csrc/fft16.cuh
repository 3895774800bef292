// In-register radix-16 FFT building blocks on packed complex values (re = lo, im = hi of a 64-bit
// register pair) shared by the channelizer (FFT-256 = 16 x 16) and the spectrum kernel (65536 = 256 x 256).
#pragma once
#include "common.cuh"

namespace wc {

// Complex values live in one 64-bit register pair (re = lo, im = hi) so that complex add/sub and
// real-scalar multiplies issue as single FADD2/FFMA2 instructions.

// forward 4-point DFT in place: (a,b,c,d) -> (X0,X1,X2,X3); 6 packed + 4 scalar adds
__device__ __forceinline__ void dft4(u64& a, u64& b, u64& c, u64& d) {
    const u64 t0 = add2(a, c), t1 = sub2(a, c), t2 = add2(b, d), t3 = sub2(b, d);
    a = add2(t0, t2);
    c = sub2(t0, t2);
    const float t1x = lo2(t1), t1y = hi2(t1), t3x = lo2(t3), t3y = hi2(t3);
    b = pk2(t1x + t3y, t1y - t3x);  // t1 - j*t3
    d = pk2(t1x - t3y, t1y + t3x);  // t1 + j*t3
}

// v * (c - j*s)
__device__ __forceinline__ u64 twid(u64 v, float c, float s) {
    const float x = lo2(v), y = hi2(v);
    return pk2(fmaf(x, c, y * s), fmaf(y, c, -x * s));
}

// forward 16-point DFT in registers. Input natural order; X[k] ends up in v[4*(k&3) + (k>>2)].
__device__ __forceinline__ void fft16(u64 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // twiddles W16^(n2*k1) on v[4*k1 + n2]
    float x, y;
    v[5] = twid(v[5], C1, S1);                                                              // W^1
    x = lo2(v[6]);  y = hi2(v[6]);  v[6]  = mul2(pk2(x + y, y - x), bc2(R2));               // W^2
    v[7] = twid(v[7], S1, C1);                                                              // W^3
    x = lo2(v[9]);  y = hi2(v[9]);  v[9]  = mul2(pk2(x + y, y - x), bc2(R2));               // W^2
    x = lo2(v[10]); y = hi2(v[10]); v[10] = pk2(y, -x);                                     // W^4
    x = lo2(v[11]); y = hi2(v[11]); v[11] = mul2(pk2(y - x, -(x + y)), bc2(R2));            // W^6
    v[13] = twid(v[13], S1, C1);                                                            // W^3
    x = lo2(v[14]); y = hi2(v[14]); v[14] = mul2(pk2(y - x, -(x + y)), bc2(R2));            // W^6
    v[15] = twid(v[15], -C1, -S1);                                                          // W^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

__device__ __forceinline__ constexpr int rev4(int o) { return ((o & 3) << 2) | (o >> 2); }


}  // namespace wc
