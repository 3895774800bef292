// Shared helpers for the wcsdr_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace wc {

// ---- error plumbing: every C-ABI entry returns 0 or a negative code and leaves a message ----
void set_error(const char* fmt, ...);
const char* get_error();

#define WC_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::wc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                            cudaGetErrorString(_e));                                    \
            return -2;                                                                  \
        }                                                                               \
    } while (0)

#define WC_REQUIRE(cond, ...)                                                           \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            ::wc::set_error(__VA_ARGS__);                                               \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

int sm_count();

// Opt-in to > 48 KB of dynamic shared memory. The attribute belongs to the (function, device) pair, so it is set once
// for every device this process launches on (one bit per device ordinal in `done`), not once per process.
template <typename F>
inline cudaError_t smem_optin(F* func, int bytes, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
// tuning / A-B switches for the dev tools (WC_* environment variables): read only in dev builds (-DWC_DEV, i.e.
// WC_DEV=1 python csrc/build.py). The product library never looks at the environment: env_int is the default, inlined.
#ifdef WC_DEV
int env_int(const char* name, int dflt);
#else
inline int env_int(const char*, int dflt) { return dflt; }
#endif

// ---- device helpers ----
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) — global -> shared.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// plain arrival (release, CTA scope): split-barrier use — every thread arrives, waits later on the phase parity
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WC_DONE_%=;\n"
        "bra WC_WAIT_%=;\n"
        "WC_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 4-byte asynchronous global -> shared copies (LDGSTS): staging of per-channel rows for the sequential per-channel kernels
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// capture.freq_shift (capture.py:166-193): exp(j * fl32(k32 * fl32(n))) with k32 = float32(-2*pi*round(off)/fs)
// nf = float32(n), the value numpy's float32 arange holds.
__device__ __forceinline__ void nco_f32f(float k32, float nf, float& c, float& s) {
    // theta = fl32(k32 * fl32(n)) exactly as numpy computes it; then an accurate cos/sin of that float32 angle. The
    // turn count theta/(2 pi) is formed in float-float arithmetic — product by the leading part of 1/(2 pi), its exact
    // rounding error (FMA), the trailing part — so the fractional turn is good to ~3e-8 (2e-7 rad, the float32 rounding
    // of the reduced angle itself) for |theta| far beyond 2^24, without touching the FP64 pipe or its conversions:
    // the float64 reduction this replaces cost the analog front end most of its XU-pipe time (ncu, round 1).
    const float th = __fmul_rn(k32, nf);
    const float C_HI = 0.15915493667125702f, C_LO = 6.4206382432985265e-09f;   // 1/(2 pi) = C_HI + C_LO
    const float hi = __fmul_rn(th, C_HI);
    const float e = __fmaf_rn(th, C_HI, -hi);
    const float lo = __fmaf_rn(th, C_LO, e);
    const float fr = __fadd_rn(__fsub_rn(hi, rintf(hi)), lo);   // hi - rint(hi) is exact
    sincospif(2.0f * fr, &s, &c);
}
__device__ __forceinline__ void nco_f32(float k32, int n, float& c, float& s) { nco_f32f(k32, (float)n, c, s); }

// ---- small complex helpers on float2 ----
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// ---- packed FP32 pairs (sm_100 FFMA2/FADD2/FMUL2: one issue slot, two lanes of the FMA pipe) ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(u64 v) {
    float a;
    asm("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(a) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi2(u64 v) {
    float b;
    asm("{ .reg .b32 t; mov.b64 {t, %0}, %1; }" : "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ u64 bc2(float v) { return pk2(v, v); }  // folds into a .F32 broadcast operand
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// atan2 with a degree-13 odd minimax polynomial (relative error <= 6.5e-7, fitted offline with
// tools/fit_atan.py). No divergence, one MUFU.RCP. atan2(0,0)=0 like numpy.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = __fdividef(mn, fmaxf(mx, 1e-37f));
    const float s = t * t;
    float p = 0.008007131516933441f;
    p = fmaf(p, s, -0.037443727254867554f);
    p = fmaf(p, s, 0.08435501158237457f);
    p = fmaf(p, s, -0.13512229919433594f);
    p = fmaf(p, s, 0.198873370885849f);
    p = fmaf(p, s, -0.3332701623439789f);
    p = fmaf(p, s, 0.9999994039535522f);
    float a = p * t;
    a = (ay > ax) ? (1.57079632679489662f - a) : a;
    a = (x < 0.0f) ? (3.14159265358979324f - a) : a;
    return copysignf(a, y);
}

// degree-17 variant for the analog chain's discriminator (dsp/fm.py:65-97 feeds audio filters whose parity budget is
// tighter than the channelizer's): max relative error 1.4e-7 in float32 Horner arithmetic (tools/fit_atan.py 9) plus the
// 2-ulp quotient — the accuracy class of libm's atan2f at about half its instructions (no IEEE-division slow path).
__device__ __forceinline__ float fast_atan2f_hi(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = __fdividef(mn, fmaxf(mx, 1e-37f));
    const float s = t * t;
    float p = 0.0028179744258522987f;
    p = fmaf(p, s, -0.01606472209095955f);
    p = fmaf(p, s, 0.04292432218790054f);
    p = fmaf(p, s, -0.07548770308494568f);
    p = fmaf(p, s, 0.10676739364862442f);
    p = fmaf(p, s, -0.1421811282634735f);
    p = fmaf(p, s, 0.19995509088039398f);
    p = fmaf(p, s, -0.3333330750465393f);
    p = fmaf(p, s, 1.0f);
    float a = p * t;
    a = (ay > ax) ? (1.57079632679489662f - a) : a;
    a = (x < 0.0f) ? (3.14159265358979324f - a) : a;
    return copysignf(a, y);
}

__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// two atan2's at once: the polynomial and the products run packed, the quadrant fix-ups scalar.
// y/x = (lo, hi) pairs. Same polynomial as fast_atan2f.
__device__ __forceinline__ u64 fast_atan2f_x2(u64 y, u64 x) {
    const float x0 = lo2(x), x1 = hi2(x), y0 = lo2(y), y1 = hi2(y);
    const float ax0 = fabsf(x0), ay0 = fabsf(y0), ax1 = fabsf(x1), ay1 = fabsf(y1);
    const float mx0 = fmaxf(fmaxf(ax0, ay0), 1e-37f), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(fmaxf(ax1, ay1), 1e-37f), mn1 = fminf(ax1, ay1);
    const u64 t = mul2(pk2(mn0, mn1), pk2(rcp_approx(mx0), rcp_approx(mx1)));
    const u64 s = mul2(t, t);
    u64 p = bc2(0.008007131516933441f);
    p = fma2(p, s, bc2(-0.037443727254867554f));
    p = fma2(p, s, bc2(0.08435501158237457f));
    p = fma2(p, s, bc2(-0.13512229919433594f));
    p = fma2(p, s, bc2(0.198873370885849f));
    p = fma2(p, s, bc2(-0.3332701623439789f));
    p = fma2(p, s, bc2(0.9999994039535522f));
    const u64 a = mul2(p, t);
    float a0 = lo2(a), a1 = hi2(a);
    a0 = (ay0 > ax0) ? (1.57079632679489662f - a0) : a0;
    a1 = (ay1 > ax1) ? (1.57079632679489662f - a1) : a1;
    a0 = (x0 < 0.0f) ? (3.14159265358979324f - a0) : a0;
    a1 = (x1 < 0.0f) ? (3.14159265358979324f - a1) : a1;
    return pk2(copysignf(a0, y0), copysignf(a1, y1));
}

// Discriminator flavour of the above: returns scale * atan2(y, x) for two (y, x) pairs with the output scale folded
// into the polynomial coefficients and the quadrant constants (k[0..4] = scale * c_i, hp = scale*pi/2, pi = scale*pi).
// Degree-9 odd minimax polynomial in t = min/max (tools/fit_atan.py: max relative error 3.0e-5, i.e. <= 3e-5
// relative on every output sample, inside the 1e-4 relative-RMS budget of the FM path).
struct AtanScaled {
    float k0, k1, k2, k3, k4, hp, pi;
    float h[7];   // degree-13 set (fast_atan2f's coefficients x scale) for scaled_atan2f_hi_x2
};
__device__ __forceinline__ u64 scaled_atan2f_x2(u64 y, u64 x, const AtanScaled& c) {
    const float x0 = lo2(x), x1 = hi2(x), y0 = lo2(y), y1 = hi2(y);
    const float ax0 = fabsf(x0), ay0 = fabsf(y0), ax1 = fabsf(x1), ay1 = fabsf(y1);
    const float mx0 = fmaxf(fmaxf(ax0, ay0), 1e-37f), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(fmaxf(ax1, ay1), 1e-37f), mn1 = fminf(ax1, ay1);
    const u64 t = mul2(pk2(mn0, mn1), pk2(rcp_approx(mx0), rcp_approx(mx1)));
    const u64 s = mul2(t, t);
    u64 p = bc2(c.k4);
    p = fma2(p, s, bc2(c.k3));
    p = fma2(p, s, bc2(c.k2));
    p = fma2(p, s, bc2(c.k1));
    p = fma2(p, s, bc2(c.k0));
    const u64 a = mul2(p, t);
    float a0 = lo2(a), a1 = hi2(a);
    a0 = (ay0 > ax0) ? (c.hp - a0) : a0;
    a1 = (ay1 > ax1) ? (c.hp - a1) : a1;
    a0 = (x0 < 0.0f) ? (c.pi - a0) : a0;
    a1 = (x1 < 0.0f) ? (c.pi - a1) : a1;
    return pk2(copysignf(a0, y0), copysignf(a1, y1));
}

// Degree-13 flavour (max relative error 6.5e-7, the polynomial of fast_atan2f) for consumers that filter the discriminator
// output: after rms_normalize and the /20 low-pass of the audio mode the degree-9 polynomial's systematic error is
// 1.7e-4 of the audio (measured), above the 1e-4 budget; this one leaves 2.8e-6.
__device__ __forceinline__ u64 scaled_atan2f_hi_x2(u64 y, u64 x, const AtanScaled& c) {
    const float x0 = lo2(x), x1 = hi2(x), y0 = lo2(y), y1 = hi2(y);
    const float ax0 = fabsf(x0), ay0 = fabsf(y0), ax1 = fabsf(x1), ay1 = fabsf(y1);
    const float mx0 = fmaxf(fmaxf(ax0, ay0), 1e-37f), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(fmaxf(ax1, ay1), 1e-37f), mn1 = fminf(ax1, ay1);
    const u64 t = mul2(pk2(mn0, mn1), pk2(rcp_approx(mx0), rcp_approx(mx1)));
    const u64 s = mul2(t, t);
    u64 p = bc2(c.h[6]);
    p = fma2(p, s, bc2(c.h[5]));
    p = fma2(p, s, bc2(c.h[4]));
    p = fma2(p, s, bc2(c.h[3]));
    p = fma2(p, s, bc2(c.h[2]));
    p = fma2(p, s, bc2(c.h[1]));
    p = fma2(p, s, bc2(c.h[0]));
    const u64 a = mul2(p, t);
    float a0 = lo2(a), a1 = hi2(a);
    a0 = (ay0 > ax0) ? (c.hp - a0) : a0;
    a1 = (ay1 > ax1) ? (c.hp - a1) : a1;
    a0 = (x0 < 0.0f) ? (c.pi - a0) : a0;
    a1 = (x1 < 0.0f) ? (c.pi - a1) : a1;
    return pk2(copysignf(a0, y0), copysignf(a1, y1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace wc
