// Host check of csrc/fft16_body.inc: the same text compiled with scalar emulations of the packed primitives, compared with a
// float64 DFT. Build and run:  g++ -O1 -std=c++17 -o /tmp/fft16_check tools/fft16_host_check.cpp && /tmp/fft16_check
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#define __device__
#define __forceinline__ inline
namespace wc {
typedef uint64_t u64;
static inline u64 pk2(float lo, float hi) { uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (u64)a | ((u64)b << 32); }
static inline float lo2(u64 v) { uint32_t a = (uint32_t)v; float f; memcpy(&f, &a, 4); return f; }
static inline float hi2(u64 v) { uint32_t a = (uint32_t)(v >> 32); float f; memcpy(&f, &a, 4); return f; }
static inline u64 bc2(float v) { return pk2(v, v); }
static inline u64 add2(u64 a, u64 b) { return pk2(lo2(a) + lo2(b), hi2(a) + hi2(b)); }
static inline u64 sub2(u64 a, u64 b) { return pk2(lo2(a) - lo2(b), hi2(a) - hi2(b)); }
static inline u64 mul2(u64 a, u64 b) { return pk2(lo2(a) * lo2(b), hi2(a) * hi2(b)); }
static inline u64 fma2(u64 a, u64 b, u64 c) { return pk2(fmaf(lo2(a), lo2(b), lo2(c)), fmaf(hi2(a), hi2(b), hi2(c))); }
#include "../wavecap-sdr_b200/csrc/fft16_body.inc"
}  // namespace wc

int main() {
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    double worst = 0.0, worst_tw = 0.0;
    for (int trial = 0; trial < 200; ++trial) {
        wc::u64 v[16];
        std::complex<double> x[16];
        for (int i = 0; i < 16; ++i) {
            const float re = nd(rng), im = nd(rng);
            v[i] = wc::pk2(re, im);
            x[i] = {re, im};
        }
        wc::fft16(v);
        double err = 0, ref = 0;
        for (int k = 0; k < 16; ++k) {
            std::complex<double> X = 0;
            for (int n = 0; n < 16; ++n) X += x[n] * std::polar(1.0, -2.0 * M_PI * k * n / 16.0);
            const wc::u64 g = v[4 * (k & 3) + (k >> 2)];
            err += std::norm(std::complex<double>(wc::lo2(g), wc::hi2(g)) - X);
            ref += std::norm(X);
        }
        worst = std::fmax(worst, std::sqrt(err / ref));
        const float c = std::cos(0.1f * trial), s = std::sin(0.1f * trial);
        const wc::u64 t = wc::twid(v[3], c, s);
        const std::complex<double> e = std::complex<double>(wc::lo2(v[3]), wc::hi2(v[3])) * std::complex<double>(c, -s);
        worst_tw = std::fmax(worst_tw, std::abs(std::complex<double>(wc::lo2(t), wc::hi2(t)) - e) / std::abs(e));
    }
    printf("fft16 worst rel-RMS vs float64 DFT: %.3e   twid worst rel error: %.3e\n", worst, worst_tw);
    return (worst < 5e-7 && worst_tw < 5e-7) ? 0 : 1;
}
