// host check (g++ -x c++ -DSEQLOOP_HEADER=\"...seqloop.cuh\" -DITERS=N): clock_run / clock_run8 == the per-sample loop for random clocks, steps, rooms
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#define __device__
#define __forceinline__ inline
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fdividef(float a, float b) { return a / b * (1.0f + 3e-7f); }
static inline int __ffs(unsigned m) { return __builtin_ffs(m); }
using std::min; using std::max;
#include SEQLOOP_HEADER
template <bool GE> int ref(float& cf, float st, int room, bool& fired) {
    int n = 0; fired = false;
    while (n < room) { ++n; cf = __fadd_rn(cf, st); if (GE ? cf >= 1.0f : cf > 1.0f) { fired = true; break; } }
    return n;
}
template <bool GE, int W> long run(unsigned seed) {
    srand(seed); long bad = 0;
    for (int it = 0; it < ITERS; ++it) {
        float cf = (rand() / (float)RAND_MAX) * 1.3f - 0.2f;
        float st; int r = rand() % 10;
        if (r == 0) st = 1e-5f * (rand() % 100); else if (r == 1) st = (rand() / (float)RAND_MAX); else if (r == 2) st = -0.1f;
        else st = 0.1f + 0.02f * ((rand() / (float)RAND_MAX) - 0.5f);
        if (it % 1000 == 0) cf = NAN; if (it % 1001 == 0) st = NAN; if (it % 1003 == 0) cf = -1e9f;
        int room = 1 + rand() % 140;
        float a = cf, b = cf; bool fa = false, fb = false; int na = 0, nb;
        while (na < room) { bool f; na += (W == 8 ? wc::clock_run8<GE>(a, st, room - na, f) : wc::clock_run<GE>(a, st, room - na, f)); fa = f; if (f) break; }
        nb = ref<GE>(b, st, room, fb);
        bool same = na == nb && fa == fb && (a == b || (a != a && b != b));
        if (!same && bad++ < 5) printf("mismatch GE=%d W=%d cf=%g st=%g room=%d: %d %d %d %d %g %g\n", GE, W, cf, st, room, na, nb, fa, fb, a, b);
    }
    return bad;
}
int main() {
    long bad = run<true, 8>(1) + run<false, 8>(2) + run<true, 1>(3) + run<false, 1>(4);
    printf("mismatches: %ld\n", bad); return bad != 0;
}
