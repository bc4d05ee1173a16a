// CPU self-test of triad_round.h: for many (M,T) check that theta is the SMALLEST fp32 whose
// rounded similarity equals the rounded maximum.  Built and run by tests/test_round_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../triad_b200/csrc/triad_round.h"
using namespace triad;

template <bool BF> static long check(float M, float T) {
    float R; float th = argmax_threshold<BF>(M, T, &R);
    long bad = 0;
    if (!(round_sim<BF>(th, T) == R)) bad++;
    if (!(th <= M)) bad++;
    float below = f32_prev(th);
    if (round_sim<BF>(below, T) == R) bad++;           // theta must be minimal
    if (bad) printf("BAD bf=%d M=%.9g (0x%08x) T=%.9g theta=%.9g R=%.9g r(th)=%.9g r(below)=%.9g\n", (int)BF, M, f2u(M), T, th, R,
                    round_sim<BF>(th, T), round_sim<BF>(below, T));
    return bad;
}

int main() {
    long bad = 0, n = 0;
    const float Ts[] = {1.5f, 1.2f, 2.0f, 1.0f, 0.07f, 0.7f, 3.3333f, 1.9999999f, 1e-3f, 14.285714f};
    std::mt19937 rng(123);
    std::uniform_int_distribution<uint32_t> bits(0, 0xffffffffu);
    for (float T : Ts) {
        // every bf16 grid point and its neighbourhood in a sane magnitude range, both signs
        for (uint32_t hi = 0x0000; hi < 0x10000; ++hi) {
            uint32_t e = (hi >> 7) & 0xff;
            if (e == 0xff) continue;                     // inf / nan
            if (e > 0xf0) continue;                      // keep products finite
            for (uint32_t lo : {0x0000u, 0x0001u, 0x7fffu, 0x8000u, 0x8001u, 0xffffu}) {
                float M = u2f((hi << 16) | lo);
                bad += check<true>(M, T); bad += check<false>(M, T); n += 2;
            }
        }
        for (int i = 0; i < 2000000; ++i) {
            uint32_t u = bits(rng);
            uint32_t e = (u >> 23) & 0xff;
            if (e >= 0xf0) continue;
            float M = u2f(u);
            bad += check<true>(M, T); bad += check<false>(M, T); n += 2;
        }
    }
    // bf16_rn against a straightforward reference on random values
    for (int i = 0; i < 1000000; ++i) {
        uint32_t u = bits(rng); if (((u >> 23) & 0xff) >= 0xfe) continue;
        float x = u2f(u); float r = bf16_rn(x);
        uint32_t ru = f2u(r); if (ru & 0xffff) { bad++; printf("bf16_rn not on grid\n"); }
        float lo = u2f(u & 0xffff0000u), hi2 = u2f((u & 0xffff0000u) + 0x10000u);
        double dl = std::fabs((double)x - lo), dh = std::fabs((double)hi2 - x);
        float want = dl < dh ? lo : (dh < dl ? hi2 : ((((u >> 16) & 1) == 0) ? lo : hi2));
        if (!(want == r) && !(std::isinf(want) && std::isinf(r))) { bad++; if (bad < 20) printf("bf16_rn(%a)=%a want %a\n", x, r, want); }
        n++;
    }
    printf("checked=%ld bad=%ld\n", n, bad);
    return bad ? 1 : 0;
}
