// Rounding model of the reference's token-similarity tensor and the exact first-argmax
// threshold derived from it.  Host + device (plain integer/float code, no CUDA intrinsics)
// so the same functions are unit-tested on the CPU (tests/test_round_cpu.py builds
// tools/round_selftest.cpp) and used by the kernels.
//
// Reference semantics (src/model.py:387,389 / :505,507):
//   bf16 inputs : S = bf16_rn( float(bf16_rn(acc)) * T ),  argmax_p S  = FIRST index of the max
//   fp32 inputs : S = fp32_rn( acc * T )
// Both roundings are monotone non-decreasing in acc for T > 0, hence
//   max_p round(acc_p) = round(max_p acc_p)         and
//   argmax_first_p round(acc_p) = first p with acc_p >= theta,
// where theta is the smallest fp32 whose rounded value equals round(max).  The epilogue
// therefore needs one fmax pass and one compare pass over raw accumulators instead of two
// conversions per element.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TRIAD_HD __host__ __device__ __forceinline__
#else
#define TRIAD_HD inline
#endif

namespace triad {

TRIAD_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
TRIAD_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x; memcpy(&x, &u, 4); return x;
#endif
}

// fp32 -> nearest bf16 (ties to even), returned as the fp32 value it represents.
TRIAD_HD float bf16_rn(float x) {
    uint32_t u = f2u(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) return x;           // NaN passes through
    u += 0x7fffu + ((u >> 16) & 1u);
    return u2f(u & 0xffff0000u);
}

// IEEE single multiply, never contracted into an FMA.
TRIAD_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}

template <bool kBF16>
TRIAD_HD float round_sim(float acc, float T) {
    if (kBF16) return bf16_rn(mul_rn(bf16_rn(acc), T));
    return mul_rn(acc, T);
}

// Next representable value below x on the bf16 grid (x must itself be on the grid).
TRIAD_HD float bf16_prev(float x) {
    uint32_t u = f2u(x);
    if ((u & 0x7fffffffu) == 0u) return u2f(0x80010000u);    // +-0 -> -min subnormal
    return (u >> 31) ? u2f(u + 0x10000u) : u2f(u - 0x10000u);
}
// Next representable fp32 below x.
TRIAD_HD float f32_prev(float x) {
    uint32_t u = f2u(x);
    if ((u & 0x7fffffffu) == 0u) return u2f(0x80000001u);
    return (u >> 31) ? u2f(u + 1u) : u2f(u - 1u);
}
TRIAD_HD float f32_next(float x) {
    uint32_t u = f2u(x);
    if ((u & 0x7fffffffu) == 0u) return u2f(0x00000001u);
    return (u >> 31) ? u2f(u - 1u) : u2f(u + 1u);
}

// Order-preserving map fp32 -> uint32 (for the bisection fallback).
TRIAD_HD uint32_t f32_key(float x) { uint32_t u = f2u(x); return (u >> 31) ? ~u : (u | 0x80000000u); }
TRIAD_HD float f32_unkey(uint32_t k) { return u2f((k >> 31) ? (k & 0x7fffffffu) : ~k); }

// Smallest fp32 acc with bf16_rn(acc) == b (b on the bf16 grid; +-0 treated as equal).
TRIAD_HD float bf16_interval_lo(float b) {
    float pb = bf16_prev(b);
    float mid = 0.5f * pb + 0.5f * b;                         // exact: 9 significant bits
    bool b_even = ((f2u(b) >> 16) & 1u) == 0u;                // ties-to-even: the midpoint goes to the even one
    return b_even ? mid : f32_next(mid);
}

// theta(M,T): smallest fp32 x with round_sim(x,T) == round_sim(M,T)  (float equality, so
// -0 == +0 like torch.max).  Requires T > 0 and finite M.  *R receives round_sim(M,T).
template <bool kBF16>
TRIAD_HD float argmax_threshold(float M, float T, float* R) {
    const float r = round_sim<kBF16>(M, T);
    *R = r;
    float b = kBF16 ? bf16_rn(M) : M;
    int it = 0;
    for (; it < 8; ++it) {                                    // T in [1,2]: at most two steps
        float pb = kBF16 ? bf16_prev(b) : f32_prev(b);
        float rp = kBF16 ? bf16_rn(mul_rn(pb, T)) : mul_rn(pb, T);
        if (!(rp == r)) break;
        b = pb;
    }
    if (it == 8) {
        // Small T maps many neighbours onto one product: bisect on the ordered-integer image of
        // fp32 for the smallest x with round_sim(x) >= r  (monotone => exact).
        uint32_t lo = f32_key(-3.0e38f), hi = f32_key(b);     // round(lo) < r <= round(hi)
        if (round_sim<kBF16>(-3.0e38f, T) >= r) return -3.0e38f;
        while (hi - lo > 1u) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (round_sim<kBF16>(f32_unkey(mid), T) >= r) hi = mid; else lo = mid;
        }
        return f32_unkey(hi);
    }
    return kBF16 ? bf16_interval_lo(b) : b;
}

}  // namespace triad
