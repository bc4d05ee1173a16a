// Positive-pair regularisers of the reference's loss and a device-scalar scale.
//
// Both terms read only the B diagonal blocks token_sims[i,i] (1/B of the similarity tensor):
//   mode 0  temporal smoothness, audio-visual  (src/model.py:394-408):
//           value = mean over (i, a < Nq-1, p) of (S[i,a+1,p] - S[i,a,p])^2
//   mode 1  patch-usage sparsity, text-visual  (src/model.py:528-541):
//           probs = softmax_p(S[i]);  frac[i,p] = sum_t probs[i,t,p] / Nq   (padded tokens take part, as in the
//           reference);  value = mean over (i, p) of relu(frac - threshold)^2
// with S = round(T * raw), raw[i,a,p] = <q[i,a], v[i,p]> the output of one small batched library GEMM in the input
// dtype (bf16 inputs: raw and S are bf16-rounded exactly like the reference's matmul and `* temperature` under
// autocast, model.py:387; everything after that is evaluated in fp32 here, where the reference keeps bf16).
//
// The reference builds these with ~40 ATen elementwise / reduction kernels and their autograd graph (0.96 ms at
// B=256, 250 x 256).  Here ONE pass produces the value, G = d value / d raw (same dtype and shape as raw, the operand
// of the two small backward GEMMs dq_i = G_i v_i, dv_i = G_i^T q_i) and d value / dT = sum G_S * raw.
// Deterministic: fp64 per-block partials, reduced in block order by the finishing launch.
#include "common.cuh"

namespace triad {
namespace pospair {

constexpr int kThreads = 256;

template <typename T> struct Elt;
template <> struct Elt<__nv_bfloat16> {
    static __device__ __forceinline__ float load(const void* p, size_t k) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[k]); }
    static __device__ __forceinline__ void store(void* p, size_t k, float x) { reinterpret_cast<__nv_bfloat16*>(p)[k] = __float2bfloat16_rn(x); }
    static __device__ __forceinline__ float scaled(float raw, float Tv) { return __bfloat162float(__float2bfloat16_rn(raw * Tv)); }
};
template <> struct Elt<float> {
    static __device__ __forceinline__ float load(const void* p, size_t k) { return reinterpret_cast<const float*>(p)[k]; }
    static __device__ __forceinline__ void store(void* p, size_t k, float x) { reinterpret_cast<float*>(p)[k] = x; }
    static __device__ __forceinline__ float scaled(float raw, float Tv) { return raw * Tv; }
};

__device__ __forceinline__ void block_partials(double a, double b, double* __restrict__ partials) {
    __shared__ double ra[kThreads / 32], rb[kThreads / 32];
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) { x += ra[w]; y += rb[w]; }
        partials[2 * blockIdx.x] = x;
        partials[2 * blockIdx.x + 1] = y;
    }
}

// ---- mode 0: temporal smoothness --------------------------------------------------------------------------
// blocks walk the B*Nq token rows, threads the patches of a row (no index divisions); the three rows a-1, a, a+1
// are neighbours in memory (L1/L2 hits)
template <typename T>
__global__ void __launch_bounds__(kThreads)
smooth_kernel(const void* __restrict__ raw, const float* __restrict__ Tptr, int B, int Nq, int Nv,
              void* __restrict__ G, double* __restrict__ partials) {
    const float Tv = *Tptr;
    const double cnt = (double)B * (double)(Nq - 1) * (double)Nv;
    const float c = (Nq > 1) ? (float)(2.0 / cnt) : 0.f;
    double s2 = 0.0, sT = 0.0;
    const int rows = B * Nq;
    int a = (int)(blockIdx.x % (unsigned)Nq);
    const int a_step = (int)(gridDim.x % (unsigned)Nq);
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const size_t k0 = (size_t)r * Nv;
        float r2 = 0.f, rT = 0.f;                                       // this row's share in fp32 (an fp64 add is 1/64 rate)
        for (int p = threadIdx.x; p < Nv; p += kThreads) {
            const size_t k = k0 + p;
            const float r0 = Elt<T>::load(raw, k);
            const float s0 = Elt<T>::scaled(r0, Tv);
            float dprev = 0.f, dnext = 0.f;                             // d[a-1] = S[a]-S[a-1], d[a] = S[a+1]-S[a]
            if (a > 0) dprev = s0 - Elt<T>::scaled(Elt<T>::load(raw, k - Nv), Tv);
            if (a + 1 < Nq) dnext = Elt<T>::scaled(Elt<T>::load(raw, k + Nv), Tv) - s0;
            const float gs = c * (dprev - dnext);                       // d value / dS[i,a,p]
            r2 = fmaf(dnext, dnext, r2);                                // every difference is counted once, at its lower row
            rT = fmaf(gs, r0, rT);                                      // dS/dT = raw
            Elt<T>::store(G, k, gs * Tv);
        }
        s2 += (double)r2; sT += (double)rT;
        a += a_step;
        if (a >= Nq) a -= Nq;
    }
    block_partials(s2, sT, partials);
}

// ---- mode 1: patch-usage sparsity --------------------------------------------------------------------------
// one CTA per positive pair i; dynamic shared memory: m[Nq] | rz[Nq] | gp[Nv] | red[kThreads]
template <typename T>
__global__ void __launch_bounds__(kThreads)
sparsity_kernel(const void* __restrict__ raw, const float* __restrict__ Tptr, float threshold, int B, int Nq, int Nv,
                void* __restrict__ G, double* __restrict__ partials) {
    extern __shared__ float sm[];
    float* m = sm;                 // row maxima
    float* rz = sm + Nq;           // 1 / sum exp
    float* gp = rz + Nq;           // d value / d probs[t,p] (the same for every t)
    const float Tv = *Tptr;
    const int i = blockIdx.x;
    const size_t base = (size_t)i * Nq * Nv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (A) softmax statistics of every token row: one warp per row
    for (int t = warp; t < Nq; t += kThreads / 32) {
        float mx = -INFINITY;
        for (int p = lane; p < Nv; p += 32) mx = fmaxf(mx, Elt<T>::scaled(Elt<T>::load(raw, base + (size_t)t * Nv + p), Tv));
        mx = warp_max(mx);
        float z = 0.f;
        for (int p = lane; p < Nv; p += 32) z += __expf(Elt<T>::scaled(Elt<T>::load(raw, base + (size_t)t * Nv + p), Tv) - mx);
        z = warp_sum(z);
        if (lane == 0) { m[t] = mx; rz[t] = 1.f / z; }
    }
    __syncthreads();
    // (B) usage fraction of every patch (fixed order over the rows), its excess and the gradient seed
    const float inv_nq = 1.f / (float)Nq;
    const float cg = (float)(2.0 / ((double)B * (double)Nv)) * inv_nq;
    double v2 = 0.0;
    for (int p = threadIdx.x; p < Nv; p += kThreads) {
        float f = 0.f;
        for (int t = 0; t < Nq; ++t) f += __expf(Elt<T>::scaled(Elt<T>::load(raw, base + (size_t)t * Nv + p), Tv) - m[t]) * rz[t];
        const float e = fmaxf(f * inv_nq - threshold, 0.f);
        v2 += (double)(e * e);
        gp[p] = cg * e;
    }
    __syncthreads();
    // (C) softmax backward per row: dS[t,p] = probs[t,p] * (gp[p] - sum_p' probs[t,p'] gp[p'])
    double sT = 0.0;
    for (int t = warp; t < Nq; t += kThreads / 32) {
        float dot = 0.f, rT = 0.f;
        for (int p = lane; p < Nv; p += 32)
            dot += __expf(Elt<T>::scaled(Elt<T>::load(raw, base + (size_t)t * Nv + p), Tv) - m[t]) * rz[t] * gp[p];
        dot = warp_sum(dot);
        for (int p = lane; p < Nv; p += 32) {
            const float r0 = Elt<T>::load(raw, base + (size_t)t * Nv + p);
            const float pr = __expf(Elt<T>::scaled(r0, Tv) - m[t]) * rz[t];
            const float gs = pr * (gp[p] - dot);
            rT = fmaf(gs, r0, rT);
            Elt<T>::store(G, base + (size_t)t * Nv + p, gs * Tv);
        }
        sT += (double)rT;
    }
    block_partials(v2, sT, partials);
}

// sums[0] = scale0 * sum partials[.][0], sums[1] = sum partials[.][1]   (block order: deterministic)
__global__ void finish_kernel(const double* __restrict__ partials, int nblocks, double scale0, double* __restrict__ sums) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += 32) { a += partials[2 * k]; b += partials[2 * k + 1]; }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if (threadIdx.x == 0) { sums[0] = a * scale0; sums[1] = b; }
}

// y = x * (*scale): 16-byte vectors, scalar tail
template <typename T>
__global__ void __launch_bounds__(kThreads)
scale_kernel(const void* __restrict__ x, void* __restrict__ y, size_t n, const float* __restrict__ scale) {
    constexpr int E = 16 / sizeof(T);
    const float s = *scale;
    const size_t nvec = n / E;
    for (size_t k = (size_t)blockIdx.x * kThreads + threadIdx.x; k < nvec; k += (size_t)gridDim.x * kThreads) {
        uint4 u = reinterpret_cast<const uint4*>(x)[k];
        if constexpr (sizeof(T) == 2) {
            uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float lo = __uint_as_float(w[c] << 16) * s, hi = __uint_as_float(w[c] & 0xffff0000u) * s;
                __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
                w[c] = *reinterpret_cast<uint32_t*>(&h);
            }
        } else {
            float* f = reinterpret_cast<float*>(&u);
#pragma unroll
            for (int c = 0; c < 4; ++c) f[c] *= s;
        }
        reinterpret_cast<uint4*>(y)[k] = u;
    }
    if (blockIdx.x == 0)
        for (size_t k = nvec * E + threadIdx.x; k < n; k += kThreads) Elt<T>::store(y, k, Elt<T>::load(x, k) * s);
}

constexpr int kMaxBlocks = 148 * 8;

}  // namespace pospair
}  // namespace triad

using namespace triad;

extern "C" size_t triad_pospair_workspace_bytes(int B) {
    const int blocks = B > pospair::kMaxBlocks ? B : pospair::kMaxBlocks;
    return (size_t)blocks * 2 * sizeof(double);
}

extern "C" int triad_pospair_terms(const void* raw, int dtype, const float* temperature, int mode, float threshold,
                                   int B, int Nq, int Nv, void* G, double* sums, void* ws, size_t ws_bytes, void* stream) {
    if (!raw || !temperature || !G || !sums || !ws) return fail_msg(TRIAD_ERR_BAD_ARG, "pospair_terms: null pointer");
    if (dtype != TRIAD_DTYPE_F32 && dtype != TRIAD_DTYPE_BF16) return fail_msg(TRIAD_ERR_BAD_ARG, "pospair_terms: dtype");
    if (mode != 0 && mode != 1) return fail_msg(TRIAD_ERR_BAD_ARG, "pospair_terms: mode");
    if (B <= 0 || Nq <= 0 || Nv <= 0) return fail_msg(TRIAD_ERR_BAD_SHAPE, "pospair_terms: bad shape");
    if (ws_bytes < triad_pospair_workspace_bytes(B)) return fail_msg(TRIAD_ERR_WORKSPACE, "pospair_terms: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)ws;
    int blocks;
    double scale0;
    if (mode == 0) {
        if ((long long)B * Nq > 0x7fffffffLL) return fail_msg(TRIAD_ERR_BAD_SHAPE, "pospair_terms: B*Nq overflows int32");
        blocks = B * Nq < pospair::kMaxBlocks ? B * Nq : pospair::kMaxBlocks;
        // mean over B*(Nq-1)*Nv differences; a single token row has none: 0/0 = NaN, like torch.mean of an empty tensor
        scale0 = 1.0 / ((double)B * (double)(Nq - 1) * (double)Nv);
        if (dtype == TRIAD_DTYPE_BF16)
            pospair::smooth_kernel<__nv_bfloat16><<<blocks, pospair::kThreads, 0, st>>>(raw, temperature, B, Nq, Nv, G, partials);
        else
            pospair::smooth_kernel<float><<<blocks, pospair::kThreads, 0, st>>>(raw, temperature, B, Nq, Nv, G, partials);
        TRIAD_LAUNCH_CHECK("pospair smooth_kernel");
    } else {
        blocks = B;
        scale0 = 1.0 / ((double)B * (double)Nv);
        const size_t smem = ((size_t)2 * Nq + Nv) * sizeof(float);
        if (smem > 48 * 1024) return fail_msg(TRIAD_ERR_UNSUPPORTED, "pospair_terms: 2*Nq + Nv > 12288");
        if (dtype == TRIAD_DTYPE_BF16)
            pospair::sparsity_kernel<__nv_bfloat16><<<blocks, pospair::kThreads, smem, st>>>(raw, temperature, threshold, B, Nq, Nv, G, partials);
        else
            pospair::sparsity_kernel<float><<<blocks, pospair::kThreads, smem, st>>>(raw, temperature, threshold, B, Nq, Nv, G, partials);
        TRIAD_LAUNCH_CHECK("pospair sparsity_kernel");
    }
    pospair::finish_kernel<<<1, 32, 0, st>>>(partials, blocks, scale0, sums);
    TRIAD_LAUNCH_CHECK("pospair finish_kernel");
    return TRIAD_OK;
}

extern "C" int triad_scale(const void* x, void* y, size_t n, int dtype, const float* scale, void* stream) {
    if (!x || !y || !scale) return fail_msg(TRIAD_ERR_BAD_ARG, "scale: null pointer");
    if (dtype != TRIAD_DTYPE_F32 && dtype != TRIAD_DTYPE_BF16) return fail_msg(TRIAD_ERR_BAD_ARG, "scale: dtype");
    if (n == 0) return TRIAD_OK;
    if (((uintptr_t)x | (uintptr_t)y) & 15) return fail_msg(TRIAD_ERR_ALIGNMENT, "scale: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvec = n / (dtype == TRIAD_DTYPE_BF16 ? 8 : 4);
    size_t want = (nvec + pospair::kThreads - 1) / pospair::kThreads;
    const int blocks = (int)(want < 1 ? 1 : (want > (size_t)(148 * 16) ? (size_t)(148 * 16) : want));
    if (dtype == TRIAD_DTYPE_BF16) pospair::scale_kernel<__nv_bfloat16><<<blocks, pospair::kThreads, 0, st>>>(x, y, n, scale);
    else pospair::scale_kernel<float><<<blocks, pospair::kThreads, 0, st>>>(x, y, n, scale);
    TRIAD_LAUNCH_CHECK("scale_kernel");
    return TRIAD_OK;
}
