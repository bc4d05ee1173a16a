// Backward of the max-mean similarity through the saved argmax indices (SURVEY.md §8 a5;
// replaces autograd's backward of src/model.py:387-391, which materialises a zero dS of the
// full Bq x Bv x Nq x Nv size, scatters into it and runs two dense bmm backward GEMMs).
//
//   dq[r,:]   = T*row_scale[r] * sum_j g[i(r),j] * v[j, idx[j][r], :]        (gather)
//   dv[j,p,:] = sum_{r: idx[j][r]==p} T*row_scale[r]*g[i(r),j] * q[r,:]      (scatter)
//   dT        = sum_ij g[i,j]*clip[i,j] / T
//
// Both passes are HBM/L2- and issue-bound CUDA-core work (2*D MACs per (row,image) pair, i.e.
// 2/Nv of the forward's flops); they are not reshaped into one-hot GEMMs.
#include "common.cuh"

namespace triad {

template <typename T> struct Vec16;            // one 16-byte chunk of a row
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int kElems = 8;
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = __uint_as_float(w[k] << 16);
            f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
        uint4 u;
        uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template <> struct Vec16<float> {
    static constexpr int kElems = 4;
    __device__ static __forceinline__ void load(const float* p, float (&f)[4]) {
        const float4 u = *reinterpret_cast<const float4*>(p);
        f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
    }
    __device__ static __forceinline__ void store(float* p, const float (&f)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};

// ---------------------------------------------------------------------------------------
// dq: one CTA = one 32-row group; 8 warps x 4 rows; lanes own 16-byte chunks of D
// ---------------------------------------------------------------------------------------
constexpr int kDqRows = 32;
constexpr int kDqJ = 32;          // images staged per step

template <typename T, typename IdxT, int KCH>
__global__ void __launch_bounds__(256)
dq_gather_kernel(const T* __restrict__ v, const IdxT* __restrict__ idx, const float* __restrict__ g,
                 const float* __restrict__ row_scale, const float* __restrict__ Tptr,
                 int M, int Bv, int Nq, int Nv, int D, T* __restrict__ dq) {
    constexpr int E = Vec16<T>::kElems;
    __shared__ IdxT idx_s[kDqJ][kDqRows];
    __shared__ float w_s[kDqRows][kDqJ + 1];

    const int row0 = blockIdx.x * kDqRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = D / E;

    float acc[4][KCH][E];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < KCH; ++b)
#pragma unroll
            for (int c = 0; c < E; ++c) acc[a][b][c] = 0.f;

    for (int j0 = 0; j0 < Bv; j0 += kDqJ) {
        __syncthreads();
        // stage idx[j0..j0+31][row0..row0+31] and the matching g entries
        for (int t = threadIdx.x; t < kDqJ * kDqRows; t += 256) {
            const int jj = t / kDqRows, rr = t % kDqRows;
            const int j = j0 + jj, r = row0 + rr;
            IdxT p = 0; float w = 0.f;
            if (j < Bv && r < M) {
                p = idx[(size_t)j * M + r];
                w = g[(size_t)(r / Nq) * Bv + j];
            }
            idx_s[jj][rr] = p;
            w_s[rr][jj] = w;
        }
        __syncthreads();
        const int jn = min(kDqJ, Bv - j0);
        for (int jj = 0; jj < jn; ++jj) {
            const T* vj = v + (size_t)(j0 + jj) * Nv * D;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int rr = warp * 4 + a;
                const float w = w_s[rr][jj];
                const T* src = vj + (size_t)idx_s[jj][rr] * D;
#pragma unroll
                for (int b = 0; b < KCH; ++b) {
                    const int ch = lane + 32 * b;
                    if (ch < nchunk) {
                        float f[E];
                        Vec16<T>::load(src + ch * E, f);
#pragma unroll
                        for (int c = 0; c < E; ++c) acc[a][b][c] = fmaf(w, f[c], acc[a][b][c]);
                    }
                }
            }
        }
    }
    const float Tval = *Tptr;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = row0 + warp * 4 + a;
        if (r >= M) continue;
        const float s = Tval * row_scale[r];
#pragma unroll
        for (int b = 0; b < KCH; ++b) {
            const int ch = lane + 32 * b;
            if (ch < nchunk) {
                float f[E];
#pragma unroll
                for (int c = 0; c < E; ++c) f[c] = acc[a][b][c] * s;
                Vec16<T>::store(dq + (size_t)r * D + ch * E, f);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// dv (baseline): one CTA = one image x one D-slice of 32*kDvWarps dims; fp32 accumulators for
// every patch live in shared memory.  Each warp owns 32 dims (one per lane) of ALL patches and
// streams ALL M rows, so no two threads ever touch the same accumulator: no atomics, and the
// summation order (row order) is fixed => deterministic.
// ---------------------------------------------------------------------------------------
constexpr int kDvWarps = 4;

template <typename T, typename IdxT, typename OutT>
__global__ void __launch_bounds__(kDvWarps * 32)
dv_scatter_kernel(const T* __restrict__ q, const IdxT* __restrict__ idx, const float* __restrict__ g,
                  const float* __restrict__ row_scale, const float* __restrict__ Tptr,
                  int M, int Bv, int Nq, int Nv, int D, int DS, OutT* __restrict__ dv) {
    extern __shared__ float acc_s[];                 // [Nv][DS]
    const int nslice = D / DS;
    const int j = blockIdx.x / nslice;
    const int d0 = (blockIdx.x % nslice) * DS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float Tval = *Tptr;

    for (int t = threadIdx.x; t < Nv * DS; t += kDvWarps * 32) acc_s[t] = 0.f;
    __syncthreads();

    const IdxT* idxj = idx + (size_t)j * M;
    const int dcol = warp * 32 + lane;               // this thread's column inside the slice
    const bool active = dcol < DS;
    for (int base = 0; base < M; base += 32) {
        const int r = base + lane;
        int p = 0; float w = 0.f;
        if (r < M) {
            p = (int)idxj[r];
            w = Tval * row_scale[r] * g[(size_t)(r / Nq) * Bv + j];
        }
        const int nrow = min(32, M - base);
        const T* src = q + (size_t)base * D + d0 + dcol;
        for (int l = 0; l < nrow; ++l) {
            const float wl = __shfl_sync(0xffffffffu, w, l);
            const int pl = __shfl_sync(0xffffffffu, p, l);
            if (wl != 0.f && active) acc_s[pl * DS + dcol] += wl * (float)src[(size_t)l * D];
        }
    }
    __syncthreads();
    OutT* out = dv + (size_t)j * Nv * D + d0;
    for (int t = threadIdx.x; t < Nv * DS; t += kDvWarps * 32) {
        const int p = t / DS, d = t % DS;
        out[(size_t)p * D + d] = (OutT)acc_s[t];
    }
}

// ---------------------------------------------------------------------------------------
// dT = sum g*clip / T : single CTA, fixed-order tree (deterministic)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
dT_kernel(const float* __restrict__ g, const float* __restrict__ clip, size_t n,
          const float* __restrict__ Tptr, float* __restrict__ dT) {
    __shared__ double red[32];
    double a = 0.0;
    for (size_t k = threadIdx.x; k < n; k += 1024) a += (double)g[k] * (double)clip[k];
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        double b = red[threadIdx.x];
        b = warp_sum_d(b);
        if (threadIdx.x == 0) *dT = (float)(b / (double)*Tptr);
    }
}

static int pick_dv_slice(int Nv, int D) {
    // largest slice (multiple of 8 dividing D) whose accumulators fit ~160 KB of shared memory
    int best = 0;
    for (int ds = 8; ds <= D && ds <= 32 * kDvWarps; ds += 8)
        if (D % ds == 0 && (size_t)Nv * ds * 4 <= 160 * 1024) best = ds;
    return best;
}

template <typename T, typename IdxT>
static int bwd_typed(const void* q, const void* v, const void* idx, const float* g,
                     const float* clip, const float* row_scale, const float* Tp,
                     int Bq, int Bv, int Nq, int Nv, int D,
                     void* dq, void* dv, int dv_f32, float* dT, cudaStream_t st) {
    const int M = Bq * Nq;
    constexpr int E = Vec16<T>::kElems;
    if (dq) {
        const int kch = ceil_div(D / E, 32);
        const int grid = ceil_div(M, kDqRows);
#define TRIAD_DQ(K) dq_gather_kernel<T, IdxT, K><<<grid, 256, 0, st>>>( \
        (const T*)v, (const IdxT*)idx, g, row_scale, Tp, M, Bv, Nq, Nv, D, (T*)dq)
        switch (kch) {
            case 1: TRIAD_DQ(1); break;
            case 2: TRIAD_DQ(2); break;
            case 3: TRIAD_DQ(3); break;
            case 4: TRIAD_DQ(4); break;
            default: return fail_msg(TRIAD_ERR_UNSUPPORTED, "maxmean_bwd: D too large for dq kernel");
        }
#undef TRIAD_DQ
        TRIAD_LAUNCH_CHECK("dq_gather_kernel");
    }
    if (dv) {
        const int DS = pick_dv_slice(Nv, D);
        if (DS == 0) return fail_msg(TRIAD_ERR_UNSUPPORTED, "maxmean_bwd: Nv too large for dv kernel");
        const size_t smem = (size_t)Nv * DS * 4;
        const int grid = Bv * (D / DS);
        if (dv_f32 || sizeof(T) == 4) {
            auto kern = dv_scatter_kernel<T, IdxT, float>;
            TRIAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kDvWarps * 32, smem, st>>>((const T*)q, (const IdxT*)idx, g, row_scale, Tp, M, Bv, Nq, Nv, D, DS, (float*)dv);
        } else {
            auto kern = dv_scatter_kernel<T, IdxT, T>;
            TRIAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kDvWarps * 32, smem, st>>>((const T*)q, (const IdxT*)idx, g, row_scale, Tp, M, Bv, Nq, Nv, D, DS, (T*)dv);
        }
        TRIAD_LAUNCH_CHECK("dv_scatter_kernel");
    }
    if (dT) {
        dT_kernel<<<1, 1024, 0, st>>>(g, clip, (size_t)Bq * Bv, Tp, dT);
        TRIAD_LAUNCH_CHECK("dT_kernel");
    }
    return TRIAD_OK;
}

}  // namespace triad

using namespace triad;

extern "C" size_t triad_maxmean_bwd_workspace_bytes(int, int, int, int, int, int) { return 256; }

extern "C" int triad_maxmean_bwd(const void* q, const void* v, const void* idx, const float* g,
                                 const float* clip, const float* row_scale, const float* temperature,
                                 int Bq, int Bv, int Nq, int Nv, int D, int dtype,
                                 void* dq, void* dv, int dv_f32, float* dT,
                                 void* ws, size_t ws_bytes, void* stream) {
    (void)ws; (void)ws_bytes;
    if (!q || !v || !idx || !g || !row_scale || !temperature) return fail_msg(TRIAD_ERR_BAD_ARG, "maxmean_bwd: null pointer");
    if (dT && !clip) return fail_msg(TRIAD_ERR_BAD_ARG, "maxmean_bwd: dT needs clip");
    if (Bq <= 0 || Bv <= 0 || Nq <= 0 || Nv <= 0 || D <= 0 || D % 8 != 0 || Nv > 65535)
        return fail_msg(TRIAD_ERR_BAD_SHAPE, "maxmean_bwd: bad shape");
    if (dtype != TRIAD_DTYPE_F32 && dtype != TRIAD_DTYPE_BF16) return fail_msg(TRIAD_ERR_BAD_ARG, "maxmean_bwd: dtype");
    if (((uintptr_t)q | (uintptr_t)v | (uintptr_t)dq | (uintptr_t)dv) & 15) return fail_msg(TRIAD_ERR_ALIGNMENT, "maxmean_bwd: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = Nv > 256;
    if (dtype == TRIAD_DTYPE_BF16) {
        return wide ? bwd_typed<__nv_bfloat16, uint16_t>(q, v, idx, g, clip, row_scale, temperature, Bq, Bv, Nq, Nv, D, dq, dv, dv_f32, dT, st)
                    : bwd_typed<__nv_bfloat16, uint8_t>(q, v, idx, g, clip, row_scale, temperature, Bq, Bv, Nq, Nv, D, dq, dv, dv_f32, dT, st);
    }
    return wide ? bwd_typed<float, uint16_t>(q, v, idx, g, clip, row_scale, temperature, Bq, Bv, Nq, Nv, D, dq, dv, dv_f32, dT, st)
                : bwd_typed<float, uint8_t>(q, v, idx, g, clip, row_scale, temperature, Bq, Bv, Nq, Nv, D, dq, dv, dv_f32, dT, st);
}
