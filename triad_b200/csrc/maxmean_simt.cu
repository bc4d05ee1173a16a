// CUDA-core (fp32 FFMA) forward of the max-mean similarity, plus the small helper kernels
// shared by both forward paths (row weights, deterministic clip finalisation).
//
// This kernel is the fp32-input path (the reference's use_amp=False configuration, SURVEY.md
// cfg 1: 1e-4 parity needs true fp32 products, which kind::tf32 tensor-core MMAs do not give)
// and the on-device cross-check of the tcgen05 kernel in maxmean_tc.cu (same rounding model,
// different accumulation order).  Replaces src/model.py:384-391 / :502-512.
#include "common.cuh"

namespace triad {

// ---------------------------------------------------------------------------------------
// row_scale
// ---------------------------------------------------------------------------------------
__global__ void row_scale_kernel(const int64_t* __restrict__ mask, int Bq, int Nq,
                                 float* __restrict__ row_scale) {
    // one warp per query
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Bq) return;
    if (mask == nullptr) {
        const float w = 1.0f / (float)Nq;
        for (int t = lane; t < Nq; t += 32) row_scale[(size_t)warp * Nq + t] = w;
        return;
    }
    const int64_t* m = mask + (size_t)warp * Nq;
    float cnt = 0.f;
    for (int t = lane; t < Nq; t += 32) cnt += (float)m[t];      // mask.float() (model.py:509)
    cnt = warp_sum(cnt);
    const float denom = fmaxf(cnt, 1e-7f);                       // clamp(min=1e-7) (model.py:511)
    for (int t = lane; t < Nq; t += 32) row_scale[(size_t)warp * Nq + t] = (float)m[t] / denom;
}

int launch_row_scale(const int64_t* mask, int Bq, int Nq, float* row_scale, cudaStream_t st) {
    const int threads = 256;
    const int blocks = ceil_div(Bq * 32, threads);
    row_scale_kernel<<<blocks, threads, 0, st>>>(mask, Bq, Nq, row_scale);
    TRIAD_LAUNCH_CHECK("row_scale_kernel");
    return TRIAD_OK;
}

// ---------------------------------------------------------------------------------------
// finalize: clip[i][j] = sum over the 32-row groups that hold rows of query i
// ---------------------------------------------------------------------------------------
// A tripped pipeline watchdog (abort flag != 0: the forward kernel drained out and left partial sums unwritten) must
// not pass silently: clip becomes NaN, so the loss — and everything the caller logs — is NaN.
__global__ void finalize_clip_kernel(const float* __restrict__ part, int Bq, int Bv, int Nq,
                                     int G, int S, float* __restrict__ clip, const int* __restrict__ abort_flag) {
    const int nib = (Bq + blockDim.x - 1) / blockDim.x;
    const int i = (blockIdx.x % nib) * blockDim.x + threadIdx.x;
    const int j = blockIdx.x / nib;
    if (i >= Bq) return;
    const int r0 = i * Nq, r1 = r0 + Nq - 1;
    const int g0 = r0 >> 5, g1 = r1 >> 5;
    const float* pj = part + (size_t)j * G * S;
    float acc = 0.f;
    for (int g = g0; g <= g1; ++g) {
        const int qfirst = (g * 32) / Nq;
        acc += pj[(size_t)g * S + (i - qfirst)];
    }
    if (abort_flag && *abort_flag != 0) acc = __int_as_float(0x7fc00000);
    clip[(size_t)i * Bv + j] = acc;
}

int launch_finalize_clip(const float* part, int Bq, int Bv, int Nq, float* clip, const int* abort_flag, cudaStream_t st) {
    PartLayout pl = part_layout(Bq * Nq, Nq);
    dim3 grid((unsigned)(ceil_div(Bq, 128) * Bv));
    finalize_clip_kernel<<<grid, 128, 0, st>>>(part, Bq, Bv, Nq, pl.G, pl.S, clip, abort_flag);
    TRIAD_LAUNCH_CHECK("finalize_clip_kernel");
    return TRIAD_OK;
}

// packed rows: clip[i][j] = sum of the query's pieces (one per 32-row group its kept rows touch), ascending
__global__ void finalize_clip_packed_kernel(const float* __restrict__ part, const int* __restrict__ off, int Bq, int Bv,
                                            int pieces, float* __restrict__ clip, const int* __restrict__ abort_flag) {
    const int nib = (Bq + blockDim.x - 1) / blockDim.x;
    const int i = (blockIdx.x % nib) * blockDim.x + threadIdx.x;
    const int j = blockIdx.x / nib;
    if (i >= Bq) return;
    const int k0 = off[i], k1 = off[i + 1];
    float acc = 0.f;
    if (k1 > k0) {
        const int np = ((k1 - 1) >> 5) - (k0 >> 5) + 1;
        const float* pp = part + ((size_t)j * Bq + i) * pieces;
        for (int s = 0; s < np; ++s) acc += pp[s];
    }
    if (abort_flag && *abort_flag != 0) acc = __int_as_float(0x7fc00000);
    clip[(size_t)i * Bv + j] = acc;
}

int launch_finalize_clip_packed(const float* part, const int* pack_off, int Bq, int Bv, int Nq, float* clip,
                                const int* abort_flag, cudaStream_t st) {
    dim3 grid((unsigned)(ceil_div(Bq, 128) * Bv));
    finalize_clip_packed_kernel<<<grid, 128, 0, st>>>(part, pack_off, Bq, Bv, packed_pieces(Nq), clip, abort_flag);
    TRIAD_LAUNCH_CHECK("finalize_clip_packed_kernel");
    return TRIAD_OK;
}

// ---------------------------------------------------------------------------------------
// SIMT forward: one CTA = one 32-row group x one image; 256 threads as 16 (rows/2) x 16 (cols/4)
// ---------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

constexpr int kSimtRows = 32;
constexpr int kSimtCols = 64;
constexpr int kSimtK = 32;

template <typename T, bool kBF16Round, typename IdxT>
__global__ void __launch_bounds__(256)
maxmean_simt_kernel(const T* __restrict__ q, const T* __restrict__ v,
                    const float* __restrict__ row_scale, const float* __restrict__ Tptr, int inv_T,
                    int M, int Bv, int Nq, int Nv, int D, int G, int S, int nq_pad,
                    float* __restrict__ part, IdxT* __restrict__ idx) {
    __shared__ float Qs[kSimtK][kSimtRows + 1];
    __shared__ __align__(16) float Vs[kSimtK][kSimtCols + 4];
    __shared__ float bestv[kSimtRows][16];
    __shared__ int besti[kSimtRows][16];

    const int g = blockIdx.x % G, j = blockIdx.x / G;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = g * kSimtRows;
    float Tval = *Tptr;
    if (inv_T) Tval = 1.0f / Tval;

    float rbest[2] = {-INFINITY, -INFINITY};
    int ibest[2] = {0, 0};

    const T* vj = v + (size_t)j * Nv * D;
    for (int c0 = 0; c0 < Nv; c0 += kSimtCols) {
        float acc[2][4] = {};
        for (int k0 = 0; k0 < D; k0 += kSimtK) {
            // Q tile: 32 rows x 32 k  (4 elements per thread)
            {
                const int r = tid >> 3, kk = (tid & 7) * 4;
                const int gr = row0 + r;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k0 + kk + e;
                    Qs[kk + e][r] = (gr < M && k < D) ? to_f32<T>(q[(size_t)gr * D + k]) : 0.f;
                }
            }
            // V tile: 64 cols x 32 k  (8 elements per thread)
            {
                const int c = tid >> 2, kk = (tid & 3) * 8;
                const int gc = c0 + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = k0 + kk + e;
                    Vs[kk + e][c] = (gc < Nv && k < D) ? to_f32<T>(vj[(size_t)gc * D + k]) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kSimtK; ++kk) {
                const float a0 = Qs[kk][ty * 2], a1 = Qs[kk][ty * 2 + 1];
                const float4 b = *reinterpret_cast<const float4*>(&Vs[kk][tx * 4]);
                acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
                acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
                acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
                acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int col = c0 + tx * 4 + cc;
                if (col < Nv) {
                    const float s = round_sim<kBF16Round>(acc[rr][cc], Tval);
                    if (s > rbest[rr]) { rbest[rr] = s; ibest[rr] = col; }   // strict: keeps the first
                }
            }
    }
    bestv[ty * 2][tx] = rbest[0]; besti[ty * 2][tx] = ibest[0];
    bestv[ty * 2 + 1][tx] = rbest[1]; besti[ty * 2 + 1][tx] = ibest[1];
    __syncthreads();
    if (tid < 32) {
        const int lane = tid, r = row0 + lane;
        float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const float x = bestv[lane][t]; const int xi = besti[lane][t];
            if (x > bv || (x == bv && xi < bi)) { bv = x; bi = xi; }
        }
        float val = 0.f;
        if (r < M) {
            val = bv * row_scale[r];
            if (idx != nullptr) {
                const int qi = r / Nq;
                idx[((size_t)j * (M / Nq) + qi) * nq_pad + (r - qi * Nq)] = (IdxT)bi;
            }
        }
        store_group_partials(part, j, g, G, S, row0, M, Nq, val, lane);
    }
}

template <typename T, bool R>
static int launch_simt_t(const void* q, const void* v, const float* row_scale, const float* Tp,
                         int inv_T, int M, int Bv, int Nq, int Nv, int D, float* part, void* idx,
                         cudaStream_t st) {
    PartLayout pl = part_layout(M, Nq);
    if ((long long)pl.G * Bv > 0x7fffffffLL) return fail_msg(TRIAD_ERR_UNSUPPORTED, "SIMT forward: grid too large");
    dim3 grid((unsigned)(pl.G * Bv));
    if (Nv <= 256)
        maxmean_simt_kernel<T, R, uint8_t><<<grid, 256, 0, st>>>(
            (const T*)q, (const T*)v, row_scale, Tp, inv_T, M, Bv, Nq, Nv, D, pl.G, pl.S, nq_padded(Nq), part, (uint8_t*)idx);
    else
        maxmean_simt_kernel<T, R, uint16_t><<<grid, 256, 0, st>>>(
            (const T*)q, (const T*)v, row_scale, Tp, inv_T, M, Bv, Nq, Nv, D, pl.G, pl.S, nq_padded(Nq), part, (uint16_t*)idx);
    TRIAD_LAUNCH_CHECK("maxmean_simt_kernel");
    return TRIAD_OK;
}

int launch_maxmean_simt(const void* q, const void* v, const float* row_scale, const float* T,
                        int inv_T, int M, int Bv, int Nq, int Nv, int D, int dtype,
                        float* part, void* idx, cudaStream_t st) {
    if (dtype == TRIAD_DTYPE_BF16)
        return launch_simt_t<__nv_bfloat16, true>(q, v, row_scale, T, inv_T, M, Bv, Nq, Nv, D, part, idx, st);
    return launch_simt_t<float, false>(q, v, row_scale, T, inv_T, M, Bv, Nq, Nv, D, part, idx, st);
}

}  // namespace triad
