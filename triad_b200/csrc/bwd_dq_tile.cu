// dQ of the max-mean backward, bf16, tiled for on-chip reuse:
//
//   dq[r, :] = T * row_scale[r] * sum_j g[i(r), j] * v[j, idx[j][r], :]
//
// The arithmetic is 2*D MACs per (row, image) pair — 2/Nv of the forward — but every MAC
// needs one bf16 of a GATHERED patch row: M*Bv rows of D*2 bytes (16.8 GB at B=256, 250x256,
// D=512).  Served from L2 that is the whole cost (the generic kernel in maxmean_bwd.cu runs
// at the L2 limit).  This kernel makes the gather an ON-CHIP one:
//
//   * a CTA owns R = 512 consecutive token rows and ONE 64-element (128-byte = one cache line)
//     slice of D, and keeps the 512 x 64 fp32 accumulators in registers (128 KB = half the
//     register file of the SM) for the whole sweep over the images;
//   * for one image the 512 rows pick their winners among only Nv (256) patches, so the CTA
//     touches at most Nv distinct 128-byte lines per image — a 32 KB working set that lives in
//     L1: every line is fetched from L2 once and then re-read by the other rows that chose the
//     same patch (Nv*(1-exp(-R/Nv)) = 221 of 512 reads miss for uniformly random winners, far
//     fewer when the winners concentrate, as they do for trained embeddings);
//   * all CTAs walk the images in the same order, so the lines they miss on are hot in L2.
//
// Thread mapping: warp w owns tile rows [32w, 32w+32); its four 8-lane groups own 8 rows each;
// lane c of a group owns elements [8c, 8c+8) of the slice, so one LDG.128 per lane fetches a
// whole line per group (4 rows per warp instruction, no bank or sector waste).  The winners and
// weights g[i(r), j] of 8 images at a time are staged through shared memory (thread t stages
// tile row t), double buffered, one __syncthreads per 8 images.
#include "common.cuh"

namespace triad {
namespace dq2 {

constexpr int kRows = 512;
constexpr int kThreads = 512;
constexpr int kSlice = 64;      // bf16 elements per CTA slice (128 bytes)
constexpr int kJG = 8;          // images per staging group

__device__ __forceinline__ uint4 ldg_line(const char* p) {
    uint4 u;
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    return u;
}

__device__ __forceinline__ void fma8(float (&a)[8], float w, const uint4& u) {
    a[0] = fmaf(w, __uint_as_float(u.x << 16), a[0]);
    a[1] = fmaf(w, __uint_as_float(u.x & 0xffff0000u), a[1]);
    a[2] = fmaf(w, __uint_as_float(u.y << 16), a[2]);
    a[3] = fmaf(w, __uint_as_float(u.y & 0xffff0000u), a[3]);
    a[4] = fmaf(w, __uint_as_float(u.z << 16), a[4]);
    a[5] = fmaf(w, __uint_as_float(u.z & 0xffff0000u), a[5]);
    a[6] = fmaf(w, __uint_as_float(u.w << 16), a[6]);
    a[7] = fmaf(w, __uint_as_float(u.w & 0xffff0000u), a[7]);
}

template <typename IdxT>
__global__ void __launch_bounds__(kThreads, 1)
dq_tile_kernel(const __nv_bfloat16* __restrict__ v, const IdxT* __restrict__ idx, const float* __restrict__ g,
               const float* __restrict__ row_scale, const float* __restrict__ Tptr,
               int M, int Bv, int Nq, int Nv, int D, int nq_pad, int g_vec, __nv_bfloat16* __restrict__ dq) {
    __shared__ __align__(16) IdxT idx_s[2][kJG][kRows];
    __shared__ __align__(16) float w_s[2][kJG][kRows];

    const int t = threadIdx.x;
    const int slice = blockIdx.x;
    const int row0 = blockIdx.y * kRows;

    // ---- staging role: thread t = tile row t -------------------------------------------------
    const int rs = row0 + t;
    const bool rs_valid = rs < M;
    const int qi = rs_valid ? rs / Nq : 0;
    const size_t pitch = (size_t)(M / Nq) * nq_pad;
    const IdxT* idx_row = idx + (size_t)qi * nq_pad + (rs_valid ? rs - qi * Nq : 0);
    const float* g_row = g + (size_t)qi * Bv;

    // ---- compute role -------------------------------------------------------------------------
    const int warp = t >> 5, lane = t & 31, grp = lane >> 3, c = lane & 7;
    const int rb = warp * 32 + grp * 8;                      // first tile row of this 8-lane group
    const char* vbase = reinterpret_cast<const char*>(v) + (size_t)slice * (kSlice * 2) + c * 16;
    const size_t img_bytes = (size_t)Nv * D * 2;
    const uint32_t row_bytes = (uint32_t)D * 2u;

    float acc[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;

    auto stage = [&](int j0, int st) {
        IdxT pi[kJG];
        float pw[kJG];
        if (rs_valid && g_vec && j0 + kJG <= Bv) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(g_row + j0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(g_row + j0 + 4));
            pw[0] = a.x; pw[1] = a.y; pw[2] = a.z; pw[3] = a.w; pw[4] = b.x; pw[5] = b.y; pw[6] = b.z; pw[7] = b.w;
#pragma unroll
            for (int jj = 0; jj < kJG; ++jj) pi[jj] = __ldcs(idx_row + (size_t)(j0 + jj) * pitch);
        } else {
#pragma unroll
            for (int jj = 0; jj < kJG; ++jj) {
                const int j = j0 + jj;
                const bool ok = rs_valid && j < Bv;
                pi[jj] = ok ? __ldcs(idx_row + (size_t)j * pitch) : (IdxT)0;
                pw[jj] = ok ? __ldg(g_row + j) : 0.f;
            }
        }
#pragma unroll
        for (int jj = 0; jj < kJG; ++jj) { idx_s[st][jj][t] = pi[jj]; w_s[st][jj][t] = pw[jj]; }
    };

    stage(0, 0);
    __syncthreads();
    int st = 0;
    for (int j0 = 0; j0 < Bv; j0 += kJG, st ^= 1) {
        const bool more = j0 + kJG < Bv;
        const int jn = min(kJG, Bv - j0);
#pragma unroll
        for (int jj = 0; jj < kJG; ++jj) {
            if (jj < jn) {
                const char* vj = vbase + (size_t)(j0 + jj) * img_bytes;
                uint32_t p[8];
                if (sizeof(IdxT) == 1) {
                    const uint2 pk = *reinterpret_cast<const uint2*>(&idx_s[st][jj][rb]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { p[k] = (pk.x >> (8 * k)) & 0xffu; p[4 + k] = (pk.y >> (8 * k)) & 0xffu; }
                } else {
                    const uint4 pk = *reinterpret_cast<const uint4*>(&idx_s[st][jj][rb]);
                    p[0] = pk.x & 0xffffu; p[1] = pk.x >> 16; p[2] = pk.y & 0xffffu; p[3] = pk.y >> 16;
                    p[4] = pk.z & 0xffffu; p[5] = pk.z >> 16; p[6] = pk.w & 0xffffu; p[7] = pk.w >> 16;
                }
                const float4 wa = *reinterpret_cast<const float4*>(&w_s[st][jj][rb]);
                const float4 wb = *reinterpret_cast<const float4*>(&w_s[st][jj][rb + 4]);
                const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                uint4 d[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = ldg_line(vj + p[k] * row_bytes);
#pragma unroll
                for (int k = 0; k < 8; ++k) fma8(acc[k], w[k], d[k]);
            }
        }
        if (more) stage(j0 + kJG, st ^ 1);
        __syncthreads();
    }

    const float Tval = *Tptr;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = row0 + rb + k;
        if (r < M) {
            const float s = Tval * row_scale[r];
            uint4 o;
            uint32_t* w32 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(acc[k][2 * e] * s, acc[k][2 * e + 1] * s);
                w32[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dq + (size_t)r * D + slice * kSlice + c * 8) = o;
        }
    }
}

}  // namespace dq2

bool dq_tile_supported(int D, int dtype) { return dtype == TRIAD_DTYPE_BF16 && D % dq2::kSlice == 0; }

int launch_dq_tile(const void* v, const void* idx, int idx_bytes, const float* g, const float* row_scale,
                   const float* Tp, int M, int Bv, int Nq, int Nv, int D, void* dq, cudaStream_t st) {
    using namespace dq2;
    const dim3 grid((unsigned)(D / kSlice), (unsigned)ceil_div(M, kRows));
    const int nq_pad = nq_padded(Nq);
    const int g_vec = (Bv % 4 == 0) && (((uintptr_t)g & 15) == 0);
    if (idx_bytes == 1) {
        dq_tile_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)v, (const uint8_t*)idx, g, row_scale, Tp,
                                                          M, Bv, Nq, Nv, D, nq_pad, g_vec, (__nv_bfloat16*)dq);
    } else {
        dq_tile_kernel<uint16_t><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)v, (const uint16_t*)idx, g, row_scale, Tp,
                                                           M, Bv, Nq, Nv, D, nq_pad, g_vec, (__nv_bfloat16*)dq);
    }
    TRIAD_LAUNCH_CHECK("dq_tile_kernel");
    return TRIAD_OK;
}

}  // namespace triad
