// Row packing for masked (text) queries — src/model.py:509-512 multiplies the row maxima of padded
// tokens by 0, so those rows never reach clip, dq or dv; at the CC3M shape (77-token pad, captions of
// 8..77 tokens) they are 45 % of the GEMM.  The packed path drops them BEFORE the tensor cores:
//
//   rowmap[k] = original row (i*Nq + a) of the k-th row with a non-zero weight, in original order
//   off[i]    = number of such rows in queries < i          (off[Bq] = M' = rows kept)
//   qp[k,:]   = q[rowmap[k], :]                             (dense copy: TMA tiles need contiguous rows)
//
// Everything is computed on the device from row_scale (no host synchronisation: kernels that consume the
// packed rows read M' from device memory and are launched for the worst case M).
#include "common.cuh"

namespace triad {

// one warp per query: count the kept rows
__global__ void __launch_bounds__(256)
pack_count_kernel(const float* __restrict__ row_scale, int Bq, int Nq, int* __restrict__ cnt) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= Bq) return;
    int c = 0;
    for (int a = lane; a < Nq; a += 32) c += row_scale[(size_t)i * Nq + a] != 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[i] = c;
}

// single CTA: exclusive scan of cnt[Bq] -> off[Bq+1]
__global__ void __launch_bounds__(1024)
pack_scan_kernel(const int* __restrict__ cnt, int Bq, int* __restrict__ off) {
    __shared__ int tile[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < Bq; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int v = i < Bq ? cnt[i] : 0;
        tile[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int x = threadIdx.x >= (unsigned)o ? tile[threadIdx.x - o] : 0;
            __syncthreads();
            tile[threadIdx.x] += x;
            __syncthreads();
        }
        if (i < Bq) off[i] = carry + tile[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += tile[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[Bq] = carry;
}

// one warp per query: positions of the kept rows (ballot prefix), original order preserved
__global__ void __launch_bounds__(256)
pack_map_kernel(const float* __restrict__ row_scale, const int* __restrict__ off, int Bq, int Nq, int* __restrict__ rowmap) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= Bq) return;
    int k = off[i];
    for (int a0 = 0; a0 < Nq; a0 += 32) {
        const int a = a0 + lane;
        const bool keep = a < Nq && row_scale[(size_t)i * Nq + a] != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) rowmap[k + __popc(m & ((1u << lane) - 1u))] = i * Nq + a;
        k += __popc(m);
    }
}

// qp[k,:] = q[rowmap[k],:] for k < M' (16-byte chunks; rows >= M' are left as they are: they only feed
// accumulator rows nobody reads)
__global__ void __launch_bounds__(256)
pack_copy_kernel(const uint4* __restrict__ q, const int* __restrict__ rowmap, const int* __restrict__ Mp, int chunks_per_row,
                 uint4* __restrict__ qp) {
    const int Mk = *Mp;
    const long long total = (long long)Mk * chunks_per_row;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int k = (int)(t / chunks_per_row), c = (int)(t - (long long)k * chunks_per_row);
        qp[(size_t)k * chunks_per_row + c] = __ldg(&q[(size_t)rowmap[k] * chunks_per_row + c]);
    }
}

// ---- 8-row groups for the pipelined dq (bwd_dq_pipe.cu): the groups (i, a0 = 0, 8, 16, ...) that hold at least one
// kept row, as padded row indices x0 = i*nq_pad + a0, in order; goff[Bq] = number of active groups -----------------
__device__ __forceinline__ bool group_active(const float* __restrict__ rs, int Nq, int a0) {
    bool any = false;
    for (int a = a0; a < min(a0 + 8, Nq); ++a) any |= rs[a] != 0.f;
    return any;
}
__global__ void __launch_bounds__(256)
pack_group_count_kernel(const float* __restrict__ row_scale, int Bq, int Nq, int* __restrict__ cnt) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= Bq) return;
    int c = 0;
    for (int a0 = lane * 8; a0 < Nq; a0 += 256) c += group_active(row_scale + (size_t)i * Nq, Nq, a0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[i] = c;
}
__global__ void __launch_bounds__(256)
pack_group_map_kernel(const float* __restrict__ row_scale, const int* __restrict__ goff, int Bq, int Nq, int nq_pad,
                      int* __restrict__ glist) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= Bq) return;
    int k = goff[i];
    for (int base = 0; base < Nq; base += 256) {
        const int a0 = base + lane * 8;
        const bool keep = a0 < Nq && group_active(row_scale + (size_t)i * Nq, Nq, a0);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) glist[k + __popc(m & ((1u << lane) - 1u))] = i * nq_pad + a0;
        k += __popc(m);
    }
}

// maps (inside the pack_map_bytes region): goff[Bq+1] | glist[Bq*ceil(Nq/8)] | cnt[Bq]
int launch_pack_groups(const float* row_scale, int Bq, int Nq, void* maps, cudaStream_t st) {
    int* goff = (int*)maps;
    int* glist = goff + Bq + 1;
    int* cnt = glist + (size_t)Bq * ceil_div(Nq, 8);
    const int blocks = ceil_div(Bq * 32, 256);
    pack_group_count_kernel<<<blocks, 256, 0, st>>>(row_scale, Bq, Nq, cnt);
    TRIAD_LAUNCH_CHECK("pack_group_count_kernel");
    pack_scan_kernel<<<1, 1024, 0, st>>>(cnt, Bq, goff);
    TRIAD_LAUNCH_CHECK("pack_scan_kernel");
    pack_group_map_kernel<<<blocks, 256, 0, st>>>(row_scale, goff, Bq, Nq, nq_padded(Nq), glist);
    TRIAD_LAUNCH_CHECK("pack_group_map_kernel");
    return TRIAD_OK;
}

// scratch layout of the packing maps: off[Bq+1] | rowmap[M] (ints; off[Bq] = M')
size_t pack_map_bytes(int Bq, int Nq) { return align_up(((size_t)Bq + 1 + (size_t)Bq * Nq) * 4 + (size_t)Bq * 4, 256); }

int launch_pack_map(const float* row_scale, int Bq, int Nq, void* maps, cudaStream_t st) {
    int* off = (int*)maps;
    int* rowmap = off + Bq + 1;
    int* cnt = rowmap + (size_t)Bq * Nq;
    const int blocks = ceil_div(Bq * 32, 256);
    pack_count_kernel<<<blocks, 256, 0, st>>>(row_scale, Bq, Nq, cnt);
    TRIAD_LAUNCH_CHECK("pack_count_kernel");
    pack_scan_kernel<<<1, 1024, 0, st>>>(cnt, Bq, off);
    TRIAD_LAUNCH_CHECK("pack_scan_kernel");
    pack_map_kernel<<<blocks, 256, 0, st>>>(row_scale, off, Bq, Nq, rowmap);
    TRIAD_LAUNCH_CHECK("pack_map_kernel");
    return TRIAD_OK;
}

int launch_pack_copy(const void* q, const void* maps, int Bq, int Nq, int D, int elt_bytes, void* qp, cudaStream_t st) {
    const int* off = (const int*)maps;
    const int* rowmap = off + Bq + 1;
    const int chunks = D * elt_bytes / 16;
    pack_copy_kernel<<<148 * 8, 256, 0, st>>>((const uint4*)q, rowmap, off + Bq, chunks, (uint4*)qp);
    TRIAD_LAUNCH_CHECK("pack_copy_kernel");
    return TRIAD_OK;
}

}  // namespace triad
