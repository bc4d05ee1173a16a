// Write-bandwidth probe no. 2 for the dense regulariser's N (8.4 GB at cfg 2).  Unlike write_pattern.cu every store
// instruction writes FULL 128-byte lines (8 lanes x 16 B per row, 4 rows per warp instruction, as a TMA store box of
// 32 rows x 128 B does) and the values are incompressible (address hash), so the numbers show what the layout costs:
//   (a) row-major   N[M][Bv*Nv]                      box rows 128 KB apart            (what the forward emits today)
//   (b) tile-major  N[M/128][Bv][128][Nv]            box rows 512 B apart inside a contiguous 64 KB block
//   (c) box-major   N[M/128][Bv][Nv/64][128][64]     a 32 x 64 box is 4 KB contiguous
//   (d) linear      plain streaming write of the same bytes (upper bound)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a write_pattern2.cu -o write_pattern2
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint4 val_of(size_t a) {
    uint32_t x = (uint32_t)a * 2654435761u ^ (uint32_t)(a >> 20);
    return make_uint4(x, x * 747796405u + 1u, x ^ 0x9e3779b9u, x * 2891336453u);
}

// grid = 148 x CTAs/SM; 8 warps; warp w of a CTA owns rows [32 w', 32 w'+32) of a 128-row tile half ... simplified:
// a CTA walks (tile, image) pairs; its 8 warps cover 4 row groups x 2 column halves (2 boxes of 64 columns each).
template <int kLayout>
__global__ void __launch_bounds__(256) wr(uint4* __restrict__ N, int n_tiles, int Bv, int Nv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = warp & 3, ch = warp >> 2;                   // rows 32 rg.., column boxes {2 ch, 2 ch + 1}
    const int r4 = lane >> 3, c16 = lane & 7;                  // 4 rows per instruction, 8 x 16 B per row
    const size_t pitch16 = (size_t)Bv * Nv * 2 / 16;
    const long long n_items = (long long)n_tiles * Bv;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int t = (int)(it / Bv), j = (int)(it % Bv);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int box = ch * 2 + b;                        // 64-column box of the image's 256 columns
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = rg * 32 + k * 4 + r4;          // row in the 128-row tile
                size_t a;
                if (kLayout == 0) a = ((size_t)t * 128 + row) * pitch16 + (size_t)j * (Nv / 8) + box * 8 + c16;
                else if (kLayout == 1) a = ((size_t)t * Bv + j) * (size_t)(128 * Nv / 8) + (size_t)row * (Nv / 8) + box * 8 + c16;
                else a = (((size_t)t * Bv + j) * (Nv / 64) + box) * (size_t)(128 * 8) + (size_t)row * 8 + c16;
                N[a] = val_of(a);
            }
        }
    }
}

__global__ void __launch_bounds__(256) linear(uint4* __restrict__ N, size_t n16) {
    for (size_t a = (size_t)blockIdx.x * 256 + threadIdx.x; a < n16; a += (size_t)gridDim.x * 256) N[a] = val_of(a);
}

int main() {
    const int M = 64000, Bv = 256, Nv = 256, n_tiles = M / 128;
    const size_t bytes = (size_t)M * Bv * Nv * 2;
    uint4* N; cudaMalloc(&N, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-34s %.3f ms  %.2f TB/s\n", name, best, bytes / best / 1e9);
    };
    for (int per_sm : {1, 2, 4, 8}) {
        printf("-- %d CTAs per SM\n", per_sm);
        run("row-major (pitch 128 KB)", [&] { wr<0><<<148 * per_sm, 256>>>(N, n_tiles, Bv, Nv); });
        run("tile-major (64 KB blocks)", [&] { wr<1><<<148 * per_sm, 256>>>(N, n_tiles, Bv, Nv); });
        run("box-major (4 KB boxes)", [&] { wr<2><<<148 * per_sm, 256>>>(N, n_tiles, Bv, Nv); });
        run("linear", [&] { linear<<<148 * per_sm, 256>>>(N, bytes / 16); });
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
