// SUPERSEDED by write_pattern2.cu: this probe's stores fill a 128-byte line with 16 separate 16-byte pieces per warp
// instruction stream (partial-sector writes), so it measures the store instruction pattern, not the layout (1.8 / 2.2 TB/s
// here against 5.2 / 6.3 TB/s with full-line stores).
// Write-bandwidth probe for the dense regulariser's N (8.4 GB at cfg 2): the same bytes written
//   (a) row-major  N[M][Bv*Nv]   — a CTA owns 128 rows and appends 512 B per row per image (pitch 128 KB): what the
//       forward's epilogue does;
//   (b) tile-major N[M/128][Bv][128][Nv] — a CTA writes one contiguous 64 KB block per (row tile, image);
// each with default and streaming (evict-first) stores.  nvcc -O3 -arch=sm_100a write_pattern.cu -o write_pattern
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int kTiled, int kStream>
__global__ void __launch_bounds__(256) wr(uint4* __restrict__ N, int n_tiles, int Bv, int Nv) {
    const int row = threadIdx.x >> 1, half = threadIdx.x & 1;                 // 128 rows x 2 halves of 256 B
    const size_t ld16 = (size_t)Bv * Nv * 2 / 16;                             // row pitch in uint4
    const int v16 = Nv * 2 / 16;                                              // 16-byte pieces per image row (32)
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int j = 0; j < Bv; ++j) {
            uint4* p;
            if (kTiled) p = N + ((size_t)t * Bv + j) * (size_t)(128 * v16) + (size_t)row * v16 + half * (v16 / 2);
            else p = N + ((size_t)t * 128 + row) * ld16 + (size_t)j * v16 + half * (v16 / 2);
            const uint4 val = make_uint4(t, j, row, half);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (kStream) __stcs(p + k, val); else p[k] = val;
            }
        }
    }
}

int main() {
    const int M = 64000, Bv = 256, Nv = 256, n_tiles = M / 128;
    const size_t bytes = (size_t)M * Bv * Nv * 2;
    uint4* N; cudaMalloc(&N, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto kern) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0); kern<<<148, 256>>>(N, n_tiles, Bv, Nv); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-28s %.3f ms  %.2f TB/s\n", name, best, bytes / best / 1e9);
    };
    run("row-major, default", wr<0, 0>);
    run("row-major, streaming", wr<0, 1>);
    run("tile-major, default", wr<1, 0>);
    run("tile-major, streaming", wr<1, 1>);
    cudaEventRecord(e0); cudaMemsetAsync(N, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%-28s %.3f ms  %.2f TB/s\n", "cudaMemset", ms, bytes / ms / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
