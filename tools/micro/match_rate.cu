// Throughput of __match_any_sync (MATCH.ANY) on sm_100a: cycles per warp instruction with 8 warps / SM-quarter busy,
// for keys with few / many distinct values per warp.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a match_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int kMod>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters) {
    uint32_t key = (threadIdx.x * 2654435761u >> 7) % kMod, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t m = __match_any_sync(0xffffffffu, key);
        acc += __popc(m & ((1u << (threadIdx.x & 31)) - 1));
        key = (key + (m & 3) + 1) % kMod;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("kMod %3d: %.1f cycles per loop iteration (8 warps per CTA, 1 CTA/SM)\n", kMod, double(t1 - t0) / iters);
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}
int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 256 * 4);
    k<2><<<148, 256>>>(out, 4096); k<8><<<148, 256>>>(out, 4096); k<32><<<148, 256>>>(out, 4096); k<256><<<148, 256>>>(out, 4096);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0;
}
