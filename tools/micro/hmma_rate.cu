// Micro-benchmark: issue rate of the legacy mma.sync.m16n8k16 bf16 path on sm_100a (cycles per MMA per SM
// sub-partition), to decide whether a "diagonal weight matrix" MMA can replace unpack+FFMA2 in the dq gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(float* out, long long* cyc, int iters) {
    uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
    float d[8][4] = {};
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[u][0]), "+f"(d[u][1]), "+f"(d[u][2]), "+f"(d[u][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0;
    for (int u = 0; u < 8; ++u) for (int e = 0; e < 4; ++e) s += d[u][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 256, 512}) {
        const int iters = 2000;
        k<<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        const double per_warp = (double)h[0] / (iters * 8.0);
        const int warps_per_smsp = threads / 32 / 4;
        printf("threads %d: %.2f cycles per MMA per warp, %.2f cycles per MMA per SM sub-partition (%d warps each) -> %.0f dense flop/clk/SM\n",
               threads, per_warp, per_warp / warps_per_smsp, warps_per_smsp, 4096.0 * 4 / (per_warp / warps_per_smsp));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
