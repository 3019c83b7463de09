// Throughput of float reductions into a [nvert][ncat] table (the vertex-adjoint pattern of the reverse epilogue): every warp
// adds 32 consecutive floats to a random row.  Variants: scalar red.global.add.f32 per lane, red.global.add.v2.f32 on even
// lanes, red.global.add.v4.f32 on every fourth lane (values gathered with shuffles).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/redbw tools/redbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int VEC>
__global__ void red_kernel(float* table, int nvert, int ncat, int64_t warp_items, int items_per_warp) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    for (int it = 0; it < items_per_warp; ++it) {
        const int64_t item = w * items_per_warp + it;
        if (item >= warp_items) return;
        const uint32_t h = hash32((uint32_t)item * 2654435761u + 12345u);
        const int v = h % nvert;
        const int f0 = ((h >> 16) % (ncat / 32)) * 32;
        float* p = table + (int64_t)v * ncat + f0 + lane;
        const float x = 1.0f + lane * 1e-3f;
        if (VEC == 1) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
        } else if (VEC == 2) {
            const float x1 = __shfl_down_sync(0xffffffffu, x, 1);
            if ((lane & 1) == 0) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(x1) : "memory");
        } else {
            const float x1 = __shfl_down_sync(0xffffffffu, x, 1), x2 = __shfl_down_sync(0xffffffffu, x, 2),
                        x3 = __shfl_down_sync(0xffffffffu, x, 3);
            if ((lane & 3) == 0)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(x1), "f"(x2), "f"(x3) : "memory");
        }
    }
}

template <int VEC>
static void run(float* table, int nvert, int ncat, int64_t warp_items) {
    const int ipw = 64;
    const int64_t warps = (warp_items + ipw - 1) / ipw;
    const int block = 256;
    const int64_t grid = (warps * 32 + block - 1) / block;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(table, 0, (size_t)nvert * ncat * 4);
    red_kernel<VEC><<<(unsigned)grid, block>>>(table, nvert, ncat, warp_items, ipw);
    cudaEventRecord(e0);
    red_kernel<VEC><<<(unsigned)grid, block>>>(table, nvert, ncat, warp_items, ipw);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double sum = 0.0;
    {   // checksum: every variant must add the same total
        float* h = (float*)malloc((size_t)nvert * ncat * 4);
        cudaMemcpy(h, table, (size_t)nvert * ncat * 4, cudaMemcpyDeviceToHost);
        for (int64_t i = 0; i < (int64_t)nvert * ncat; ++i) sum += h[i];
        free(h);
    }
    printf("nvert %6d ncat %4d  vec %d: %8.3f ms  %7.1f G float adds/s  (%s)  checksum %.6e\n", nvert, ncat, VEC, ms,
           warp_items * 32.0 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()), sum);
}

int main() {
    const int64_t warp_items = (int64_t)1 << 23;      // 2^23 x 32 = 268 M float adds (one config-3 chunk of layer 0)
    for (int nvert : {8192, 32768, 262144}) {
        const int ncat = 992;
        float* table;
        cudaMalloc(&table, (size_t)nvert * ncat * 4);
        run<1>(table, nvert, ncat, warp_items);
        run<2>(table, nvert, ncat, warp_items);
        run<4>(table, nvert, ncat, warp_items);
        cudaFree(table);
    }
    return 0;
}
