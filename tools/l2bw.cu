// L2 / HBM read-bandwidth probe (design input for the tensor-core path's operand streaming).
// Blocks rotate over the buffer every iteration so that no SM re-reads its own L1-resident slice.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bw tools/l2bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void read_kernel(const uint4* __restrict__ p, size_t n, int iters, unsigned* sink) {
    unsigned acc = 0;
    const size_t per_block = n / gridDim.x;   // contiguous slice per block
    for (int it = 0; it < iters; ++it) {
        const size_t blk = (blockIdx.x + (size_t)it * 61) % gridDim.x;
        const uint4* base = p + blk * per_block;
        for (size_t i = threadIdx.x; i + 3 * blockDim.x < per_block; i += 4 * blockDim.x) {
            uint4 a, b, c, d;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(base + i));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(base + i + blockDim.x));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(base + i + 2 * blockDim.x));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(base + i + 3 * blockDim.x));
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    unsigned* sink;
    cudaMalloc(&sink, 4);
    size_t sizes_mb[] = {8, 16, 32, 64, 96, 128, 256, 2048};
    for (size_t mb : sizes_mb) {
        size_t bytes = mb << 20;
        uint4* buf;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 1, bytes);
        size_t n = bytes / 16;
        int iters = (int)((size_t)(16ull << 30) / bytes);
        if (iters < 2) iters = 2;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        read_kernel<<<148 * 4, 512>>>(buf, n, 2, sink);
        cudaEventRecord(a);
        read_kernel<<<148 * 4, 512>>>(buf, n, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("working set %5zu MB: %8.1f GB/s (ld.global.cg, block-rotated)\n", mb, (double)bytes * iters / (ms * 1e-3) / 1e9);
        cudaFree(buf);
    }
    return 0;
}
