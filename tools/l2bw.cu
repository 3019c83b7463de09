// L2 / HBM read-bandwidth probe (design input for the tensor-core path's operand streaming).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bw tools/l2bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void read_kernel(const uint4* __restrict__ p, size_t n, int iters, unsigned* sink) {
    unsigned acc = 0;
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = tid; i < n; i += stride * 4) {
            uint4 a = p[i];
            uint4 b = (i + stride < n) ? p[i + stride] : a;
            uint4 c = (i + 2 * stride < n) ? p[i + 2 * stride] : a;
            uint4 d = (i + 3 * stride < n) ? p[i + 3 * stride] : a;
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
        read_kernel<<<148 * 8, 256>>>(buf, n, 2, sink);
        cudaEventRecord(a);
        read_kernel<<<148 * 8, 256>>>(buf, n, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("working set %5zu MB: %8.1f GB/s\n", mb, (double)bytes * iters / (ms * 1e-3) / 1e9);
        cudaFree(buf);
    }
    return 0;
}
