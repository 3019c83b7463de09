// tcgen05 tensor-core path (fp16 hi/lo split operands, fp32 accumulation in TMEM).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "kernels.h"

namespace stpde {

struct TcContext {
    void* impl[64];
};

size_t tc_fixed_bytes(const stpde_desc_t* d, int n_layers, const int* widths, const int* np);
size_t tc_per_point_bytes(const stpde_desc_t* d, int kc, int ncorner, int max_even, int max_odd);
int tc_prepare(TcContext& tc, const stpde_desc_t* d, int n_layers, const int* widths, const int* np, const int* kh,
               const int* in_features, const float* const* W, char* fixed_ws, char* chunk_ws, size_t chunk_bytes,
               const JetSpec& spec, int pc, int ncorner, int* status, cudaStream_t st);
int tc_run_chunk(TcContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                 const float* Wx0, const float* Vb, int ncat, const int* cat_off, const float* const* unused,
                 char* ws, const size_t* off_wx, float* act_last, cudaStream_t st);
const char* tc_last_error();

}  // namespace stpde
