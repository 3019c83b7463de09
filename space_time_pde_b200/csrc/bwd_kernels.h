// Internal declarations of the reverse-mode CUDA-core kernels (bwd_kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace stpde {

struct BlendBwdArgs {
    int dim, rows, O, Kp, n_feat, ld_out, ldz, act, ncat, cat_off, three;
    int z_half;             // 1: z_in holds fp16 (single-pass training), 0: fp32
    int acc_copies;         // set by the launcher: private per-warp accumulator copies in shared memory (8) or 1
    float beta;
    int64_t total_pts, p0;
    ChunkBuffers cb;
    const float* gy;        // [b*p][O]
    const float* gjets;     // [n_jet][b*p][O] or nullptr
    const float* scale;     // device {S, 1/S}
    const float* Wlast;     // packed [O][Kp]
    const float* act_last;  // [KC][rows][Kp] fp32 activations of the last hidden layer
    const float* z_in;      // [KC][rows][ldz] its pre-activations
    __half* out_hi;         // [KC][rows][ld_out] zbar planes of the last hidden layer
    __half* out_lo;
    float* g_vb;            // [nvert][ncat]
    float* g_wx;            // &gW[n-2][0][kh], row stride g_wx_ld
    int g_wx_ld;
    float* g_wlast;         // [O][n_feat]
    float* g_blast;         // [O]
    float* g_beta;          // adjoint of the Swish beta (may be null)
    int* status;
};

struct VertexBwdArgs {
    int n_layers, ncat;
    int cat_off[kMaxLayers], in_features[kMaxLayers], kh[kMaxLayers];
    const float* W[kMaxLayers];
    float* gW[kMaxLayers];
    float* gB[kMaxLayers];
    const float* grid;
    const float* g_vb;
    const float* scale;
};

void launch_grad_scale(const JetSpec& spec, const GridGeom& g, const float* gy, const float* gjets, int64_t plane_elems,
                       unsigned* maxes, float* scale, int target_exp, cudaStream_t st);
// up to 2 * kMaxLayers + 2 caller-owned gradient buffers handled by ONE launch
struct BufferList {
    float* ptr[2 * kMaxLayers + 2];
    int64_t n[2 * kMaxLayers + 2];
    int count;
    void add(float* p, int64_t len) {
        if (p && len > 0) { ptr[count] = p; n[count] = len; ++count; }
    }
};
void launch_zero_buffers(const BufferList& bl, cudaStream_t st);
void launch_scale_buffers(const BufferList& bl, const float* scale, cudaStream_t st);   // *= scale[1]
void launch_split_weights_t(const float* W, int N, int in_features, int kh, int fp, int ldz, int pack, const unsigned* absmax,
                            __half* hi, __half* lo, cudaStream_t st);
int launch_blend_backward(const JetSpec& spec, const BlendBwdArgs& a, cudaStream_t st);
void launch_vertex_backward(const GridGeom& g, int nvert_total, const VertexBwdArgs& a, float* ggrid, cudaStream_t st);

}  // namespace stpde
