// Host side of the reverse-mode sweep on the tensor cores: per chunk of points the forward is recomputed with all
// activations / pre-activations kept (chunk-level checkpointing), then every hidden layer runs
//   wgrad_l : gW_l[:, :kh] += zbar_l^T . a_{l-1}         (tc_wgrad_pair_kernel, MN-major operands, split-K)
//   dgrad_l : abar_{l-1} = W_l^T . zbar_l + reverse jet activation -> zbar_{l-1}   (tc_layer_pair_kernel MODE 2 / 3)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include "kernels.h"
#include "tc_kernels.cuh"

namespace stpde {

// TcBwdContext::z_half: the saved pre-activations z_l are fp16 planes (the forward that wrote them ran in the single-pass
// mode); set by tc_bwd_prepare from its z_half argument, read by the forward-save, dgrad and blend_backward launches.
struct TcBwdLayer {               // hidden layer l = 1 .. n_layers-2
    int n_feat, kh, ldz, ld_in, last, pack, pack_t;   // row groups per tile: forward (tc_layer_pack) / dgrad (block-diagonal W^T)
    CUtensorMap w_hi, w_lo;       // forward: W_l planes [np256][ld_in], box 64 x 128
    CUtensorMap fa_hi, fa_lo;     // forward: a_{l-1} planes (ld_in, rows, kc), box 64 x 8 x kc
    CUtensorMap wt_hi, wt_lo;     // dgrad:   W_l^T planes [fp256][ldz], box 64 x 128
    CUtensorMap zb_hi, zb_lo;     // dgrad:   zbar_l planes (ldz, rows, kc), box 64 x 8 x kc
    CUtensorMap ga_hi, ga_lo;     // wgrad:   a_{l-1} planes, box 64 features x 64 rows
    CUtensorMap gb_hi, gb_lo;     // wgrad:   zbar_l planes, box 64 features x 64 rows
    __half *w_hi_ptr, *w_lo_ptr, *wt_hi_ptr, *wt_lo_ptr;
    __half *a_in[2];              // a_{l-1} hi / lo
    __half *a_out[2];             // a_l hi / lo (unused for the last hidden layer: fp32 act_last instead)
    float* z;                     // z_l [kc][rows][ldz]
    __half* zb[2];                // zbar_l hi / lo
};

struct TcBwdContext {
    int n_layers, kc, rows, passes, num_sms, use_pair_wide;
    int ld0, n0, z_half;
    TcBwdLayer layer[kMaxLayers];
    float* wscale;
    unsigned* absmax;
    int* status;
};

size_t tc_bwd_fixed_bytes(int n_layers, const int* widths);
size_t tc_bwd_per_point_bytes(int n_layers, const int* widths, int kc, int ncorner);

int tc_bwd_prepare(TcBwdContext& tc, int precision, int n_layers, const int* widths, const int* in_features,
                   const float* const* W, char* fixed_ws, char* chunk_ws, size_t chunk_bytes, int kc, int rows,
                   int* status, bool split_weights, bool z_half, cudaStream_t st);

// forward recompute of one chunk: layer 0 + hidden layers, all operand planes and pre-activations kept;
// act_last = fp32 activations of the last hidden layer [kc][rows][np_last]
int tc_bwd_forward_chunk(TcBwdContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                         const float* Vb, int ncat, const int* cat_off, const float* const* Wx, float* act_last,
                         int np_last, const struct TcFinal* fused_final, cudaStream_t st);
// the training forward may fuse the final layer + blend into the last hidden layer's epilogue (it still writes act_last)
bool tc_bwd_can_fuse_final(const TcBwdContext& tc, const JetSpec& spec, int dim, int n_out);

// reverse sweep of one chunk; zbar of the last hidden layer must already be in layer[n-2].zb (blend_backward).
// gW[l] : gradient of layer l's weight [widths[l]][in_features[l]] (scaled by S), g_vb [nvert][ncat]
int tc_bwd_backward_chunk(TcBwdContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                          const float* Vb, int ncat, const int* cat_off, const int* in_features, const float* const* Wx,
                          float* const* gW, float* g_vb, float* g_beta, cudaStream_t st);

// launch wrappers instantiated in tc_bwd_a/b/c.cu (one translation unit per kernel mode)
int tc_launch_pair_save(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                        const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);
int tc_launch_pair_bwd(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                       const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);
int tc_launch_pair_bwd0(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                        const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);

// single-CTA variants (M = 128) for layers narrower than a CTA-pair tile
int tc_launch_single_save(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                          const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);
int tc_launch_single_bwd(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                         const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);
int tc_launch_single_bwd0(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                          const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);

// shared with tc_path.cu
int tc_make_map_2d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1);
int tc_make_map_3d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2);
bool tc_encode_available();
void tc_launch_split_weights(const float* W, int N, int in_features, int kh, int np, int kp, int pack, unsigned* absmax,
                             float* wscale, __half* hi, __half* lo, cudaStream_t st);
void tc_launch_layer0_planes(int kc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb, int N, int ld,
                             const float* Wx, const float* Vb, int ncat, int three, __half* out_hi, __half* out_lo,
                             int* status, cudaStream_t st);

}  // namespace stpde
