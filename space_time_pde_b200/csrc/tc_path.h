// Host side of the tcgen05 tensor-core path (fp16 hi/lo split operands, fp32 accumulation in TMEM).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "kernels.h"

namespace stpde {

struct TcEnv {
    int use_pair;
    uint32_t wait_ns;
    int fuse_final;      // STPDE_FUSE_FINAL=0: keep the separate final_blend kernel
    int l0_tma;          // STPDE_L0_TMA=0: layer 0 writes its planes with st.global instead of TMA stores
    int l0_rows_per_warp;     // STPDE_L0_RPW: rows one warp of the layer-0 kernel walks (block = 8 warps x this many rows)
    int wgrad_tile_fastest;   // STPDE_WGRAD_ORDER=0: round-1 unit order of the weight-gradient kernel (slices of a tile adjacent)
    int pack_narrow;          // STPDE_PACK=0: narrow layers (<= 64 features) keep one row group per 128-lane tile
    int z_half;               // STPDE_Z_HALF=0: the single-pass training mode keeps fp32 pre-activation planes
    int pdl;                  // STPDE_PDL=0: tensor-core kernels are launched without programmatic stream serialization
};

// Final linear layer + multilinear blend fused into the epilogue of the last hidden layer (inference, d = 3, the
// Rayleigh-Benard jet set, <= 4 outputs, last hidden layer narrower than a CTA-pair tile): the launch writes y / jets.
struct TcFinal {
    const float* w_last;   // [O][ldw] padded last-layer weights
    const float* b_last;   // [O]
    int n_out, ldw;
    float* y;
    float* jets;
    int64_t p0, total_pts;
};
bool tc_can_fuse_final(const struct TcContext& tc, const JetSpec& spec, int dim, int n_out);
const TcEnv& tc_env();   // environment switches, read once per process
// Row groups per tile of the single-CTA forward kernel (tc::LayerArgs::pack) for a hidden layer of n_feat features:
// the last hidden layer stores round_up(n_feat, 16) features (4 groups when that is <= 32), the others 64-feature
// multiples (2 groups when n_feat <= 64).  The weight operand of a packed layer is block diagonal [128][pack * kp].
int tc_layer_pack(int n_feat, bool last);

struct TcLayerPlan {
    int n_feat, np128, kp_in, ld_out, n_store, last, cat_off, pack;
    CUtensorMap w_hi, w_lo, a_hi, a_lo;
    __half *w_hi_ptr, *w_lo_ptr;
};

struct TcContext {
    int n_layers, kc, rows, passes, num_sms, use_pair;
    TcLayerPlan layer[kMaxLayers];     // hidden layers 1..n_layers-2
    __half* act[2][2];                 // [buffer parity][hi/lo] planes [KC][rows][ld]
    int ld0, n0;                       // row stride / true width of layer 0's output planes
    float* wscale;                     // device [kMaxLayers]
    unsigned* absmax;                  // device [kMaxLayers]
    int* status;
};

// bytes of the call-invariant region (split weights, scales) and of the per-point activation planes
size_t tc_fixed_bytes(int n_layers, const int* widths);
size_t tc_per_point_bytes(int n_layers, const int* widths, int kc, int ncorner);

// Once per call: split the hidden-layer weights into scaled fp16 hi/lo planes, build the TMA maps.
int tc_prepare(TcContext& tc, int precision, int n_layers, const int* widths, const int* in_features,
               const float* const* W, char* fixed_ws, char* chunk_ws, size_t chunk_bytes, int kc, int rows,
               int* status, bool split_weights, cudaStream_t st);

// Per chunk: layer 0 (closed form) -> hidden layers on the tensor cores; the last hidden layer's
// activations are written as fp32 [KC][rows][np_last] into act_last for final_blend.
int tc_run_chunk(TcContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                 const float* Vb, int ncat, const int* cat_off, char* ws, const size_t* off_wx, float* act_last,
                 int np_last, const TcFinal* fused_final, cudaStream_t st);

const char* tc_last_error();

}  // namespace stpde
