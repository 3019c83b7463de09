// Launch wrappers of the tensor-core layer kernels.  The kernel templates are instantiated for K = 1..10 jet
// components in three separate translation units (tc_layers_a/b/c.cu) so that the build parallelises.
#pragma once
#include <cstring>

#include "tc_kernels.cuh"
#include "tc_path.h"

namespace stpde {

int tc_fail(int code, const char* msg);
int tc_launch_layer(int kc, const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);
int tc_launch_layer_pair(int kc, const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st);

// The K = 1..10 instantiations of every kernel family are split over two translation units (K <= 5 / K >= 6) so that
// the build parallelises: *_lo.cu / *_hi.cu include the family's .inc with STPDE_KC_HALF 0 / 1.
#define STPDE_TC_DISPATCH_KC_LO(kc, CALL)              \
    switch (kc) {                                      \
        case 1: { constexpr int KC = 1; CALL; } break; \
        case 2: { constexpr int KC = 2; CALL; } break; \
        case 3: { constexpr int KC = 3; CALL; } break; \
        case 4: { constexpr int KC = 4; CALL; } break; \
        default: { constexpr int KC = 5; CALL; } break; \
    }
#ifdef STPDE_ONLY_RB2
#define STPDE_TC_DISPATCH_KC_HI(kc, CALL) { constexpr int KC = 6; CALL; }
#else
#define STPDE_TC_DISPATCH_KC_HI(kc, CALL)              \
    switch (kc) {                                      \
        case 6: { constexpr int KC = 6; CALL; } break; \
        case 7: { constexpr int KC = 7; CALL; } break; \
        case 8: { constexpr int KC = 8; CALL; } break; \
        case 9: { constexpr int KC = 9; CALL; } break; \
        default: { constexpr int KC = 10; CALL; } break; \
    }
#endif
#ifdef STPDE_ONLY_RB2   // kernel experiments (tools/build_variant.sh): instantiate K = 6 only, builds in a fraction of the time
#define STPDE_TC_DISPATCH_KC(kc, CALL) { constexpr int KC = 6; CALL; }
#else
#define STPDE_TC_DISPATCH_KC(kc, CALL)                 \
    switch (kc) {                                      \
        case 1: { constexpr int KC = 1; CALL; } break; \
        case 2: { constexpr int KC = 2; CALL; } break; \
        case 3: { constexpr int KC = 3; CALL; } break; \
        case 4: { constexpr int KC = 4; CALL; } break; \
        case 5: { constexpr int KC = 5; CALL; } break; \
        case 6: { constexpr int KC = 6; CALL; } break; \
        case 7: { constexpr int KC = 7; CALL; } break; \
        case 8: { constexpr int KC = 8; CALL; } break; \
        case 9: { constexpr int KC = 9; CALL; } break; \
        default: { constexpr int KC = 10; CALL; } break; \
    }
#endif

// Rayleigh-Benard jet set [value | d0, d1, d2 | d11, d22] -> kernels specialised at compile time (tc::kSpecRb2)
static inline bool spec_is_rb2(const JetSpec& s) {
    return s.kc == 6 && s.n_first == 3 && s.n_second == 2 && s.pa[4] == 2 && s.pb[4] == 2 && s.pa[5] == 3 && s.pb[5] == 3;
}

// Fills a.out_map for the output planes of a forward layer launch (tc_path.cu).
int tc_encode_out_maps(tc::LayerArgs& a, int kc, void* plane0, void* plane1, bool f32, bool reverse = false);

// Launch of a tensor-core kernel with programmatic stream serialization (STPDE_PDL=0: plain launch): the kernel's
// prologue runs while the previous kernel in the stream drains; griddep_wait() in the kernel orders the data.
static inline cudaError_t tc_launch_ex(const void* func, unsigned grid, unsigned block, size_t smem, cudaStream_t st, void** args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_env().pdl ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, func, args);
}

#ifdef STPDE_TC_LAUNCH_IMPL
#define STPDE_TC_LAUNCH(KERNEL, GRID, SMEM, ST, M0, M1, M2, M3, SPEC_, ARGS_)                                           \
    do {                                                                                                                 \
        void* kargs_[] = {(void*)&(M0), (void*)&(M1), (void*)&(M2), (void*)&(M3), (void*)&(SPEC_), (void*)&(ARGS_)};     \
        if (tc_launch_ex((const void*)(KERNEL), (unsigned)(GRID), tc::kThreads, (SMEM), (ST), kargs_) != cudaSuccess)    \
            return tc_fail(STPDE_ECUDA, "tensor-core kernel launch failed");                                             \
    } while (0)
template <int KC, int SPEC = 0>
static int launch_layer(const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a,
                        cudaStream_t st) {
    constexpr int NR = tc::rows_per_tile(KC);
    constexpr int N = KC * NR;
    const size_t smem = (size_t)tc::kStages * (2 * tc::kTileF * tc::kBlockK * 2 + 2 * N * tc::kBlockK * 2) +
                        tc::epi_staging_total<KC, tc::kModeFwd, true>() + 1024 + 256;
    static DeviceOnce configured;               // function attributes are per device
    if (configured.first_use()) {
        if (cudaFuncSetAttribute(tc::tc_layer_kernel<KC, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "cudaFuncSetAttribute(tc_layer_kernel) failed");
        configured.mark();
    }
    const int rpt = NR * (a.pack > 1 ? a.pack : 1);            // packed narrow layers: pack row groups per tile
    const int n_tiles = a.pack > 1 ? (a.rows + rpt - 1) / rpt : ((a.n_store + tc::kTileF - 1) / tc::kTileF) * ((a.rows + NR - 1) / NR);
    const int grid = n_tiles < tc.num_sms ? n_tiles : tc.num_sms;
    STPDE_TC_LAUNCH((tc::tc_layer_kernel<KC, SPEC>), grid, smem, st, L.w_hi, L.w_lo, L.a_hi, L.a_lo, spec, a);
    return STPDE_OK;
}

template <int KC, int SPEC = 0>
static int launch_layer_pair(const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a,
                             cudaStream_t st) {
    constexpr int NR = tc::rows_per_tile(KC);
    const size_t smem = (size_t)tc::kPairSmemBudget + 1024 + 512;
    static DeviceOnce configured;
    if (configured.first_use()) {
        if (cudaFuncSetAttribute(tc::tc_layer_pair_kernel<KC, tc::kModeFwd, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "cudaFuncSetAttribute(tc_layer_pair_kernel) failed");
        configured.mark();
    }
    const int n_tiles = ((a.n_store + 2 * tc::kTileF - 1) / (2 * tc::kTileF)) * ((a.rows + NR - 1) / NR);
    const int max_pairs = tc.num_sms / 2;
    const int n_pairs = n_tiles < max_pairs ? n_tiles : max_pairs;
    STPDE_TC_LAUNCH((tc::tc_layer_pair_kernel<KC, tc::kModeFwd, SPEC>), 2 * n_pairs, smem, st, L.w_hi, L.w_lo, L.a_hi, L.a_lo, spec, a);
    return STPDE_OK;
}

// CTA-pair kernel in one of the training modes (tc_kernels.cuh kMode*): the maps are passed explicitly because the
// reverse sweep runs the same kernel on W^T / zbar planes.  Always the pair kernel: TMA zero-fills the feature
// rows beyond the layer width.
template <int KC, int MODE, int SPEC = 0>
static int launch_layer_pair_mode(int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                                  const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    constexpr int NR = tc::rows_per_tile(KC);
    const size_t smem = (size_t)tc::kPairSmemBudget + 1024 + 512;
    static DeviceOnce configured;
    if (configured.first_use()) {
        if (cudaFuncSetAttribute(tc::tc_layer_pair_kernel<KC, MODE, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "cudaFuncSetAttribute(tc_layer_pair_kernel, training mode) failed");
        configured.mark();
    }
    const int n_tiles = ((a.n_store + 2 * tc::kTileF - 1) / (2 * tc::kTileF)) * ((a.rows + NR - 1) / NR);
    const int max_pairs = num_sms / 2;
    const int n_pairs = n_tiles < max_pairs ? n_tiles : max_pairs;
    STPDE_TC_LAUNCH((tc::tc_layer_pair_kernel<KC, MODE, SPEC>), 2 * n_pairs, smem, st, w_hi, w_lo, a_hi, a_lo, spec, a);
    return STPDE_OK;
}

// Single-CTA kernel (M = 128 features per CTA, one CTA per SM) in a training mode: used for the narrow layers, where
// a 256-feature CTA-pair tile would leave one CTA (and most epilogue warps) idle.
template <int KC, int MODE, int SPEC = 0>
static int launch_layer_mode(int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                             const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    constexpr int NR = tc::rows_per_tile(KC);
    constexpr int N = KC * NR;
    const size_t smem = (size_t)tc::kStages * (2 * tc::kTileF * tc::kBlockK * 2 + 2 * N * tc::kBlockK * 2) +
                        tc::epi_staging_total<KC, MODE, true>() + 1024 + 256;
    static DeviceOnce configured;
    if (configured.first_use()) {
        if (cudaFuncSetAttribute(tc::tc_layer_kernel<KC, SPEC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "cudaFuncSetAttribute(tc_layer_kernel, training mode) failed");
        configured.mark();
    }
    const int rpt = NR * (a.pack > 1 ? a.pack : 1);
    const int n_tiles = rpt > NR ? (a.rows + rpt - 1) / rpt : ((a.n_store + tc::kTileF - 1) / tc::kTileF) * ((a.rows + NR - 1) / NR);
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    STPDE_TC_LAUNCH((tc::tc_layer_kernel<KC, SPEC, MODE>), grid, smem, st, w_hi, w_lo, a_hi, a_lo, spec, a);
    return STPDE_OK;
}

#endif  // STPDE_TC_LAUNCH_IMPL

}  // namespace stpde
