// Host orchestration + small helper kernels of the tensor-core path.
#include "tc_path.h"
#include "tc_bwd.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "profile.h"
#include "tc_kernels.cuh"
#include "tc_launch.cuh"

namespace stpde {

static thread_local char g_tc_err[256] = "";
const char* tc_last_error() { return g_tc_err; }
int tc_fail(int code, const char* msg) { snprintf(g_tc_err, sizeof(g_tc_err), "%s", msg); return code; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ W, int N, int in_features, int kh, unsigned* __restrict__ out) {
    float m = 0.f;
    const int64_t total = (int64_t)N * kh;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(e / kh), k = (int)(e % kh);
        m = fmaxf(m, fabsf(W[(int64_t)n * in_features + k]));
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}

// Whi/Wlo [np128][pack * kp] = fp16 split of W[:, :kh] * 2^sw with max|W| * 2^sw in [2^13, 2^14).
// pack > 1 (narrow layers, tc_layer_pack): block diagonal - copy q of W sits in rows q * 128/pack ... and columns
// q * kp ..., everything else is zero.
__global__ void split_weights_kernel(const float* __restrict__ W, int N, int in_features, int kh, int np128, int kp, int pack,
                                     const unsigned* __restrict__ absmax, float* __restrict__ wscale,
                                     __half* __restrict__ hi, __half* __restrict__ lo) {
    const float amax = __uint_as_float(*absmax);
    int e2 = 0;
    if (amax > 0.f) frexpf(amax, &e2);           // amax = m * 2^e2, m in [0.5, 1)
    const int sw = amax > 0.f ? 14 - e2 : 0;
    const float up = ldexpf(1.f, sw);
    if (blockIdx.x == 0 && threadIdx.x == 0) *wscale = ldexpf(1.f, -(sw + tc::kActScaleLog2));
    const int ldk = pack * kp, lanes = pack > 1 ? 128 / pack : np128;
    const int64_t total = (int64_t)np128 * ldk;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / ldk), col = (int)(e % ldk);
        const int q = col / kp, k = col - q * kp, n = row - q * lanes;
        float x = (n >= 0 && n < lanes && n < N && k < kh) ? W[(int64_t)n * in_features + k] * up : 0.f;
        __half h = __float2half_rn(x);
        hi[e] = h;
        lo[e] = __float2half_rn(x - __half2float(h));
    }
}

// layer 0 jets written as scaled fp16 hi/lo planes [KC][rows][ld] (pad columns n >= N are zero).
// One thread = 4 consecutive features of one row per iteration (8-byte stores, 256 B per warp and plane); the
// per-feature constants live in registers, the next row's operands are prefetched while the current row is
// evaluated (the kernel is latency-bound on the Vb gather otherwise).
// THREE (lo plane written) is a template parameter: the single-pass mode used to compute the lo halves and drop them
// (30 of its 207 instructions per warp-row; the kernel is issue-bound in that mode, ncu r02).
// TMA_OUT: the 24 (48) halves a thread produces per row go to a per-warp smem staging buffer [plane][component][128
// features] with 8-byte st.shared and leave through one cp.async.bulk.tensor store per plane and row (box 128 features x
// 1 row x K) instead of 6 (12) st.global with 64-bit address arithmetic each.
template <int KC, bool THREE, bool TMA_OUT>
__global__ void __launch_bounds__(256, 3) layer0_jets_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                                const __grid_constant__ CUtensorMap map_lo, JetSpec spec, int dim,
                                                                int act, float beta, int rows, int rpw, int N, int ld,
                                                                const int* __restrict__ vtx, const float* __restrict__ xrel,
                                                                const float* __restrict__ Wx, const float* __restrict__ Vb, int ncat,
                                                                __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                                int* __restrict__ status) {
    constexpr int F = 4;
    const int fg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 32 + fg) * F;
    __shared__ __align__(128) uint8_t staging[TMA_OUT ? 8 : 1][TMA_OUT ? (THREE ? 2 : 1) * KC * 256 : 16];
    if (!TMA_OUT && n0 >= ld) return;
    const uint32_t stg = tc::smem_u32(&staging[TMA_OUT ? rl : 0][0]);
    float wx[F][kMaxDim];
#pragma unroll
    for (int e = 0; e < F; ++e)
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) wx[e][k] = (k < dim && n0 + e < N) ? __ldg(Wx + (n0 + e) * dim + k) : 0.f;
    const float act_scale = (float)(1 << tc::kActScaleLog2);
    // a_c = sigma^(order_c)(z0) * coef[c]: coef = 2^4 (value), W0x[:,dir] * 2^4, W0x[:,a] * W0x[:,b] * 2^4; 0 for pad features
    float coef[KC][F];
    float cmax = 0.f, smax = 0.f;
#pragma unroll
    for (int e = 0; e < F; ++e) {
        const float fm = (n0 + e < N) ? act_scale : 0.f;
        coef[0][e] = fm;
#pragma unroll
        for (int c = 1; c < KC; ++c) {
            float wa = 1.f, wb = 1.f;
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) {
                if (spec.kind[c] == 1 && k == spec.dir[c]) wa = wx[e][k];
                if (spec.kind[c] == 2 && k == spec.dir[spec.pa[c]]) wa = wx[e][k];
                if (spec.kind[c] == 2 && k == spec.dir[spec.pb[c]]) wb = wx[e][k];
            }
            coef[c][e] = wa * wb * fm;
        }
#pragma unroll
        for (int c = 0; c < KC; ++c) cmax = fmaxf(cmax, fabsf(coef[c][e]));
    }
    const bool vec_ok = (n0 + F <= N) && (ncat % 4 == 0);
    const int n_first = spec.n_first;
    auto load_row = [&](int r, float* xr, float* vb) {
        const int rc = min(r, rows - 1);
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) xr[k] = __ldg(xrel + (int64_t)k * rows + rc);     // planes >= dim are zero
        const float* vrow = Vb + (int64_t)__ldg(vtx + rc) * ncat + n0;
        if (vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(vrow));
            vb[0] = a.x; vb[1] = a.y; vb[2] = a.z; vb[3] = a.w;
        } else {
#pragma unroll
            for (int e = 0; e < F; ++e) vb[e] = (n0 + e < N) ? __ldg(vrow + e) : 0.f;
        }
    };
    float xr[kMaxDim], vb[F];
    const int rbase = blockIdx.y * (8 * rpw) + rl;
    if (rbase < rows) load_row(rbase, xr, vb);
    dispatch_act(act, [&](auto act_c) {
    constexpr int kAct = decltype(act_c)::value;
#pragma unroll 1
    for (int j = 0; j < rpw; ++j) {
        const int r = rbase + 8 * j;
        if (r >= rows) break;
        float xr_n[kMaxDim], vb_n[F];
        load_row(r + 8, xr_n, vb_n);                      // prefetch (clamped to a valid row)
        float s0[F], s1[F], s2[F];
#pragma unroll
        for (int e = 0; e < F; ++e) {
            float z = vb[e];
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) z = fmaf(wx[e][k], xr[k], z);
            act_jet_fast(kAct, beta, z, s0[e], s1[e], s2[e]);
            smax = fmaxf(smax, fmaxf(fabsf(s0[e]), fmaxf(fabsf(s1[e]), fabsf(s2[e]))));
        }
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            float x[F];
            if (c == 0) {
#pragma unroll
                for (int e = 0; e < F; ++e) x[e] = s0[e] * coef[c][e];
            } else if (c <= n_first) {                    // warp-uniform branch instead of per-element selects
#pragma unroll
                for (int e = 0; e < F; ++e) x[e] = s1[e] * coef[c][e];
            } else {
#pragma unroll
                for (int e = 0; e < F; ++e) x[e] = s2[e] * coef[c][e];
            }
            uint32_t ph[F / 2], pl[F / 2];
#pragma unroll
            for (int e = 0; e < F; e += 2) {
                const __half2 h = __floats2half2_rn(x[e], x[e + 1]);
                ph[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                if constexpr (THREE) {
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(x[e] - hf.x, x[e + 1] - hf.y);
                    pl[e >> 1] = *reinterpret_cast<const uint32_t*>(&l);
                }
            }
            if constexpr (TMA_OUT) {
                if (c == 0) {                                  // the previous row's stores have read the buffer
                    if (fg == 0) tc::bulk_wait_read0();
                    __syncwarp();
                }
                tc::sts_v2(stg + c * 256 + fg * 8, ph[0], ph[1]);
                if constexpr (THREE) tc::sts_v2(stg + (KC + c) * 256 + fg * 8, pl[0], pl[1]);
            } else {
                const int64_t off = ((int64_t)c * rows + r) * ld + n0;
                *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(ph[0], ph[1]);
                if constexpr (THREE) *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(pl[0], pl[1]);
            }
        }
        if constexpr (TMA_OUT) {
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (fg == 0) {
                tc::tma_store_3d(&map_hi, stg, blockIdx.x * 128, r, 0);
                if constexpr (THREE) tc::tma_store_3d(&map_lo, stg + KC * 256, blockIdx.x * 128, r, 0);
                tc::bulk_commit();
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) xr[k] = xr_n[k];
#pragma unroll
        for (int e = 0; e < F; ++e) vb[e] = vb_n[e];
    }
    });
    if (TMA_OUT && fg == 0) tc::bulk_wait0();
    if (!(smax * cmax < 65000.f)) atomicOr(status, kStatusRange);
}

// ---------------------------------------------------------------------------------------------
// sizes
// ---------------------------------------------------------------------------------------------
size_t tc_fixed_bytes(int n_layers, const int* widths) {
    size_t off = 1024;  // wscale + absmax
    for (int l = 1; l <= n_layers - 2; ++l) {
        size_t plane = (size_t)round_up(widths[l], 128) * round_up(widths[l - 1], 64) * sizeof(__half) *
                       tc_layer_pack(widths[l], l == n_layers - 2);
        off += 2 * align_up(plane, 1024);
    }
    return off;
}

static void plane_lds(int n_layers, const int* widths, int& max_even, int& max_odd) {
    max_even = max_odd = 64;
    for (int l = 0; l <= n_layers - 3; ++l) {     // outputs that feed a tensor-core layer
        int ld = round_up(widths[l], 64);
        if (l % 2 == 0) max_even = ld > max_even ? ld : max_even;
        else max_odd = ld > max_odd ? ld : max_odd;
    }
}

size_t tc_per_point_bytes(int n_layers, const int* widths, int kc, int ncorner) {
    int me, mo;
    plane_lds(n_layers, widths, me, mo);
    return (size_t)2 * kc * ncorner * ((size_t)me + mo) * sizeof(__half);
}

// ---------------------------------------------------------------------------------------------
// TMA maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_map_2d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {d0 * sizeof(__half)};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_map_3d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                       uint32_t b2) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * sizeof(__half), d0 * d1 * sizeof(__half)};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Process-wide switches read from the environment ONCE (getenv is not free and not thread-safe against setenv):
// STPDE_TC_PAIR=0 disables the CTA-pair kernels, STPDE_WAIT_NS overrides the mbarrier suspend-time hint.
const TcEnv& tc_env() {
    static const TcEnv env = [] {
        TcEnv e;
        const char* p = getenv("STPDE_TC_PAIR");
        e.use_pair = p ? atoi(p) : 1;
        const char* w = getenv("STPDE_WAIT_NS");
        e.wait_ns = w ? (uint32_t)atoll(w) : 0x989680u;
        const char* f = getenv("STPDE_FUSE_FINAL");
        e.fuse_final = f ? atoi(f) : 1;
        const char* l0 = getenv("STPDE_L0_TMA");
        e.l0_tma = l0 ? atoi(l0) : 1;
        const char* rpw = getenv("STPDE_L0_RPW");
        e.l0_rows_per_warp = rpw && atoi(rpw) > 0 ? atoi(rpw) : 32;   // measured r02/s28: 8 -> 32 rows is 9-12 % faster
        const char* wo = getenv("STPDE_WGRAD_ORDER");
        e.wgrad_tile_fastest = wo ? atoi(wo) : 1;
        const char* pk = getenv("STPDE_PACK");
        e.pack_narrow = pk ? atoi(pk) : 1;
        const char* zh = getenv("STPDE_Z_HALF");
        e.z_half = zh ? atoi(zh) : 1;
        const char* pdl = getenv("STPDE_PDL");
        e.pdl = pdl ? atoi(pdl) : 1;
        return e;
    }();
    return env;
}

int tc_layer_pack(int n_feat, bool last) {
    if (!tc_env().pack_narrow) return 1;
    if (last) return round_up(n_feat, 16) <= 32 ? 4 : 1;
    return n_feat <= 64 ? 2 : 1;
}

// TMA store maps of a layer's output planes [kc][rows][ld_out]: box = 32 features x stage_rows(kc) rows x kc components,
// no swizzle (the epilogue warps write their staging buffers in exactly that order).
int tc_encode_out_maps(tc::LayerArgs& a, int kc, void* plane0, void* plane1, bool f32, bool reverse) {
    cuuint64_t dims[3] = {(cuuint64_t)a.ld_out, (cuuint64_t)a.rows, (cuuint64_t)kc};
    const size_t es = f32 ? sizeof(float) : sizeof(__half);
    cuuint64_t strides[2] = {(cuuint64_t)a.ld_out * es, (cuuint64_t)a.ld_out * a.rows * es};
    // rows per store: stage_rows(kc), twice that when only one fp16 plane is written (same staging bytes)
    const int srb = reverse ? tc::stage_rows_bwd(kc) : tc::stage_rows(kc);
    const int box_rows = (!f32 && !plane1 && srb < 8) ? 2 * srb : srb;
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)kc};
    cuuint32_t estr[3] = {1, 1, 1};
    void* planes[2] = {plane0, plane1};
    for (int i = 0; i < 2; ++i) {
        if (!planes[i]) { a.out_map[i] = a.out_map[0]; continue; }
        CUresult r = encode_fn()(&a.out_map[i], f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                                 planes[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return tc_fail(STPDE_ECUDA, "cuTensorMapEncodeTiled failed (output planes)");
    }
    return STPDE_OK;
}

// helpers shared with the reverse-mode path (tc_bwd.cu)
int tc_make_map_2d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
    return make_map_2d(m, base, d0, d1, b0, b1);
}
int tc_make_map_3d(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2) {
    return make_map_3d(m, base, d0, d1, d2, b0, b1, b2);
}
bool tc_encode_available() { return encode_fn() != nullptr; }
void tc_launch_split_weights(const float* W, int N, int in_features, int kh, int np, int kp, int pack, unsigned* absmax,
                             float* wscale, __half* hi, __half* lo, cudaStream_t st) {
    absmax_kernel<<<148, 256, 0, st>>>(W, N, in_features, kh, absmax);
    split_weights_kernel<<<148 * 4, 256, 0, st>>>(W, N, in_features, kh, np, kp, pack, absmax, wscale, hi, lo);
}

int tc_prepare(TcContext& tc, int precision, int n_layers, const int* widths, const int* in_features,
               const float* const* W, char* fixed_ws, char* chunk_ws, size_t chunk_bytes, int kc, int rows,
               int* status, bool split_weights, cudaStream_t st) {
    if (!encode_fn()) return tc_fail(STPDE_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
    if (n_layers < 3) return tc_fail(STPDE_EUNSUPPORTED, "the tensor-core path needs at least one hidden contraction");
    memset(&tc, 0, sizeof(tc));
    tc.n_layers = n_layers;
    tc.kc = kc;
    tc.rows = rows;
    tc.passes = precision == STPDE_PREC_FP16X3 ? 3 : 1;
    tc.status = status;
    tc.use_pair = tc_env().use_pair;
    // (Fusing layer 0 into layer 1's operand producer was tried in round 1 and measured slower - 905 ms vs 435 + 122 ms
    // at BASELINE config 2, the 4 feature-tile passes regenerate the operand 4 times - and removed in round 2.)
    // The tensor-core kernels use the MUFU-based activation jets (act_jet_fast): measured on B200, switching to the
    // libdevice-accurate versions changes the fp16x3 error by < 2 % (the 2^-22 operand rounding dominates) and
    // costs 4 % of the step.
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, dev);

    tc.wscale = (float*)fixed_ws;
    tc.absmax = (unsigned*)(fixed_ws + 512);
    if (split_weights) cudaMemsetAsync(tc.absmax, 0, 256, st);
    size_t off = 1024;
    int me, mo;
    plane_lds(n_layers, widths, me, mo);
    const size_t plane_even = (size_t)kc * rows * me * sizeof(__half), plane_odd = (size_t)kc * rows * mo * sizeof(__half);
    if (2 * (align_up(plane_even, 1024) + align_up(plane_odd, 1024)) > chunk_bytes)
        return tc_fail(STPDE_ENOMEM, "workspace too small for the fp16 activation planes");
    char* p = chunk_ws;
    tc.act[0][0] = (__half*)p; p += align_up(plane_even, 1024);
    tc.act[0][1] = (__half*)p; p += align_up(plane_even, 1024);
    tc.act[1][0] = (__half*)p; p += align_up(plane_odd, 1024);
    tc.act[1][1] = (__half*)p;
    tc.ld0 = round_up(widths[0], 64);
    tc.n0 = widths[0];

    if (split_weights) prof_begin(kSlotSetup, st);
    for (int l = 1; l <= n_layers - 2; ++l) {
        TcLayerPlan& L = tc.layer[l];
        L.n_feat = widths[l];
        L.np128 = round_up(widths[l], 128);
        L.kp_in = round_up(widths[l - 1], 64);
        L.last = (l == n_layers - 2);
        L.ld_out = L.last ? round_up(widths[l], 16) : round_up(widths[l], 64);
        L.n_store = L.ld_out;
        L.pack = tc_layer_pack(widths[l], L.last != 0);
        const size_t plane = align_up((size_t)L.np128 * L.kp_in * L.pack * sizeof(__half), 1024);
        L.w_hi_ptr = (__half*)(fixed_ws + off); off += plane;
        L.w_lo_ptr = (__half*)(fixed_ws + off); off += plane;
        const int kh = widths[l - 1];
        if (split_weights) {     // (a call that reuses the previous call's setup finds the planes and scales in place)
            absmax_kernel<<<148, 256, 0, st>>>(W[l], widths[l], in_features[l], kh, tc.absmax + l);
            split_weights_kernel<<<148 * 4, 256, 0, st>>>(W[l], widths[l], in_features[l], kh, L.np128, L.kp_in, L.pack,
                                                          tc.absmax + l, tc.wscale + l, L.w_hi_ptr, L.w_lo_ptr);
        }
        int rc = make_map_2d(&L.w_hi, L.w_hi_ptr, (uint64_t)L.kp_in * L.pack, L.np128, tc::kBlockK, tc::kTileF);
        rc |= make_map_2d(&L.w_lo, L.w_lo_ptr, (uint64_t)L.kp_in * L.pack, L.np128, tc::kBlockK, tc::kTileF);
        __half* in_hi = tc.act[(l - 1) & 1][0];
        __half* in_lo = tc.act[(l - 1) & 1][1];
        rc |= make_map_3d(&L.a_hi, in_hi, L.kp_in, rows, kc, tc::kBlockK, 8, kc);
        rc |= make_map_3d(&L.a_lo, in_lo, L.kp_in, rows, kc, tc::kBlockK, 8, kc);
        if (rc) { if (split_weights) prof_end(kSlotSetup, st, 0); return tc_fail(STPDE_ECUDA, "cuTensorMapEncodeTiled failed"); }
    }
    if (split_weights) prof_end(kSlotSetup, st, 2 * (n_layers - 2));
    return STPDE_OK;
}

// layer-0 launch: TMA-store variant unless STPDE_L0_TMA=0
template <int KC>
static void launch_layer0_any(const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb, int N, int ld,
                              const float* Wx, const float* Vb, int ncat, bool three, __half* out_hi, __half* out_lo,
                              int* status, cudaStream_t st) {
    // rows per block: 8 warps x rpw rows (the per-thread feature constants are amortised over rpw rows); small chunks
    // keep >= 16 blocks per SM
    int rpw = tc_env().l0_rows_per_warp;
    while (rpw > 8 && (int64_t)((ld + 127) / 128) * ((cb.rows + 8 * rpw - 1) / (8 * rpw)) < 16 * 148) rpw >>= 1;
    dim3 grid((ld + 127) / 128, (cb.rows + 8 * rpw - 1) / (8 * rpw));
    CUtensorMap m_hi, m_lo;
    bool tma = tc_env().l0_tma && encode_fn();
    if (tma) {
        cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)cb.rows, (cuuint64_t)KC};
        cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(__half), (cuuint64_t)ld * cb.rows * sizeof(__half)};
        cuuint32_t box[3] = {(cuuint32_t)(ld < 128 ? ld : 128), 1, (cuuint32_t)KC};
        cuuint32_t es[3] = {1, 1, 1};
        void* planes[2] = {out_hi, three ? (void*)out_lo : (void*)out_hi};
        CUtensorMap* maps[2] = {&m_hi, &m_lo};
        for (int i = 0; i < 2 && tma; ++i)
            tma = encode_fn()(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, planes[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        if (ld < 128) tma = false;                        // (the staging layout assumes 128-feature boxes)
    }
    if (!tma) { memset(&m_hi, 0, sizeof(m_hi)); memset(&m_lo, 0, sizeof(m_lo)); }
#define STPDE_L0_LAUNCH(THREE_, TMA_) layer0_jets_tc_kernel<KC, THREE_, TMA_><<<grid, 256, 0, st>>>(m_hi, m_lo, spec, dim, act, beta, cb.rows, rpw, N, ld, \
                                                                                      cb.vtx, cb.xrel, Wx, Vb, ncat, out_hi, out_lo, status)
    if (three) { if (tma) STPDE_L0_LAUNCH(true, true); else STPDE_L0_LAUNCH(true, false); }
    else       { if (tma) STPDE_L0_LAUNCH(false, true); else STPDE_L0_LAUNCH(false, false); }
#undef STPDE_L0_LAUNCH
}

void tc_launch_layer0_planes(int kc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb, int N, int ld,
                             const float* Wx, const float* Vb, int ncat, int three, __half* out_hi, __half* out_lo,
                             int* status, cudaStream_t st) {
    STPDE_TC_DISPATCH_KC(kc, (launch_layer0_any<KC>(spec, dim, act, beta, cb, N, ld, Wx, Vb, ncat, three != 0, out_hi, out_lo, status, st)));
}

bool tc_can_fuse_final(const TcContext& tc, const JetSpec& spec, int dim, int n_out) {
    const TcLayerPlan& L = tc.layer[tc.n_layers - 2];
    const bool pair = tc.use_pair && L.n_feat >= 2 * tc::kTileF;
    return tc_env().fuse_final && dim == 3 && spec_is_rb2(spec) && n_out <= 4 && !pair && L.n_feat <= tc::kTileF;
}

int tc_run_chunk(TcContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                 const float* Vb, int ncat, const int* cat_off, char* ws, const size_t* off_wx, float* act_last,
                 int np_last, const TcFinal* fused_final, cudaStream_t st) {
    if (cb.rows != tc.rows) return tc_fail(STPDE_EINVAL, "chunk geometry changed after tc_prepare");
    {
        ProfScope ps(kSlotLayer0, st);
        tc_launch_layer0_planes(spec.kc, spec, dim, act, beta, cb, tc.n0, tc.ld0, (const float*)(ws + off_wx[0]), Vb, ncat,
                                tc.passes == 3, tc.act[0][0], tc.act[0][1], tc.status, st);
    }
    for (int l = 1; l <= tc.n_layers - 2; ++l) {
        const TcLayerPlan& L = tc.layer[l];
        tc::LayerArgs a;
        memset(&a, 0, sizeof(a));
        a.rows = cb.rows;
        a.n_feat = L.n_feat;
        a.kp_in = L.kp_in;
        a.ld_out = L.last ? np_last : L.ld_out;
        a.n_store = a.ld_out;
        a.last = L.last;
        a.passes = tc.passes;
        a.pack = L.pack;
        a.dim = dim;
        a.act = act;
        a.ncat = ncat;
        a.cat_off = cat_off[l];
        a.beta = beta;
        a.wscale = tc.wscale + l;
        a.Wx = (const float*)(ws + off_wx[l]);
        a.Vb = Vb;
        a.vtx = cb.vtx;
        a.xrel = cb.xrel;
        a.out_hi = tc.act[l & 1][0];
        a.out_lo = tc.act[l & 1][1];
        a.out_f32 = act_last;
        a.status = tc.status;
        int rc = STPDE_OK;
        ProfScope ps(kSlotGemm + l - 1, st);
        a.wait_ns = tc_env().wait_ns;
        if (L.last && fused_final) {
            a.fuse_final = 1;
            a.n_out = fused_final->n_out;
            a.ldw_last = fused_final->ldw;
            a.w_last = fused_final->w_last;
            a.b_last = fused_final->b_last;
            a.pc = cb.pc;
            a.p0 = fused_final->p0;
            a.total_pts = fused_final->total_pts;
            a.wfac = cb.wfac;
            a.dfac = cb.dfac;
            a.dxr = cb.dxr;
            a.y = fused_final->y;
            a.jets = fused_final->jets;
        }
        rc = L.last ? tc_encode_out_maps(a, spec.kc, act_last, nullptr, true)
                    : tc_encode_out_maps(a, spec.kc, a.out_hi, tc.passes == 3 ? (void*)a.out_lo : nullptr, false);
        if (rc) return rc;
        if (tc.use_pair && L.n_feat >= 2 * tc::kTileF) {
            rc = tc_launch_layer_pair(spec.kc, tc, L, spec, a, st);
        } else {
            rc = tc_launch_layer(spec.kc, tc, L, spec, a, st);
        }
        if (rc) return rc;
    }
    return STPDE_OK;
}

}  // namespace stpde
