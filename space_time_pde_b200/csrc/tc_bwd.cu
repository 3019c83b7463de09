// Host orchestration of the reverse-mode sweep on the tensor cores (see tc_bwd.h).
#include "tc_bwd.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bwd_kernels.h"
#include "profile.h"
#include "tc_launch.cuh"
#include "tc_path.h"
#include "tc_wgrad.cuh"

namespace stpde {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// sizes
// ---------------------------------------------------------------------------------------------
size_t tc_bwd_fixed_bytes(int n_layers, const int* widths) {
    size_t off = 1024;   // wscale + absmax
    for (int l = 1; l <= n_layers - 2; ++l) {
        const int pack = tc_layer_pack(widths[l], l == n_layers - 2);       // block-diagonal forward operand [128][pack * ld]
        const size_t w_plane = (size_t)(pack > 1 ? 128 * pack : round_up(widths[l], 256)) * round_up(widths[l - 1], 64) * sizeof(__half);
        const size_t wt_plane = (size_t)round_up(widths[l - 1], 256) * round_up(widths[l], 64) * sizeof(__half);
        off += 2 * align_up(w_plane, 1024) + 2 * align_up(wt_plane, 1024);
    }
    return off;
}

// bytes per (point, corner) row and jet component
static size_t row_bytes(int n_layers, const int* widths) {
    size_t b = 0;
    for (int l = 0; l <= n_layers - 3; ++l) b += (size_t)2 * round_up(widths[l], 64) * sizeof(__half);   // a_l hi / lo
    for (int l = 1; l <= n_layers - 2; ++l) b += (size_t)round_up(widths[l], 64) * sizeof(float);        // z_l
    int me = 64, mo = 64;                                                                                // zbar ping-pong
    for (int l = 1; l <= n_layers - 2; ++l) {
        const int ld = round_up(widths[l], 64);
        if (l & 1) mo = ld > mo ? ld : mo; else me = ld > me ? ld : me;
    }
    b += (size_t)2 * (me + mo) * sizeof(__half);
    return b;
}

size_t tc_bwd_per_point_bytes(int n_layers, const int* widths, int kc, int ncorner) {
    return row_bytes(n_layers, widths) * kc * ncorner;
}

// ---------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------
int tc_bwd_prepare(TcBwdContext& tc, int precision, int n_layers, const int* widths, const int* in_features,
                   const float* const* W, char* fixed_ws, char* chunk_ws, size_t chunk_bytes, int kc, int rows,
                   int* status, bool split_weights, bool z_half, cudaStream_t st) {
    if (!tc_encode_available()) return tc_fail(STPDE_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
    if (n_layers < 3) return tc_fail(STPDE_EUNSUPPORTED, "the tensor-core backward needs at least one hidden contraction");
    if (rows % tc::kWgKBlock) return tc_fail(STPDE_EINVAL, "chunk rows must be a multiple of 64");
    memset(&tc, 0, sizeof(tc));
    tc.n_layers = n_layers;
    tc.kc = kc;
    tc.rows = rows;
    tc.passes = precision == STPDE_PREC_FP16 ? 1 : 3;
    tc.status = status;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, dev);
    tc.wscale = (float*)fixed_ws;
    tc.absmax = (unsigned*)(fixed_ws + 512);
    if (split_weights) cudaMemsetAsync(tc.absmax, 0, 256, st);
    tc.ld0 = round_up(widths[0], 64);
    tc.n0 = widths[0];
    tc.use_pair_wide = tc_env().use_pair;
    tc.z_half = z_half ? 1 : 0;

    // chunk planes
    char* p = chunk_ws;
    char* const p_end = chunk_ws + chunk_bytes;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 1024); return r; };
    const size_t rk = (size_t)rows * kc;
    __half* a_planes[kMaxLayers][2];
    for (int l = 0; l <= n_layers - 3; ++l)
        for (int h = 0; h < 2; ++h) a_planes[l][h] = (__half*)take(rk * round_up(widths[l], 64) * sizeof(__half));
    int me = 64, mo = 64;
    for (int l = 1; l <= n_layers - 2; ++l) {
        const int ld = round_up(widths[l], 64);
        if (l & 1) mo = ld > mo ? ld : mo; else me = ld > me ? ld : me;
    }
    __half* zb_planes[2][2];
    for (int par = 0; par < 2; ++par)
        for (int h = 0; h < 2; ++h) zb_planes[par][h] = (__half*)take(rk * (par ? mo : me) * sizeof(__half));
    for (int l = 1; l <= n_layers - 2; ++l) tc.layer[l].z = (float*)take(rk * round_up(widths[l], 64) * sizeof(float));
    if (p > p_end) return tc_fail(STPDE_ENOMEM, "workspace too small for the backward activation planes");

    size_t off = 1024;
    int rc = 0;
    prof_begin(kSlotSetup, st);
    for (int l = 1; l <= n_layers - 2; ++l) {
        TcBwdLayer& L = tc.layer[l];
        L.n_feat = widths[l];
        L.kh = widths[l - 1];
        L.ldz = round_up(widths[l], 64);
        L.ld_in = round_up(widths[l - 1], 64);
        L.last = (l == n_layers - 2);
        L.pack = tc_layer_pack(widths[l], L.last != 0);
        // dgrad of a layer whose INPUT has <= 64 features (MODE 2 only: layer 1's dgrad reverses the closed-form layer 0):
        // 2 row groups per 128-lane tile, block-diagonal W^T [128][2 * ldz] (same bytes as the [256][ldz] plane)
        L.pack_t = (l >= 2 && widths[l - 1] <= 64 && tc_env().pack_narrow) ? 2 : 1;
        const int np256 = L.pack > 1 ? 128 : round_up(widths[l], 256), fp256 = L.pack_t > 1 ? 128 : round_up(widths[l - 1], 256);
        const size_t w_plane = align_up((size_t)np256 * L.ld_in * L.pack * sizeof(__half), 1024);
        const size_t wt_plane = align_up((size_t)fp256 * L.ldz * L.pack_t * sizeof(__half), 1024);
        L.w_hi_ptr = (__half*)(fixed_ws + off); off += w_plane;
        L.w_lo_ptr = (__half*)(fixed_ws + off); off += w_plane;
        L.wt_hi_ptr = (__half*)(fixed_ws + off); off += wt_plane;
        L.wt_lo_ptr = (__half*)(fixed_ws + off); off += wt_plane;
        if (split_weights) {     // (a backward that reuses the training forward's workspace finds them in place)
            tc_launch_split_weights(W[l], widths[l], in_features[l], L.kh, np256, L.ld_in, L.pack, tc.absmax + l, tc.wscale + l,
                                    L.w_hi_ptr, L.w_lo_ptr, st);
            launch_split_weights_t(W[l], widths[l], in_features[l], L.kh, fp256, L.ldz, L.pack_t, tc.absmax + l, L.wt_hi_ptr, L.wt_lo_ptr, st);
        }
        for (int h = 0; h < 2; ++h) {
            L.a_in[h] = a_planes[l - 1][h];
            L.a_out[h] = L.last ? nullptr : a_planes[l][h];
            L.zb[h] = zb_planes[l & 1][h];
        }
        rc |= tc_make_map_2d(&L.w_hi, L.w_hi_ptr, (uint64_t)L.ld_in * L.pack, np256, tc::kBlockK, tc::kTileF);
        rc |= tc_make_map_2d(&L.w_lo, L.w_lo_ptr, (uint64_t)L.ld_in * L.pack, np256, tc::kBlockK, tc::kTileF);
        rc |= tc_make_map_3d(&L.fa_hi, L.a_in[0], L.ld_in, rows, kc, tc::kBlockK, 8, kc);
        rc |= tc_make_map_3d(&L.fa_lo, L.a_in[1], L.ld_in, rows, kc, tc::kBlockK, 8, kc);
        rc |= tc_make_map_2d(&L.wt_hi, L.wt_hi_ptr, (uint64_t)L.ldz * L.pack_t, fp256, tc::kBlockK, tc::kTileF);
        rc |= tc_make_map_2d(&L.wt_lo, L.wt_lo_ptr, (uint64_t)L.ldz * L.pack_t, fp256, tc::kBlockK, tc::kTileF);
        rc |= tc_make_map_3d(&L.zb_hi, L.zb[0], L.ldz, rows, kc, tc::kBlockK, 8, kc);
        rc |= tc_make_map_3d(&L.zb_lo, L.zb[1], L.ldz, rows, kc, tc::kBlockK, 8, kc);
        rc |= tc_make_map_3d(&L.ga_hi, L.a_in[0], L.ld_in, rows, kc, 64, tc::kWgKBlock, 1);
        rc |= tc_make_map_3d(&L.ga_lo, L.a_in[1], L.ld_in, rows, kc, 64, tc::kWgKBlock, 1);
        rc |= tc_make_map_3d(&L.gb_hi, L.zb[0], L.ldz, rows, kc, 64, tc::kWgKBlock, 1);
        rc |= tc_make_map_3d(&L.gb_lo, L.zb[1], L.ldz, rows, kc, 64, tc::kWgKBlock, 1);
    }
    prof_end(kSlotSetup, st, 3 * (n_layers - 2));
    if (rc) return tc_fail(STPDE_ECUDA, "cuTensorMapEncodeTiled failed (backward maps)");
    return STPDE_OK;
}

// ---------------------------------------------------------------------------------------------
// forward recompute
// ---------------------------------------------------------------------------------------------
static void base_args(tc::LayerArgs& a, const TcBwdContext& tc, int dim, int act, float beta, const ChunkBuffers& cb,
                      const float* Vb, int ncat) {
    memset(&a, 0, sizeof(a));
    a.rows = cb.rows;
    a.passes = tc.passes;
    a.dim = dim;
    a.act = act;
    a.beta = beta;
    a.ncat = ncat;
    a.Vb = Vb;
    a.vtx = cb.vtx;
    a.xrel = cb.xrel;
    a.status = tc.status;
    a.wait_ns = tc_env().wait_ns;
}

bool tc_bwd_can_fuse_final(const TcBwdContext& tc, const JetSpec& spec, int dim, int n_out) {
    const TcBwdLayer& L = tc.layer[tc.n_layers - 2];
    const bool pair = L.n_feat >= 2 * tc::kTileF && tc.use_pair_wide;
    return tc_env().fuse_final && dim == 3 && spec_is_rb2(spec) && n_out <= 4 && !pair && L.n_feat <= tc::kTileF;
}

int tc_bwd_forward_chunk(TcBwdContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                         const float* Vb, int ncat, const int* cat_off, const float* const* Wx, float* act_last,
                         int np_last, const TcFinal* fused_final, cudaStream_t st) {
    if (cb.rows != tc.rows) return tc_fail(STPDE_EINVAL, "chunk geometry changed after tc_bwd_prepare");
    {
        ProfScope ps(kSlotLayer0, st);
        tc_launch_layer0_planes(spec.kc, spec, dim, act, beta, cb, tc.n0, tc.ld0, Wx[0], Vb, ncat, tc.passes == 3,
                                tc.layer[1].a_in[0], tc.layer[1].a_in[1], tc.status, st);
    }
    for (int l = 1; l <= tc.n_layers - 2; ++l) {
        const TcBwdLayer& L = tc.layer[l];
        tc::LayerArgs a;
        base_args(a, tc, dim, act, beta, cb, Vb, ncat);
        a.n_feat = L.n_feat;
        a.kp_in = L.ld_in;
        a.ld_out = L.last ? np_last : L.ldz;
        a.n_store = a.ld_out;
        a.last = L.last;
        a.pack = L.pack;
        a.cat_off = cat_off[l];
        a.wscale = tc.wscale + l;
        a.Wx = Wx[l];
        a.out_hi = L.a_out[0];
        a.out_lo = L.a_out[1];
        a.out_f32 = act_last;
        a.z_out = L.z;
        a.ldz = L.ldz;
        a.z_half = tc.z_half;
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
        int rc = L.last ? tc_encode_out_maps(a, spec.kc, act_last, nullptr, true)
                        : tc_encode_out_maps(a, spec.kc, a.out_hi, tc.passes == 3 ? (void*)a.out_lo : nullptr, false);
        if (rc) return rc;
        ProfScope ps(kSlotGemm + l - 1, st);
        // features >= 256: CTA-pair tile (M = 256); narrower layers: one 128-feature CTA per SM
        rc = L.n_feat >= 2 * tc::kTileF && tc.use_pair_wide
                     ? tc_launch_pair_save(spec.kc, tc.num_sms, L.w_hi, L.w_lo, L.fa_hi, L.fa_lo, spec, a, st)
                     : tc_launch_single_save(spec.kc, tc.num_sms, L.w_hi, L.w_lo, L.fa_hi, L.fa_lo, spec, a, st);
        if (rc) return rc;
    }
    return STPDE_OK;
}

// ---------------------------------------------------------------------------------------------
// reverse sweep
// ---------------------------------------------------------------------------------------------
static int launch_wgrad(const TcBwdContext& tc, const TcBwdLayer& L, float* gW, int ldw, cudaStream_t st) {
    tc::WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = tc.rows;
    a.kc = tc.kc;
    a.nf_a = L.kh;
    a.nf_b = L.n_feat;
    a.nt = L.ldz <= 128 ? 128 : 256;
    a.n_ft = (L.kh + 255) / 256;
    a.n_gt = (L.n_feat + a.nt - 1) / a.nt;
    const int n_pairs_max = tc.num_sms / 2;
    const int64_t total_kb = (int64_t)tc.kc * (tc.rows / tc::kWgKBlock);
    // K slices: enough units for >= 2 rounds over the CTA pairs, picked so that the last round is as full as possible
    const int tiles = a.n_ft * a.n_gt;
    int64_t s_lo = (2 * n_pairs_max + tiles - 1) / tiles, s_hi = (8 * n_pairs_max + tiles - 1) / tiles;
    if (s_hi > total_kb / 4) s_hi = total_kb / 4;
    if (s_lo > s_hi) s_lo = s_hi;
    if (s_lo < 1) s_lo = s_hi = 1;
    int64_t slices = s_lo;
    double best = 0.0;
    for (int64_t sl = s_lo; sl <= s_hi; ++sl) {
        const int64_t units = tiles * sl, rounds = (units + n_pairs_max - 1) / n_pairs_max;
        const double eff = (double)units / (double)(rounds * n_pairs_max);
        if (eff > best + 1e-9) { best = eff; slices = sl; }
    }
    a.n_slices = (int)slices;
    a.tile_fastest = tc_env().wgrad_tile_fastest;
    a.passes = tc.passes;
    a.out_scale = 1.f / (float)(1 << tc::kActScaleLog2);
    a.gW = gW;
    a.ldw = ldw;
    a.status = tc.status;
    const size_t smem = (size_t)tc::kPairSmemBudget + 1024 + 512;
    static DeviceOnce configured;
    if (configured.first_use()) {
        if (cudaFuncSetAttribute(tc::tc_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "cudaFuncSetAttribute(tc_wgrad_pair_kernel) failed");
        configured.mark();
    }
    const int n_units = a.n_ft * a.n_gt * a.n_slices;
    const int n_pairs = n_units < n_pairs_max ? n_units : n_pairs_max;
    {
        void* kargs[] = {(void*)&L.ga_hi, (void*)&L.ga_lo, (void*)&L.gb_hi, (void*)&L.gb_lo, (void*)&a};
        if (tc_launch_ex((const void*)tc::tc_wgrad_pair_kernel, 2 * n_pairs, tc::kThreads, smem, st, kargs) != cudaSuccess)
            return tc_fail(STPDE_ECUDA, "tc_wgrad_pair_kernel launch failed");
    }
    return STPDE_OK;
}

int tc_bwd_backward_chunk(TcBwdContext& tc, const JetSpec& spec, int dim, int act, float beta, const ChunkBuffers& cb,
                          const float* Vb, int ncat, const int* cat_off, const int* in_features, const float* const* Wx,
                          float* const* gW, float* g_vb, float* g_beta, cudaStream_t st) {
    for (int l = tc.n_layers - 2; l >= 1; --l) {
        const TcBwdLayer& L = tc.layer[l];
        {
            ProfScope ps(kSlotWgrad + l - 1, st);
            int rc = launch_wgrad(tc, L, gW[l], in_features[l], st);
            if (rc) return rc;
        }
        // dgrad: adjoints of layer l-1 (features = widths[l-1]), contraction over this layer's features
        tc::LayerArgs a;
        base_args(a, tc, dim, act, beta, cb, Vb, ncat);
        a.n_feat = L.kh;
        a.kp_in = L.ldz;
        a.ld_out = L.ld_in;
        a.n_store = L.ld_in;
        a.cat_off = cat_off[l - 1];
        a.wscale = tc.wscale + l;
        a.Wx = Wx[l - 1];
        a.g_vb = g_vb;
        a.g_beta = g_beta;
        const int kh_below = l >= 2 ? tc.layer[l - 1].kh : 0;      // activation columns of layer l-1's weight
        a.g_wx = gW[l - 1] + kh_below;
        a.g_wx_ld = in_features[l - 1];
        ProfScope ps(kSlotDgrad + l - 1, st);
        const bool wide = L.kh >= 2 * tc::kTileF && tc.use_pair_wide;   // M extent of the dgrad = features of layer l-1
        int rc = STPDE_OK;
        if (l >= 2) {
            a.pack = wide ? 0 : L.pack_t;
            a.z_half = tc.z_half;
            a.z_in = tc.layer[l - 1].z;
            a.ldz = tc.layer[l - 1].ldz;
            a.out_hi = tc.layer[l - 1].zb[0];
            a.out_lo = tc.layer[l - 1].zb[1];
            rc = tc_encode_out_maps(a, spec.kc, a.out_hi, tc.passes == 3 ? (void*)a.out_lo : nullptr, false, true);
            if (rc) return rc;
            rc = wide ? tc_launch_pair_bwd(spec.kc, tc.num_sms, L.wt_hi, L.wt_lo, L.zb_hi, L.zb_lo, spec, a, st)
                      : tc_launch_single_bwd(spec.kc, tc.num_sms, L.wt_hi, L.wt_lo, L.zb_hi, L.zb_lo, spec, a, st);
        } else {
            rc = wide ? tc_launch_pair_bwd0(spec.kc, tc.num_sms, L.wt_hi, L.wt_lo, L.zb_hi, L.zb_lo, spec, a, st)
                      : tc_launch_single_bwd0(spec.kc, tc.num_sms, L.wt_hi, L.wt_lo, L.zb_hi, L.zb_lo, spec, a, st);
        }
        if (rc) return rc;
    }
    return STPDE_OK;
}

}  // namespace stpde
