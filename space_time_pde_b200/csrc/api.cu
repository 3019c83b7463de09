// C-ABI entry points (include/stpde.h): descriptor validation, workspace planning and the
// per-chunk launch sequence.  No torch types, no allocation on the device path, never throws.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bwd_kernels.h"
#include "kernels.h"
#include "profile.h"
#include "tc_bwd.h"
#include "tc_path.h"

namespace stpde {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) return fail(STPDE_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

// float32 clip bounds / cube sizes exactly as regular_nd_grid_interpolation.py:47-51 forms them
static int make_geom(GridGeom& g, int dim, const int32_t* size, int channels, const float* xmin, const float* xmax,
                     const int64_t* gstrides, const int64_t* qstrides) {
    if (dim < 1 || dim > kMaxDim) return fail(STPDE_EINVAL, "dim must be in 1..%d (got %d)", kMaxDim, dim);
    if (channels < 1) return fail(STPDE_EINVAL, "channels must be >= 1");
    memset(&g, 0, sizeof(g));
    g.dim = dim;
    g.channels = channels;
    g.nvert = 1;
    for (int k = 0; k < dim; ++k) {
        if (size[k] < 1) return fail(STPDE_EINVAL, "grid_size[%d] = %d", k, size[k]);
        g.size[k] = size[k];
        g.nvert *= size[k];
        volatile float span = xmax[k] - xmin[k];
        volatile float eps = 1e-6f * span;
        volatile float lo = xmin[k] + eps;
        volatile float hi = xmax[k] - eps;
        volatile float cs = span / ((float)size[k] - 1.0f);
        g.lo[k] = lo;
        g.hi[k] = hi;
        g.cubesize[k] = cs;
    }
    if (gstrides) for (int k = 0; k < dim + 2; ++k) g.gstride[k] = gstrides[k];
    if (qstrides) for (int k = 0; k < 3; ++k) g.qstride[k] = qstrides[k];
    return STPDE_OK;
}

struct Plan {
    GridGeom geom;
    JetSpec spec;
    int n_layers, O, ncorner, nvert_total;
    int widths[kMaxLayers], np[kMaxLayers], np64[kMaxLayers], kh[kMaxLayers], kp[kMaxLayers], in_features[kMaxLayers];
    int cat_off[kMaxLayers], ncat;
    int max_even, max_odd;
    int64_t total_pts;
    // workspace layout
    size_t off_wh[kMaxLayers], off_wx[kMaxLayers], off_vb, off_tc, fixed_bytes, per_point_bytes;
    // reverse mode (always on the tensor cores): fixed region = [0, off_tc) as above, then the split weights and
    // their transposes, the adjoint of Vb and the scale scratch
    size_t off_bwd_gvb, off_bwd_scale, bwd_fixed_bytes, bwd_per_point_bytes;
};

static int make_plan(Plan& P, const stpde_desc_t* d, const int64_t* gstrides, const int64_t* qstrides) {
    if (!d) return fail(STPDE_EINVAL, "null descriptor");
    if (d->batch < 0 || d->npts < 0) return fail(STPDE_EINVAL, "negative batch/npts");
    int rc = make_geom(P.geom, d->dim, d->grid_size, d->channels, d->xmin, d->xmax, gstrides, qstrides);
    if (rc) return rc;
    if (d->n_layers < 2 || d->n_layers > kMaxLayers) return fail(STPDE_EINVAL, "n_layers must be in 2..%d", kMaxLayers);
    if (d->act_kind < 0 || d->act_kind > STPDE_ACT_LEAKYRELU) return fail(STPDE_EINVAL, "unknown activation %d", d->act_kind);
    P.n_layers = d->n_layers;
    P.O = d->widths[d->n_layers - 1];
    if (P.O < 1 || P.O > kMaxOut) return fail(STPDE_EUNSUPPORTED, "out_features must be in 1..%d (got %d)", kMaxOut, P.O);
    P.ncorner = 1 << d->dim;
    P.nvert_total = d->batch * P.geom.nvert;
    P.total_pts = (int64_t)d->batch * d->npts;
    const int D = d->dim + d->channels;
    P.ncat = 0;
    P.max_even = P.max_odd = 0;
    for (int l = 0; l < P.n_layers; ++l) {
        if (d->widths[l] < 1) return fail(STPDE_EINVAL, "widths[%d] = %d", l, d->widths[l]);
        P.widths[l] = d->widths[l];
        P.np[l] = round_up(d->widths[l], 16);
        P.np64[l] = round_up(d->widths[l], 64);
        P.kh[l] = l == 0 ? 0 : d->widths[l - 1];
        P.kp[l] = l == 0 ? 0 : P.np[l - 1];
        P.in_features[l] = P.kh[l] + (l < P.n_layers - 1 ? D : 0);
        if (l < P.n_layers - 1) {
            P.cat_off[l] = P.ncat;
            P.ncat += d->widths[l];
            if (l % 2 == 0) P.max_even = P.np[l] > P.max_even ? P.np[l] : P.max_even;
            else P.max_odd = P.np[l] > P.max_odd ? P.np[l] : P.max_odd;
        }
    }
    // jet specification
    JetSpec& s = P.spec;
    memset(&s, 0, sizeof(s));
    if (d->n_first < 0 || d->n_first > STPDE_MAX_FIRST || d->n_second < 0 || d->n_second > STPDE_MAX_SECOND)
        return fail(STPDE_EINVAL, "bad jet specification");
    s.n_first = d->n_first;
    s.n_second = d->n_second;
    s.kc = 1 + s.n_first + s.n_second;
    if (s.kc > kMaxComp) return fail(STPDE_EUNSUPPORTED, "%d jet components > %d per call; split the second-order set", s.kc, kMaxComp);
    for (int i = 0; i < s.n_first; ++i) {
        if (d->first_dirs[i] < 0 || d->first_dirs[i] >= d->dim) return fail(STPDE_EINVAL, "first_dirs[%d] out of range", i);
        s.first_dirs[i] = d->first_dirs[i];
    }
    for (int i = 0; i < s.n_second; ++i) {
        int ca = -1, cb = -1;
        for (int j = 0; j < s.n_first; ++j) {
            if (s.first_dirs[j] == d->second_pairs[i][0]) ca = j + 1;
            if (s.first_dirs[j] == d->second_pairs[i][1]) cb = j + 1;
        }
        if (ca < 0 || cb < 0) return fail(STPDE_EINVAL, "second_pairs[%d] needs both directions in first_dirs", i);
        s.sec_a[i] = ca;
        s.sec_b[i] = cb;
    }
    for (int c = 1; c < s.kc; ++c) {
        if (c <= s.n_first) { s.kind[c] = 1; s.dir[c] = s.first_dirs[c - 1]; }
        else {
            s.kind[c] = 2; s.pa[c] = s.sec_a[c - 1 - s.n_first]; s.pb[c] = s.sec_b[c - 1 - s.n_first];
            s.sel_a[c][s.pa[c] - 1] = 1.f;
            s.sel_b[c][s.pb[c] - 1] = 1.f;
        }
    }
    // fixed workspace region
    size_t off = 256;  // status / scratch words
    for (int l = 0; l < P.n_layers; ++l) {
        P.off_wh[l] = P.off_wx[l] = 0;
        if (l >= 1) {
            P.off_wh[l] = off;
            int rows = (l == P.n_layers - 1) ? P.widths[l] : P.np64[l];
            off = align_up(off + (size_t)rows * P.kp[l] * sizeof(float), 256);
        }
        if (l < P.n_layers - 1) {
            P.off_wx[l] = off;
            off = align_up(off + (size_t)P.widths[l] * d->dim * sizeof(float), 256);
        }
    }
    P.off_vb = off;
    off = align_up(off + (size_t)(P.nvert_total > 0 ? P.nvert_total : 1) * P.ncat * sizeof(float), 256);
    off = align_up(off, 1024);
    P.off_tc = off;
    const bool use_tc = d->precision != STPDE_PREC_FP32;
    if (d->precision < STPDE_PREC_FP32 || d->precision > STPDE_PREC_FP16) return fail(STPDE_EINVAL, "unknown precision %d", d->precision);
    off = align_up(off + (use_tc ? tc_fixed_bytes(P.n_layers, P.widths) : 0), 1024);
    P.fixed_bytes = off;
    const int kc = s.kc;
    P.per_point_bytes = (size_t)P.ncorner * (4 + 4 * kMaxDim) + (size_t)4 * 5 * d->dim;
    if (use_tc)   // fp32 only for the last hidden layer (input of final_blend) + fp16 hi/lo planes
        P.per_point_bytes += (size_t)kc * P.ncorner * 4 * P.np[P.n_layers - 2] +
                             tc_per_point_bytes(P.n_layers, P.widths, kc, P.ncorner);
    else
        P.per_point_bytes += (size_t)kc * P.ncorner * 4 * ((size_t)P.max_even + P.max_odd);
    {
        size_t boff = align_up(P.off_tc + (P.n_layers >= 3 ? tc_bwd_fixed_bytes(P.n_layers, P.widths) : 0), 1024);
        P.off_bwd_gvb = boff;
        boff = align_up(boff + (size_t)(P.nvert_total > 0 ? P.nvert_total : 1) * P.ncat * sizeof(float), 1024);
        P.off_bwd_scale = boff;
        P.bwd_fixed_bytes = boff + 1024;
        P.bwd_per_point_bytes = (size_t)P.ncorner * (4 + 4 * kMaxDim) + (size_t)4 * 5 * d->dim +
                                (size_t)kc * P.ncorner * 4 * P.np[P.n_layers - 2] +
                                (P.n_layers >= 3 ? tc_bwd_per_point_bytes(P.n_layers, P.widths, kc, P.ncorner) : 0);
    }
    return STPDE_OK;
}

static size_t chunk_region_bytes(const Plan& P, int64_t pc) { return (size_t)pc * P.per_point_bytes + 16 * 1024; }

static int64_t default_chunk_points(const Plan& P) {
    size_t budget_mb = 6144;
    if (const char* e = getenv("STPDE_WORKSPACE_MB")) budget_mb = (size_t)atoll(e);
    int64_t pc = (int64_t)((budget_mb << 20) / (P.per_point_bytes ? P.per_point_bytes : 1));
    pc = pc / 128 * 128;
    int64_t need = (P.total_pts + 127) / 128 * 128;
    if (pc > need) pc = need;
    if (pc < 128) pc = 128;
    return pc;
}

static int64_t balanced_chunk_points(int64_t total_pts, int64_t pc_max);

static int run_forward(const Plan& P, const stpde_desc_t* d, const float* grid, const float* q,
                       const float* const* W, const float* const* B, float* y, float* jets, char* ws,
                       size_t ws_bytes, int* status, cudaStream_t st) {
    if (P.total_pts == 0) return STPDE_OK;
    if (ws_bytes < P.fixed_bytes + chunk_region_bytes(P, 128))
        return fail(STPDE_ENOMEM, "workspace %zu B < minimum %zu B", ws_bytes, P.fixed_bytes + chunk_region_bytes(P, 128));
    int64_t pc = (int64_t)((ws_bytes - P.fixed_bytes - 16 * 1024) / P.per_point_bytes) / 128 * 128;
    int64_t need = (P.total_pts + 127) / 128 * 128;
    if (pc > need) pc = need;
    if (pc > (1 << 24)) pc = 1 << 24;
    pc = balanced_chunk_points(P.total_pts, pc);
    const int dim = d->dim, kc = P.spec.kc;
    const int64_t rows = pc * P.ncorner;
    if (rows * (int64_t)(P.max_even > P.max_odd ? P.max_even : P.max_odd) * kc >= (int64_t)1 << 40)
        return fail(STPDE_EUNSUPPORTED, "chunk too large");

    // carve the chunk region
    char* p = ws + P.fixed_bytes;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
    ChunkBuffers cb;
    cb.pc = (int)pc;
    cb.rows = (int)rows;
    cb.vtx = (int*)take(rows * 4);
    cb.xrel = (float*)take((size_t)kMaxDim * rows * 4);   // [kMaxDim][rows], planes >= dim are zero
    cb.wfac = (float*)take((size_t)dim * 2 * pc * 4);
    cb.dfac = (float*)take((size_t)dim * 2 * pc * 4);
    cb.dxr = (float*)take((size_t)dim * pc * 4);
    const bool use_tc = d->precision != STPDE_PREC_FP32;
    float* act[2] = {nullptr, nullptr};
    if (use_tc) {
        act[(P.n_layers - 2) & 1] = (float*)take((size_t)kc * rows * P.np[P.n_layers - 2] * 4);
    } else {
        act[0] = (float*)take((size_t)kc * rows * P.max_even * 4);
        act[1] = (float*)take((size_t)kc * rows * (P.max_odd > 0 ? P.max_odd : 1) * 4);
    }
    p = (char*)align_up((size_t)p, 1024);
    char* tc_chunk = p;

    // once per call: pack weights, per-vertex latent/bias terms
    NetDesc net;
    memset(&net, 0, sizeof(net));
    net.n_layers = P.n_layers;
    net.ncat = P.ncat;
    for (int l = 0; l < P.n_layers; ++l) {
        net.cat_off[l] = P.cat_off[l];
        net.in_features[l] = P.in_features[l];
        net.kh[l] = P.kh[l];
        net.W[l] = W[l];
        net.B[l] = B[l];
    }
    float* Vb = (float*)(ws + P.off_vb);
    // reserved[1] = 1: the caller vouches that the call-invariant region (packed / split weights, Vb) is still valid
    const bool reuse_setup = d->reserved[1] == 1;
    if (!reuse_setup) {
        prof_begin(kSlotSetup, st);
        for (int l = 0; l < P.n_layers; ++l) {
            float* wh = l >= 1 ? (float*)(ws + P.off_wh[l]) : nullptr;
            float* wx = l < P.n_layers - 1 ? (float*)(ws + P.off_wx[l]) : nullptr;
            int rows_w = (l == P.n_layers - 1) ? P.widths[l] : P.np64[l];
            if (l >= 1) launch_pack_weights(W[l], P.widths[l], P.in_features[l], P.kh[l], dim, rows_w, P.kp[l], wh, wx, st);
            else launch_pack_weights(W[l], P.widths[l], P.in_features[l], 0, dim, 0, 1, nullptr, wx, st);
        }
        launch_vertex_bias(P.geom, P.nvert_total, net, grid, Vb, st);
        prof_end(kSlotSetup, st, P.n_layers + 1);
    }

    TcContext tc;
    if (use_tc) {
        int rc = tc_prepare(tc, d->precision, P.n_layers, P.widths, P.in_features, W, ws + P.off_tc, tc_chunk,
                            (size_t)(ws + ws_bytes - tc_chunk), kc, (int)rows, status, !reuse_setup, st);
        if (rc) return fail(rc, "%s", tc_last_error());
    }

    bool fused_final = false;
    for (int64_t p0 = 0; p0 < P.total_pts; p0 += pc) {
        {
            ProfScope ps(kSlotPrep, st);
            launch_prep_points(P.geom, d->npts, P.total_pts, p0, cb, q, status, st);
        }
        const int L = P.n_layers - 1;
        if (use_tc) {
            // last linear layer + corner blend inside the last hidden layer's epilogue when the shape allows it
            TcFinal fin;
            fin.w_last = (const float*)(ws + P.off_wh[L]);
            fin.b_last = B[L];
            fin.n_out = P.O;
            fin.ldw = P.kp[L];
            fin.y = y;
            fin.jets = jets;
            fin.p0 = p0;
            fin.total_pts = P.total_pts;
            fused_final = tc_can_fuse_final(tc, P.spec, dim, P.O);
            int rc = tc_run_chunk(tc, P.spec, dim, d->act_kind, d->act_param, cb, Vb, P.ncat, P.cat_off, ws, P.off_wx,
                                  act[(P.n_layers - 2) & 1], P.np[P.n_layers - 2], fused_final ? &fin : nullptr, st);
            if (rc) return fail(rc, "%s", tc_last_error());
        } else {
            {
                ProfScope ps(kSlotLayer0, st);
                launch_layer0(P.spec, dim, d->act_kind, d->act_param, cb.rows, P.widths[0], P.np[0], cb.vtx, cb.xrel,
                              (const float*)(ws + P.off_wx[0]), Vb, P.ncat, act[0], st);
            }
            for (int l = 1; l < P.n_layers - 1; ++l) {
                ProfScope ps(kSlotGemm + l - 1, st);
                launch_layer_gemm(P.spec, dim, d->act_kind, d->act_param, cb.rows, P.widths[l], P.np64[l], P.kp[l],
                                  P.np[l], act[(l - 1) & 1], (const float*)(ws + P.off_wh[l]),
                                  (const float*)(ws + P.off_wx[l]), Vb, P.ncat, P.cat_off[l], cb.vtx, cb.xrel,
                                  act[l & 1], st);
            }
        }
        if (!fused_final) {
            ProfScope ps(kSlotFinal, st);
            launch_final_blend(P.spec, dim, cb.rows, cb.pc, P.total_pts, p0, P.kp[L], P.O, act[(L - 1) & 1],
                               (const float*)(ws + P.off_wh[L]), B[L], cb, y, jets, st);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

// Every chunk runs the kernels over the full chunk geometry, so a short last chunk wastes the difference (and its padding
// rows all address vertex 0: their zero-valued adjoint atomics serialise on one table row - measured 4x on a 122k + 9k
// split).  Split the call into equal chunks instead: same count, each at most pc_max points.
static int64_t balanced_chunk_points(int64_t total_pts, int64_t pc_max) {
    if (pc_max < 128 || total_pts <= pc_max) return pc_max;
    const int64_t n_chunks = (total_pts + pc_max - 1) / pc_max;
    const int64_t pc = ((total_pts + n_chunks - 1) / n_chunks + 127) / 128 * 128;
    return pc < pc_max ? pc : pc_max;
}

static size_t bwd_chunk_region_bytes(const Plan& P, int64_t pc) { return (size_t)pc * P.bwd_per_point_bytes + 64 * 1024; }

static int64_t default_bwd_chunk_points(const Plan& P) {
    size_t budget_mb = 6144;
    if (const char* e = getenv("STPDE_WORKSPACE_MB")) budget_mb = (size_t)atoll(e);
    int64_t pc = (int64_t)((budget_mb << 20) / (P.bwd_per_point_bytes ? P.bwd_per_point_bytes : 1));
    pc = pc / 128 * 128;
    int64_t need = (P.total_pts + 127) / 128 * 128;
    if (pc > need) pc = need;
    if (pc < 128) pc = 128;
    return pc;
}

// Reverse-mode sweep: gradients of  sum(gy * y) + sum(gjets * jets)  w.r.t. the decoder weights / biases and the
// latent grid.  Per chunk of points: recompute the forward keeping every operand plane, then blend_backward ->
// (wgrad, dgrad) per hidden layer on the tensor cores; the per-vertex adjoint gVb is folded into the latent-column
// weights, the biases and the grid once at the end.
// points per chunk of the reverse-mode layout for a workspace of ws_bytes (0: too small)
static int64_t bwd_chunk_points(const Plan& P, size_t ws_bytes) {
    if (ws_bytes < P.bwd_fixed_bytes + bwd_chunk_region_bytes(P, 128)) return 0;
    int64_t pc = (int64_t)((ws_bytes - P.bwd_fixed_bytes - 64 * 1024) / P.bwd_per_point_bytes) / 128 * 128;
    int64_t need = (P.total_pts + 127) / 128 * 128;
    if (pc > need) pc = need;
    if (pc > (1 << 22)) pc = 1 << 22;
    return balanced_chunk_points(P.total_pts, pc);
}

enum { kBwdFull = 0, kBwdForwardOnly = 1, kBwdReuse = 2 };

// mode kBwdFull        : recompute the forward per chunk, then the reverse sweep
//      kBwdForwardOnly : training forward (stpde_jet_forward_train): the single chunk's planes stay in the workspace
//      kBwdReuse       : reverse sweep on the planes a kBwdForwardOnly call left in the SAME workspace
static int run_backward(int mode, const Plan& P, const stpde_desc_t* d, const float* grid, const float* q,
                        const float* const* W, const float* const* B, const float* gy, const float* gjets,
                        float* const* gW, float* const* gB, float* ggrid, float* gbeta, float* y, float* jets, char* ws,
                        size_t ws_bytes, int* status, cudaStream_t st) {
    const int dim = d->dim, kc = P.spec.kc, L = P.n_layers - 1;
    BufferList grads;               // decoder gradients: zeroed here, scaled by 1/S at the end, one launch each
    grads.count = 0;
    if (mode != kBwdForwardOnly) {
        for (int l = 0; l < P.n_layers; ++l) {
            grads.add(gW[l], (int64_t)P.widths[l] * P.in_features[l]);
            grads.add(gB[l], P.widths[l]);
        }
        grads.add(gbeta, 1);
        BufferList zero = grads;
        zero.add(ggrid, (int64_t)P.nvert_total * d->channels);
        launch_zero_buffers(zero, st);
    }
    if (P.total_pts == 0) return STPDE_OK;
    int64_t pc = bwd_chunk_points(P, ws_bytes);
    if (pc == 0)
        return fail(STPDE_ENOMEM, "workspace %zu B < minimum %zu B", ws_bytes, P.bwd_fixed_bytes + bwd_chunk_region_bytes(P, 128));
    if (mode != kBwdFull && pc < P.total_pts)
        return fail(STPDE_ENOMEM, "the training forward keeps ONE chunk: %lld points do not fit the workspace (%lld)",
                    (long long)P.total_pts, (long long)pc);
    const int64_t rows = pc * P.ncorner;
    if (rows * (int64_t)P.np64[0] * kc >= (int64_t)1 << 40 || rows >= ((int64_t)1 << 31))
        return fail(STPDE_EUNSUPPORTED, "chunk too large");

    char* p = ws + P.bwd_fixed_bytes;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
    ChunkBuffers cb;
    cb.pc = (int)pc;
    cb.rows = (int)rows;
    cb.vtx = (int*)take(rows * 4);
    cb.xrel = (float*)take((size_t)kMaxDim * rows * 4);
    cb.wfac = (float*)take((size_t)dim * 2 * pc * 4);
    cb.dfac = (float*)take((size_t)dim * 2 * pc * 4);
    cb.dxr = (float*)take((size_t)dim * pc * 4);
    const int np_last = P.np[P.n_layers - 2];
    float* act_last = (float*)take((size_t)kc * rows * np_last * 4);
    p = (char*)align_up((size_t)p, 1024);
    char* tc_chunk = p;

    NetDesc net;
    memset(&net, 0, sizeof(net));
    net.n_layers = P.n_layers;
    net.ncat = P.ncat;
    for (int l = 0; l < P.n_layers; ++l) {
        net.cat_off[l] = P.cat_off[l];
        net.in_features[l] = P.in_features[l];
        net.kh[l] = P.kh[l];
        net.W[l] = W[l];
        net.B[l] = B[l];
    }
    float* Vb = (float*)(ws + P.off_vb);
    float* g_vb = (float*)(ws + P.off_bwd_gvb);
    unsigned* maxes = (unsigned*)(ws + P.off_bwd_scale);
    float* scale = (float*)(ws + P.off_bwd_scale + 256);
    const float* Wx[kMaxLayers] = {nullptr};
    prof_begin(kSlotSetup, st);
    // kBwdReuse: the training forward left all of this in the workspace; a training forward whose caller vouches
    // (reserved[1] = 1) that the previous training forward in this workspace saw the same grid / weights skips it too -
    // a step that walks its batch in chunks splits the weights and builds the vertex table once, not once per chunk
    const bool prep = mode != kBwdReuse && !(mode == kBwdForwardOnly && d->reserved[1] == 1);
    for (int l = 0; l < P.n_layers; ++l) {
        float* wx = l < L ? (float*)(ws + P.off_wx[l]) : nullptr;
        Wx[l] = wx;
        if (!prep) continue;
        if (l == L) launch_pack_weights(W[l], P.widths[l], P.in_features[l], P.kh[l], dim, P.widths[l], P.kp[l], (float*)(ws + P.off_wh[l]), nullptr, st);
        else launch_pack_weights(W[l], P.widths[l], P.in_features[l], P.kh[l], dim, 0, 1, nullptr, wx, st);
    }
    if (prep) launch_vertex_bias(P.geom, P.nvert_total, net, grid, Vb, st);
    CUDA_TRY(cudaMemsetAsync(g_vb, 0, (size_t)P.nvert_total * P.ncat * sizeof(float), st));
    if (mode != kBwdForwardOnly) {
        // reserved[0] = headroom bits below the default adjoint scale (the binding retries with more headroom when
        // the range flag comes back)
        int target_exp = 10 - d->reserved[0];
        target_exp = target_exp > 14 ? 14 : (target_exp < -40 ? -40 : target_exp);
        launch_grad_scale(P.spec, P.geom, gy, kc > 1 ? gjets : nullptr, P.total_pts * P.O, maxes, scale, target_exp, st);
    }
    prof_end(kSlotSetup, st, P.n_layers + 4);

    TcBwdContext tc;
    // The saved pre-activations are fp16 planes when the forward that writes (wrote) them runs in the single-pass mode:
    // this call's own precision, or - reverse sweep on a stash (kBwdReuse) - what the caller reports in reserved[2].
    const bool z_half = tc_env().z_half && (mode == kBwdReuse ? d->reserved[2] == 1 : d->precision == STPDE_PREC_FP16);
    int rc = tc_bwd_prepare(tc, d->precision, P.n_layers, P.widths, P.in_features, W, ws + P.off_tc, tc_chunk,
                            (size_t)(ws + ws_bytes - tc_chunk), kc, (int)rows, status, prep, z_half, st);
    if (rc) return fail(rc, "%s", tc_last_error());

    const TcBwdLayer& TL = tc.layer[P.n_layers - 2];
    for (int64_t p0 = 0; p0 < P.total_pts; p0 += pc) {
        if (mode != kBwdReuse) {
            {
                ProfScope ps(kSlotPrep, st);
                launch_prep_points(P.geom, d->npts, P.total_pts, p0, cb, q, status, st);
            }
            // training forward: y / jets come out of the last hidden layer's epilogue when the shape allows it
            TcFinal fin;
            fin.w_last = (const float*)(ws + P.off_wh[L]);
            fin.b_last = B[L];
            fin.n_out = P.O;
            fin.ldw = P.kp[L];
            fin.y = y;
            fin.jets = jets;
            fin.p0 = p0;
            fin.total_pts = P.total_pts;
            const bool fuse = mode == kBwdForwardOnly && tc_bwd_can_fuse_final(tc, P.spec, dim, P.O);
            rc = tc_bwd_forward_chunk(tc, P.spec, dim, d->act_kind, d->act_param, cb, Vb, P.ncat, P.cat_off, Wx, act_last,
                                      np_last, fuse ? &fin : nullptr, st);
            if (rc) return fail(rc, "%s", tc_last_error());
            if (mode == kBwdForwardOnly) {
                if (!fuse) {
                    ProfScope ps(kSlotFinal, st);
                    launch_final_blend(P.spec, dim, cb.rows, cb.pc, P.total_pts, p0, P.kp[L], P.O, act_last,
                                       (const float*)(ws + P.off_wh[L]), B[L], cb, y, jets, st);
                }
                break;
            }
        }
        {
            BlendBwdArgs a;
            memset(&a, 0, sizeof(a));
            a.dim = dim; a.rows = cb.rows; a.O = P.O; a.Kp = P.kp[L]; a.n_feat = P.widths[L - 1];
            a.ld_out = TL.ldz; a.ldz = TL.ldz; a.act = d->act_kind; a.beta = d->act_param;
            a.ncat = P.ncat; a.cat_off = P.cat_off[L - 1]; a.three = tc.passes == 3;
            a.total_pts = P.total_pts; a.p0 = p0; a.cb = cb;
            a.gy = gy; a.gjets = gjets; a.scale = scale;
            a.Wlast = (const float*)(ws + P.off_wh[L]);
            a.act_last = act_last; a.z_in = TL.z; a.z_half = tc.z_half;
            a.out_hi = TL.zb[0]; a.out_lo = TL.zb[1];
            a.g_vb = g_vb;
            a.g_wx = gW[L - 1] + P.kh[L - 1]; a.g_wx_ld = P.in_features[L - 1];
            a.g_wlast = gW[L]; a.g_blast = gB[L]; a.g_beta = gbeta;
            a.status = status;
            ProfScope ps(kSlotBwdBlend, st);
            rc = launch_blend_backward(P.spec, a, st);
            if (rc) return fail(rc, "blend_backward: decoder too wide for the shared-memory staging");
        }
        rc = tc_bwd_backward_chunk(tc, P.spec, dim, d->act_kind, d->act_param, cb, Vb, P.ncat, P.cat_off, P.in_features, Wx,
                                   gW, g_vb, gbeta, st);
        if (rc) return fail(rc, "%s", tc_last_error());
    }
    if (mode != kBwdForwardOnly) {
        ProfScope ps(kSlotBwdVertex, st, 3);
        VertexBwdArgs v;
        memset(&v, 0, sizeof(v));
        v.n_layers = P.n_layers; v.ncat = P.ncat;
        for (int l = 0; l < P.n_layers; ++l) {
            v.cat_off[l] = P.cat_off[l]; v.in_features[l] = P.in_features[l]; v.kh[l] = P.kh[l];
            v.W[l] = W[l]; v.gW[l] = gW[l]; v.gB[l] = gB[l];
        }
        v.grid = grid; v.g_vb = g_vb; v.scale = scale;
        launch_vertex_backward(P.geom, P.nvert_total, v, ggrid, st);
        launch_scale_buffers(grads, scale, st);
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

}  // namespace stpde

using namespace stpde;

extern "C" {

int stpde_version(void) { return STPDE_VERSION; }

const char* stpde_last_error(void) { return g_err; }

size_t stpde_desc_size(void) { return sizeof(stpde_desc_t); }

int stpde_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

size_t stpde_workspace_bytes(const stpde_desc_t* desc) {
    Plan P;
    if (make_plan(P, desc, nullptr, nullptr) != STPDE_OK) return 0;
    return P.fixed_bytes + chunk_region_bytes(P, default_chunk_points(P));
}

int stpde_interp_coefficients(int32_t batch, int32_t npts, int32_t dim, const int32_t* grid_size, int32_t channels,
                              const float* grid, const int64_t* grid_strides, const float* q,
                              const int64_t* q_strides, const float* xmin, const float* xmax, float* corner_values,
                              float* weights, float* x_relative, int32_t* status, void* stream) {
    GridGeom g;
    int rc = make_geom(g, dim, grid_size, channels, xmin, xmax, grid_strides, q_strides);
    if (rc) return rc;
    launch_interp_coeff(g, batch, npts, grid, q, corner_values, weights, x_relative, status, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

int stpde_interp(int32_t batch, int32_t npts, int32_t dim, const int32_t* grid_size, int32_t channels,
                 const float* grid, const int64_t* grid_strides, const float* q, const int64_t* q_strides,
                 const float* xmin, const float* xmax, float* out, int32_t* status, void* stream) {
    GridGeom g;
    int rc = make_geom(g, dim, grid_size, channels, xmin, xmax, grid_strides, q_strides);
    if (rc) return rc;
    launch_interp(g, batch, npts, grid, q, out, status, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

int stpde_jet_forward(const stpde_desc_t* desc, const float* grid, const int64_t* grid_strides, const float* q,
                      const int64_t* q_strides, const float* const* W, const float* const* B, float* y, float* jets,
                      void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
    Plan P;
    int rc = make_plan(P, desc, grid_strides, q_strides);
    if (rc) return rc;
    const bool empty = P.total_pts == 0;      // empty batches come with null data pointers (torch.empty(0).data_ptr() == 0)
    if (!grid || !W || !B || !workspace || !status || (!empty && (!q || !y))) return fail(STPDE_EINVAL, "null pointer argument");
    if (!empty && P.spec.kc > 1 && !jets) return fail(STPDE_EINVAL, "jets buffer required when derivatives are requested");
    return run_forward(P, desc, grid, q, W, B, y, jets, (char*)workspace, workspace_bytes, status, (cudaStream_t)stream);
}

size_t stpde_backward_workspace_bytes(const stpde_desc_t* desc) {
    Plan P;
    if (make_plan(P, desc, nullptr, nullptr) != STPDE_OK) return 0;
    if (P.n_layers < 3) { fail(STPDE_EUNSUPPORTED, "the fused backward needs at least 3 linear layers"); return 0; }
    return P.bwd_fixed_bytes + bwd_chunk_region_bytes(P, default_bwd_chunk_points(P));
}

int64_t stpde_backward_chunk_points(const stpde_desc_t* desc, size_t workspace_bytes) {
    Plan P;
    if (make_plan(P, desc, nullptr, nullptr) != STPDE_OK || P.n_layers < 3) return 0;
    return bwd_chunk_points(P, workspace_bytes);
}

int stpde_jet_forward_train(const stpde_desc_t* desc, const float* grid, const int64_t* grid_strides, const float* q,
                            const int64_t* q_strides, const float* const* W, const float* const* B, float* y, float* jets,
                            void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
    Plan P;
    int rc = make_plan(P, desc, grid_strides, q_strides);
    if (rc) return rc;
    if (P.n_layers < 3) return fail(STPDE_EUNSUPPORTED, "the training forward needs at least 3 linear layers");
    const bool empty = P.total_pts == 0;
    if (!grid || !W || !B || !workspace || !status || (!empty && (!q || !y))) return fail(STPDE_EINVAL, "null pointer argument");
    if (!empty && P.spec.kc > 1 && !jets) return fail(STPDE_EINVAL, "jets buffer required when derivatives are requested");
    return run_backward(kBwdForwardOnly, P, desc, grid, q, W, B, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, y, jets,
                        (char*)workspace, workspace_bytes, status, (cudaStream_t)stream);
}

int stpde_jet_backward(const stpde_desc_t* desc, const float* grid, const int64_t* grid_strides, const float* q,
                       const int64_t* q_strides, const float* const* W, const float* const* B, const float* gy,
                       const float* gjets, float* const* gW, float* const* gB, float* ggrid, float* gbeta,
                       void* workspace, size_t workspace_bytes, int32_t reuse_forward, int32_t* status, void* stream) {
    Plan P;
    int rc = make_plan(P, desc, grid_strides, q_strides);
    if (rc) return rc;
    if (P.n_layers < 3) return fail(STPDE_EUNSUPPORTED, "the fused backward needs at least 3 linear layers");
    const bool empty = P.total_pts == 0;
    if (!grid || !W || !B || !gW || !gB || !workspace || !status || (!empty && (!q || !gy)))
        return fail(STPDE_EINVAL, "null pointer argument");
    if (!empty && P.spec.kc > 1 && !gjets) return fail(STPDE_EINVAL, "gjets required when derivatives were requested");
    for (int l = 0; l < P.n_layers; ++l)
        if (!W[l] || !B[l] || !gW[l] || !gB[l]) return fail(STPDE_EINVAL, "null weight / gradient pointer for layer %d", l);
    return run_backward(reuse_forward ? kBwdReuse : kBwdFull, P, desc, grid, q, W, B, gy, gjets, gW, gB, ggrid, gbeta,
                        nullptr, nullptr, (char*)workspace, workspace_bytes, status, (cudaStream_t)stream);
}

int stpde_jet_forward_host(const stpde_desc_t* desc, const float* grid, const float* q, const float* const* W,
                           const float* const* B, float* y, float* jets) {
    Plan P;
    int64_t gs[kMaxDim + 2], qs[3];
    if (!desc) return fail(STPDE_EINVAL, "null descriptor");
    {
        int64_t s = desc->channels;
        gs[desc->dim + 1] = 1;
        for (int k = desc->dim; k >= 1; --k) { gs[k] = s; s *= desc->grid_size[k - 1]; }
        gs[0] = s;
        qs[2] = 1; qs[1] = desc->dim; qs[0] = (int64_t)desc->npts * desc->dim;
    }
    int rc = make_plan(P, desc, gs, qs);
    if (rc) return rc;
    const size_t ws_bytes = P.fixed_bytes + chunk_region_bytes(P, default_chunk_points(P));
    const size_t grid_bytes = (size_t)desc->batch * gs[0] * sizeof(float);
    const size_t q_bytes = (size_t)P.total_pts * desc->dim * sizeof(float);
    const size_t y_bytes = (size_t)P.total_pts * P.O * sizeof(float);
    const int n_jet = P.spec.kc - 1;
    std::vector<void*> owned;
    auto dmalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) return nullptr;
        owned.push_back(p);
        return p;
    };
    auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
    cudaStream_t st = 0;
    float* d_grid = (float*)dmalloc(grid_bytes);
    float* d_q = (float*)dmalloc(q_bytes);
    float* d_y = (float*)dmalloc(y_bytes);
    float* d_j = (float*)dmalloc(y_bytes * (n_jet > 0 ? n_jet : 1));
    char* d_ws = (char*)dmalloc(ws_bytes);
    int* d_status = (int*)dmalloc(4);
    const float* dW[kMaxLayers];
    const float* dB[kMaxLayers];
    bool ok = d_grid && d_q && d_y && d_j && d_ws && d_status;
    for (int l = 0; ok && l < P.n_layers; ++l) {
        size_t wb = (size_t)P.widths[l] * P.in_features[l] * sizeof(float);
        float* w = (float*)dmalloc(wb);
        float* b = (float*)dmalloc(P.widths[l] * sizeof(float));
        ok = w && b;
        if (ok) {
            cudaMemcpyAsync(w, W[l], wb, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(b, B[l], P.widths[l] * sizeof(float), cudaMemcpyHostToDevice, st);
            dW[l] = w; dB[l] = b;
        }
    }
    if (!ok) { cleanup(); return fail(STPDE_ENOMEM, "cudaMalloc failed"); }
    cudaMemcpyAsync(d_grid, grid, grid_bytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_q, q, q_bytes, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_status, 0, 4, st);
    rc = run_forward(P, desc, d_grid, d_q, dW, dB, d_y, d_j, d_ws, ws_bytes, d_status, st);
    int h_status = 0;
    if (rc == STPDE_OK) {
        cudaMemcpyAsync(y, d_y, y_bytes, cudaMemcpyDeviceToHost, st);
        if (n_jet > 0 && jets) cudaMemcpyAsync(jets, d_j, y_bytes * n_jet, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&h_status, d_status, 4, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(STPDE_ECUDA, "stream sync: %s", cudaGetErrorString(e));
        else if (h_status & kStatusIndex) rc = fail(STPDE_EINDEX, "query point outside the latent grid (xmin != 0?)");
        else if (h_status & kStatusRange) rc = fail(STPDE_ERANGE, "activation left the fp16 range of the split-precision path");
    }
    cleanup();
    return rc;
}

// shared validation of a forward residual program (stack discipline is checked on the host so that the kernels cannot
// run out of their 16-entry stacks)
static int check_forward_program(const int32_t* prog, int32_t prog_words, int32_t n_consts, int32_t dim, int32_t out_features,
                                 int32_t n_jet, int32_t n_eq) {
    if (prog_words < 0 || prog_words > 640 || (prog_words & 1)) return fail(STPDE_EUNSUPPORTED, "program too long (%d words)", prog_words);
    if (n_consts < 0 || n_consts > 128) return fail(STPDE_EUNSUPPORTED, "too many constants (%d)", n_consts);
    int sp = 0, eq = 0;
    for (int w = 0; w < prog_words; w += 2) {
        int op = prog[w], arg = prog[w + 1];
        switch (op) {
            case 0: if (arg < 0 || arg >= n_consts) return fail(STPDE_EINVAL, "const index"); ++sp; break;
            case 1: if (arg < 0 || arg >= dim) return fail(STPDE_EINVAL, "q index"); ++sp; break;
            case 2: if (arg < 0 || arg >= out_features) return fail(STPDE_EINVAL, "y index"); ++sp; break;
            case 3: if (arg < 0 || arg >= n_jet * out_features) return fail(STPDE_EINVAL, "jet index"); ++sp; break;
            case 4: case 5: if (sp < 2) return fail(STPDE_EINVAL, "stack underflow"); --sp; break;
            case 6: case 7: if (sp < 1) return fail(STPDE_EINVAL, "stack underflow"); break;
            case 8: if (sp != 1) return fail(STPDE_EINVAL, "equation leaves %d values", sp); sp = 0; ++eq; break;
            default: return fail(STPDE_EINVAL, "opcode %d", op);
        }
        if (sp > 16) return fail(STPDE_EUNSUPPORTED, "expression too deep");
    }
    if (eq != n_eq || sp != 0) return fail(STPDE_EINVAL, "program has %d equations, expected %d", eq, n_eq);
    return STPDE_OK;
}

int stpde_residuals(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet, const float* q,
                    const int64_t* q_strides, const float* y, const float* jets, const int32_t* prog,
                    int32_t prog_words, const float* consts, int32_t n_consts, int32_t n_eq, float* residuals,
                    void* stream) {
    static thread_local ResidualProgram rp;
    int rc = check_forward_program(prog, prog_words, n_consts, dim, out_features, n_jet, n_eq);
    if (rc) return rc;
    rp.n_words = prog_words;
    memcpy(rp.words, prog, prog_words * sizeof(int32_t));
    memcpy(rp.consts, consts, n_consts * sizeof(float));
    {
        ProfScope ps(kSlotResidual, (cudaStream_t)stream);
        launch_residuals(rp, npts, (int64_t)batch * npts, dim, out_features, n_jet, q, q_strides, y, jets, residuals,
                         (cudaStream_t)stream);
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

int stpde_residuals_backward(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                             const float* q, const int64_t* q_strides, const float* y, const float* jets,
                             const int32_t* prog, int32_t prog_words, const float* consts, int32_t n_consts, int32_t n_eq,
                             const float* gres, float* gy, float* gjets, void* stream) {
    static thread_local ResidualProgramBig rp;
    if (prog_words < 0 || prog_words > 2048 || (prog_words & 1)) return fail(STPDE_EUNSUPPORTED, "adjoint program too long (%d words)", prog_words);
    if (n_consts < 0 || n_consts > 256) return fail(STPDE_EUNSUPPORTED, "too many constants (%d)", n_consts);
    if ((int64_t)batch * npts > 0 && (!gres || !gy || (n_jet > 0 && !gjets))) return fail(STPDE_EINVAL, "null pointer argument");
    const int n_out = out_features * (1 + n_jet);
    int sp = 0, out = 0;
    for (int w = 0; w < prog_words; w += 2) {
        int op = prog[w], arg = prog[w + 1];
        switch (op) {
            case 0: if (arg < 0 || arg >= n_consts) return fail(STPDE_EINVAL, "const index"); ++sp; break;
            case 1: if (arg < 0 || arg >= dim) return fail(STPDE_EINVAL, "q index"); ++sp; break;
            case 2: if (arg < 0 || arg >= out_features) return fail(STPDE_EINVAL, "y index"); ++sp; break;
            case 3: if (arg < 0 || arg >= n_jet * out_features) return fail(STPDE_EINVAL, "jet index"); ++sp; break;
            case 9: if (arg < 0 || arg >= n_eq) return fail(STPDE_EINVAL, "gres index"); ++sp; break;
            case 4: case 5: if (sp < 2) return fail(STPDE_EINVAL, "stack underflow"); --sp; break;
            case 6: case 7: if (sp < 1) return fail(STPDE_EINVAL, "stack underflow"); break;
            case 8: if (sp != 1) return fail(STPDE_EINVAL, "adjoint program leaves %d values", sp); sp = 0; ++out; break;
            default: return fail(STPDE_EINVAL, "opcode %d", op);
        }
        if (sp > 16) return fail(STPDE_EUNSUPPORTED, "expression too deep");
    }
    if (out != n_out || sp != 0) return fail(STPDE_EINVAL, "adjoint program has %d outputs, expected %d", out, n_out);
    rp.n_words = prog_words;
    memcpy(rp.words, prog, prog_words * sizeof(int32_t));
    memcpy(rp.consts, consts, n_consts * sizeof(float));
    {
        ProfScope ps(kSlotResidual, (cudaStream_t)stream);
        launch_residuals_backward(rp, npts, (int64_t)batch * npts, dim, out_features, n_jet, n_eq, q, q_strides, y, jets, gres,
                                  gy, gjets, (cudaStream_t)stream);
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}


int32_t stpde_residual_loss_blocks(int64_t total_points) { return residual_loss_blocks(total_points); }

int stpde_residual_loss(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet, const float* q,
                        const int64_t* q_strides, const float* y, const float* jets, const float* target,
                        const int32_t* prog, int32_t prog_words, const float* consts, int32_t n_consts, int32_t n_eq,
                        int32_t loss_kind, float* partial, void* stream) {
    static thread_local ResidualProgram rp;
    if (loss_kind < 0 || loss_kind > 2) return fail(STPDE_EINVAL, "loss kind %d", loss_kind);
    int rc = check_forward_program(prog, prog_words, n_consts, dim, out_features, n_jet, n_eq);
    if (rc) return rc;
    if (n_eq > 16) return fail(STPDE_EUNSUPPORTED, "the fused loss handles at most 16 equations");
    if (!partial) return fail(STPDE_EINVAL, "null pointer argument");
    rp.n_words = prog_words;
    memcpy(rp.words, prog, prog_words * sizeof(int32_t));
    memcpy(rp.consts, consts, n_consts * sizeof(float));
    {
        ProfScope ps(kSlotResidual, (cudaStream_t)stream);
        launch_residual_loss(rp, npts, (int64_t)batch * npts, out_features, n_eq, loss_kind, q, q_strides, y, jets, target,
                             partial, (cudaStream_t)stream);
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

int stpde_residual_loss_backward(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                                 const float* q, const int64_t* q_strides, const float* y, const float* jets,
                                 const float* target, const int32_t* prog, int32_t prog_words, const float* consts,
                                 int32_t n_consts, const int32_t* adj_prog, int32_t adj_words, const float* adj_consts,
                                 int32_t n_adj_consts, int32_t n_eq, int32_t loss_kind, const float* g_sums, float* gy,
                                 float* gjets, void* stream) {
    static thread_local ResidualProgram rp;
    static thread_local ResidualProgramBig ap;
    if (loss_kind < 0 || loss_kind > 2) return fail(STPDE_EINVAL, "loss kind %d", loss_kind);
    int rc = check_forward_program(prog, prog_words, n_consts, dim, out_features, n_jet, n_eq);
    if (rc) return rc;
    if (n_eq > 16) return fail(STPDE_EUNSUPPORTED, "the fused loss handles at most 16 equations");
    if (adj_words < 0 || adj_words > 2048 || (adj_words & 1)) return fail(STPDE_EUNSUPPORTED, "adjoint program too long (%d words)", adj_words);
    if (n_adj_consts < 0 || n_adj_consts > 256) return fail(STPDE_EUNSUPPORTED, "too many constants (%d)", n_adj_consts);
    if ((int64_t)batch * npts > 0 && (!g_sums || !gy || (n_jet > 0 && !gjets))) return fail(STPDE_EINVAL, "null pointer argument");
    const int n_out = out_features * (1 + n_jet);
    int sp = 0, out = 0;
    for (int w = 0; w < adj_words; w += 2) {
        int op = adj_prog[w], arg = adj_prog[w + 1];
        switch (op) {
            case 0: if (arg < 0 || arg >= n_adj_consts) return fail(STPDE_EINVAL, "const index"); ++sp; break;
            case 1: if (arg < 0 || arg >= dim) return fail(STPDE_EINVAL, "q index"); ++sp; break;
            case 2: if (arg < 0 || arg >= out_features) return fail(STPDE_EINVAL, "y index"); ++sp; break;
            case 3: if (arg < 0 || arg >= n_jet * out_features) return fail(STPDE_EINVAL, "jet index"); ++sp; break;
            case 9: if (arg < 0 || arg >= n_eq) return fail(STPDE_EINVAL, "gres index"); ++sp; break;
            case 4: case 5: if (sp < 2) return fail(STPDE_EINVAL, "stack underflow"); --sp; break;
            case 6: case 7: if (sp < 1) return fail(STPDE_EINVAL, "stack underflow"); break;
            case 8: if (sp != 1) return fail(STPDE_EINVAL, "adjoint program leaves %d values", sp); sp = 0; ++out; break;
            default: return fail(STPDE_EINVAL, "opcode %d", op);
        }
        if (sp > 16) return fail(STPDE_EUNSUPPORTED, "expression too deep");
    }
    if (out != n_out || sp != 0) return fail(STPDE_EINVAL, "adjoint program has %d outputs, expected %d", out, n_out);
    rp.n_words = prog_words;
    memcpy(rp.words, prog, prog_words * sizeof(int32_t));
    memcpy(rp.consts, consts, n_consts * sizeof(float));
    ap.n_words = adj_words;
    memcpy(ap.words, adj_prog, adj_words * sizeof(int32_t));
    memcpy(ap.consts, adj_consts, n_adj_consts * sizeof(float));
    {
        ProfScope ps(kSlotResidual, (cudaStream_t)stream);
        launch_residual_loss_backward(rp, ap, npts, (int64_t)batch * npts, out_features, n_jet, n_eq, loss_kind, q, q_strides,
                                      y, jets, target, g_sums, gy, gjets, (cudaStream_t)stream);
    }
    CUDA_TRY(cudaGetLastError());
    return STPDE_OK;
}

}  // extern "C"
