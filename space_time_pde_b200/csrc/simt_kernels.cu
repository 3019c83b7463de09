// FP32 (CUDA-core) kernels of the decode + jet path, and the grid-interpolation kernels.
//
// Pipeline for one chunk of query points (see DESIGN.md "Data layout"):
//   prep_points   : clip, cell lookup, per-corner relative coordinates and blend factors
//   vertex_bias   : Vb_l[v] = b_l + W_l[:, latent cols] . latent[v]   (once per latent grid)
//   layer0_jets   : layer-0 jets are closed-form (no contraction over activations)
//   layer_gemm    : z = Wh . a_prev for all jet components, fused jet-activation epilogue
//   final_blend   : last linear layer + multilinear blend with product rule -> y, jets
//
// All derivative components inside the MLP are taken w.r.t. the cell-local coordinate x_rel
// (tangent seeds are unit vectors); the 1/cubesize and clip-gradient factors are applied once in
// final_blend.
#include "common.cuh"
#include "kernels.h"

namespace stpde {

// ----------------------------------------------------------------------------------------------
// cell lookup shared by every kernel (reference regular_nd_grid_interpolation.py:47-52,69-70)
// ----------------------------------------------------------------------------------------------
struct Cell {
    float qc[kMaxDim];
    float clipgrad[kMaxDim];
    int ind0[kMaxDim];
    float xyz0[kMaxDim], xyz1[kMaxDim];
    bool bad;
};

__device__ __forceinline__ Cell cell_lookup(const GridGeom& g, const float* __restrict__ q, int64_t qoff) {
    Cell c;
    c.bad = false;
#pragma unroll
    for (int k = 0; k < kMaxDim; ++k) {
        if (k < g.dim) {
            float x = q[qoff + k * g.qstride[2]];
            float qmin = fminf(x, g.hi[k]);
            float qc = fmaxf(qmin, g.lo[k]);
            // torch.min / torch.max backward: 1 to the selected operand, 0.5 on exact ties
            float gmin = x < g.hi[k] ? 1.f : (x == g.hi[k] ? 0.5f : 0.f);
            float gmax = qmin > g.lo[k] ? 1.f : (qmin == g.lo[k] ? 0.5f : 0.f);
            if (x != x) { qc = x; }  // NaN propagates like torch.min/max
            c.qc[k] = qc;
            c.clipgrad[k] = gmin * gmax;
            float fl = floorf(__fdiv_rn(qc, g.cubesize[k]));
            int i0 = (fl >= -2.0e9f && fl <= 2.0e9f) ? (int)fl : INT32_MIN / 2;
            c.ind0[k] = i0;
            c.xyz0[k] = __fmul_rn((float)i0, g.cubesize[k]);
            c.xyz1[k] = __fmul_rn((float)i0 + 1.f, g.cubesize[k]);
        }
    }
    return c;
}

// python-style index: negative wraps once, anything else out of range is flagged and clamped
__device__ __forceinline__ int wrap_index(int i, int n, bool& bad) {
    if (i < 0) i += n;
    if (i < 0 || i >= n) { bad = true; i = min(max(i, 0), n - 1); }
    return i;
}

// ----------------------------------------------------------------------------------------------
// regular_nd_grid_interpolation_coefficients / regular_nd_grid_interpolation
// ----------------------------------------------------------------------------------------------
__global__ void interp_coeff_kernel(GridGeom g, int batch, int npts, const float* __restrict__ grid,
                                    const float* __restrict__ q, float* __restrict__ corner_values,
                                    float* __restrict__ weights, float* __restrict__ x_rel,
                                    int* __restrict__ status) {
    const int ncorner = 1 << g.dim;
    const int64_t total = (int64_t)batch * npts * ncorner * g.channels;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int ch = (int)(e % g.channels);
        int64_t r = e / g.channels;
        int j = (int)(r % ncorner);
        int64_t pt = r / ncorner;
        int b = (int)(pt / npts);
        int p = (int)(pt % npts);
        Cell c = cell_lookup(g, q, b * g.qstride[0] + p * g.qstride[1]);
        bool bad = false;
        int64_t off = b * g.gstride[0] + ch * g.gstride[g.dim + 1];
        float w = 1.f;
        for (int k = 0; k < g.dim; ++k) {
            int bit = (j >> (g.dim - 1 - k)) & 1;
            int idx = wrap_index(c.ind0[k] + bit, g.size[k], bad);
            off += idx * g.gstride[1 + k];
            float pos = bit ? c.xyz1[k] : c.xyz0[k];
            float opp = bit ? c.xyz0[k] : c.xyz1[k];
            float f = __fdiv_rn(fabsf(c.qc[k] - opp), g.cubesize[k]);
            w = (k == 0) ? f : __fmul_rn(w, f);
            if (ch == 0) x_rel[r * g.dim + k] = __fdiv_rn(c.qc[k] - pos, g.cubesize[k]);
        }
        corner_values[e] = grid[off];
        if (ch == 0) weights[r] = w;
        if (bad) atomicOr(status, kStatusIndex);
    }
}

__global__ void interp_kernel(GridGeom g, int batch, int npts, const float* __restrict__ grid,
                              const float* __restrict__ q, float* __restrict__ out, int* __restrict__ status) {
    const int ncorner = 1 << g.dim;
    const int64_t total = (int64_t)batch * npts * g.channels;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int ch = (int)(e % g.channels);
        int64_t pt = e / g.channels;
        int b = (int)(pt / npts);
        int p = (int)(pt % npts);
        Cell c = cell_lookup(g, q, b * g.qstride[0] + p * g.qstride[1]);
        bool bad = false;
        float acc = 0.f;
        for (int j = 0; j < ncorner; ++j) {
            int64_t off = b * g.gstride[0] + ch * g.gstride[g.dim + 1];
            float w = 1.f;
            for (int k = 0; k < g.dim; ++k) {
                int bit = (j >> (g.dim - 1 - k)) & 1;
                int idx = wrap_index(c.ind0[k] + bit, g.size[k], bad);
                off += idx * g.gstride[1 + k];
                float opp = bit ? c.xyz0[k] : c.xyz1[k];
                float f = __fdiv_rn(fabsf(c.qc[k] - opp), g.cubesize[k]);
                w = (k == 0) ? f : __fmul_rn(w, f);
            }
            acc = __fadd_rn(acc, __fmul_rn(grid[off], w));  // torch.sum over corners, in order
        }
        out[e] = acc;
        if (bad) atomicOr(status, kStatusIndex);
    }
}

// ----------------------------------------------------------------------------------------------
// prep_points: one thread per query point of the chunk
// ----------------------------------------------------------------------------------------------
__global__ void prep_points_kernel(GridGeom g, int npts, int64_t total_pts, int64_t p0, ChunkBuffers cb,
                                   const float* __restrict__ q, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cb.pc) return;
    const int ncorner = 1 << g.dim;
    const int64_t gp = p0 + i;
    if (gp >= total_pts) {  // padding point: neutral, finite
        for (int j = 0; j < ncorner; ++j) {
            cb.vtx[(int64_t)i * ncorner + j] = 0;
            for (int k = 0; k < kMaxDim; ++k) cb.xrel[(int64_t)k * cb.rows + (int64_t)i * ncorner + j] = 0.f;
        }
        for (int k = 0; k < g.dim; ++k) {
            cb.wfac[(k * 2 + 0) * cb.pc + i] = 0.f; cb.wfac[(k * 2 + 1) * cb.pc + i] = 0.f;
            cb.dfac[(k * 2 + 0) * cb.pc + i] = 0.f; cb.dfac[(k * 2 + 1) * cb.pc + i] = 0.f;
            cb.dxr[k * cb.pc + i] = 0.f;
        }
        return;
    }
    const int b = (int)(gp / npts);
    const int p = (int)(gp % npts);
    Cell c = cell_lookup(g, q, b * g.qstride[0] + p * g.qstride[1]);
    bool bad = false;
    int idx[kMaxDim][2];
    float xr[kMaxDim][2];
#pragma unroll
    for (int k = 0; k < kMaxDim; ++k) {
        if (k < g.dim) {
            idx[k][0] = wrap_index(c.ind0[k], g.size[k], bad);
            idx[k][1] = wrap_index(c.ind0[k] + 1, g.size[k], bad);
            float d0 = c.qc[k] - c.xyz0[k];
            float d1 = c.qc[k] - c.xyz1[k];
            xr[k][0] = __fdiv_rn(d0, g.cubesize[k]);  // relative coordinate w.r.t. the low corner
            xr[k][1] = __fdiv_rn(d1, g.cubesize[k]);  // ... w.r.t. the high corner
            // blend factor of a corner with bit b is |q - pos_opposite| / cubesize
            float s0 = (d1 > 0.f) ? 1.f : (d1 < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
            float s1 = (d0 > 0.f) ? 1.f : (d0 < 0.f ? -1.f : 0.f);
            float ginv = __fdiv_rn(c.clipgrad[k], g.cubesize[k]);
            cb.wfac[(k * 2 + 0) * cb.pc + i] = __fdiv_rn(fabsf(d1), g.cubesize[k]);
            cb.wfac[(k * 2 + 1) * cb.pc + i] = __fdiv_rn(fabsf(d0), g.cubesize[k]);
            cb.dfac[(k * 2 + 0) * cb.pc + i] = s0 * ginv;
            cb.dfac[(k * 2 + 1) * cb.pc + i] = s1 * ginv;
            cb.dxr[k * cb.pc + i] = ginv;
        }
    }
    for (int j = 0; j < ncorner; ++j) {
        int v = 0;
        for (int k = 0; k < g.dim; ++k) {
            int bit = (j >> (g.dim - 1 - k)) & 1;
            v = v * g.size[k] + idx[k][bit];
            cb.xrel[(int64_t)k * cb.rows + (int64_t)i * ncorner + j] = xr[k][bit];
        }
        for (int k = g.dim; k < kMaxDim; ++k) cb.xrel[(int64_t)k * cb.rows + (int64_t)i * ncorner + j] = 0.f;
        cb.vtx[(int64_t)i * ncorner + j] = b * g.nvert + v;
    }
    if (bad) atomicOr(status, kStatusIndex);
}

// ----------------------------------------------------------------------------------------------
// vertex_bias: Vb[v][cat] = b_l[n] + sum_ch W_l[n][koff_l + dim + ch] * latent[v][ch]
// block = 128 features x 16 vertices
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vertex_bias_kernel(GridGeom g, int nvert_total, NetDesc net,
                                                          const float* __restrict__ grid, float* __restrict__ Vb) {
    extern __shared__ float lat[];  // [16][channels]
    const int v0 = blockIdx.y * 16;
    const int c = g.channels;
    for (int e = threadIdx.x; e < 16 * c; e += blockDim.x) {
        int vi = e / c, ch = e % c;
        int v = v0 + vi;
        float val = 0.f;
        if (v < nvert_total) {
            int b = v / g.nvert;
            int rem = v % g.nvert;
            int64_t off = b * g.gstride[0] + ch * g.gstride[g.dim + 1];
            for (int k = g.dim - 1; k >= 0; --k) {
                off += (rem % g.size[k]) * g.gstride[1 + k];
                rem /= g.size[k];
            }
            val = grid[off];
        }
        lat[e] = val;
    }
    __syncthreads();
    const int cat = blockIdx.x * blockDim.x + threadIdx.x;
    if (cat >= net.ncat) return;
    int l = 0;
    while (l + 1 < net.n_layers - 1 && cat >= net.cat_off[l + 1]) ++l;
    const int n = cat - net.cat_off[l];
    const float* wrow = net.W[l] + (int64_t)n * net.in_features[l] + net.kh[l] + g.dim;
    float acc[16];
    const float bias = net.B[l][n];
#pragma unroll
    for (int vi = 0; vi < 16; ++vi) acc[vi] = bias;
    for (int ch = 0; ch < c; ++ch) {
        float w = wrow[ch];
#pragma unroll
        for (int vi = 0; vi < 16; ++vi) acc[vi] = fmaf(w, lat[vi * c + ch], acc[vi]);
    }
#pragma unroll
    for (int vi = 0; vi < 16; ++vi)
        if (v0 + vi < nvert_total) Vb[(int64_t)(v0 + vi) * net.ncat + cat] = acc[vi];
}

// ----------------------------------------------------------------------------------------------
// pack_weights: Wh_l [Np][Kp] zero padded (activation columns only), Wx_l [N][dim]
// ----------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ W, int N, int in_features, int kh, int dim,
                                    int Np, int Kp, float* __restrict__ Wh, float* __restrict__ Wx) {
    const int64_t total = (int64_t)Np * Kp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(e / Kp), k = (int)(e % Kp);
        Wh[e] = (n < N && k < kh) ? W[(int64_t)n * in_features + k] : 0.f;
    }
    if (Wx != nullptr) {
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)N * dim;
             e += (int64_t)gridDim.x * blockDim.x) {
            int n = (int)(e / dim), k = (int)(e % dim);
            Wx[e] = W[(int64_t)n * in_features + kh + k];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// layer 0: z0 = Vb0[vtx] + W0x . xrel ; tangents are the constant columns of W0x; curvature is 0
// out[c][r][n], n < Np (pad columns written as 0)
// ----------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(256) layer0_jets_kernel(JetSpec spec, int dim, int act, float beta, int rows,
                                                          int N, int Np, const int* __restrict__ vtx,
                                                          const float* __restrict__ xrel,
                                                          const float* __restrict__ Wx, const float* __restrict__ Vb,
                                                          int ncat, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Np) return;
    float wx[kMaxDim];
#pragma unroll
    for (int k = 0; k < kMaxDim; ++k) wx[k] = (k < dim && n < N) ? Wx[n * dim + k] : 0.f;
    const int r0 = blockIdx.y * 16;
    for (int r = r0; r < min(r0 + 16, rows); ++r) {
        float zc[KC];
        if (n < N) {
            float z = Vb[(int64_t)vtx[r] * ncat + n];
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k)
                if (k < dim) z = fmaf(wx[k], xrel[(int64_t)k * rows + r], z);
            float s0, s1, s2;
            act_jet(act, beta, z, s0, s1, s2);
            zc[0] = s0;
#pragma unroll
            for (int c = 1; c < KC; ++c) {
                if (c <= spec.n_first) zc[c] = s1 * wx[spec.first_dirs[c - 1]];
                else {
                    int s = c - 1 - spec.n_first;
                    zc[c] = s2 * wx[spec.first_dirs[spec.sec_a[s] - 1]] * wx[spec.first_dirs[spec.sec_b[s] - 1]];
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < KC; ++c) zc[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < KC; ++c) out[((int64_t)c * rows + r) * Np + n] = zc[c];
    }
}

// ----------------------------------------------------------------------------------------------
// layer_gemm: z[c][r][n] = sum_k a[c][r][k] * Wh[n][k], fused jet activation epilogue.
// CTA tile 64 rows x 64 features x KC components, BK = 16, 256 threads, 4x4xKC per thread,
// cp.async double buffering.  Rows/cols of a thread are strided by 16 (bank-conflict-free LDS.128).
// ----------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBN = 64, kBK = 16, kLd = 20;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int KC>
__global__ void __launch_bounds__(256) layer_gemm_kernel(JetSpec spec, int dim, int act, float beta, int rows,
                                                         int N, int Np, int Kp, int NpOut,
                                                         const float* __restrict__ actIn,   // [KC][rows][Kp]
                                                         const float* __restrict__ Wh,      // [Np64][Kp]
                                                         const float* __restrict__ Wx,      // [N][dim]
                                                         const float* __restrict__ Vb, int ncat, int cat_off,
                                                         const int* __restrict__ vtx, const float* __restrict__ xrel,
                                                         float* __restrict__ out) {        // [KC][rows][NpOut]
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                  // [2][KC][64][kLd]
    float* Bs = smem + 2 * KC * kBM * kLd;             // [2][64][kLd]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * kBN, r0 = blockIdx.y * kBM;

    // loader mapping: thread -> (row = tid / 4, 16B chunk = tid % 4)
    const int lrow = tid >> 2, lchunk = tid & 3;
    const bool arow_ok = (r0 + lrow) < rows;
    const float* a_src = actIn + ((int64_t)min(r0 + lrow, rows - 1)) * Kp + lchunk * 4;
    const float* b_src = Wh + ((int64_t)(n0 + lrow)) * Kp + lchunk * 4;  // Wh is padded to a multiple of 64 rows
    (void)arow_ok;

    auto load_tile = [&](int buf, int k0) {
#pragma unroll
        for (int c = 0; c < KC; ++c)
            cp_async16(As + ((buf * KC + c) * kBM + lrow) * kLd + lchunk * 4, a_src + (int64_t)c * rows * Kp + k0);
        cp_async16(Bs + (buf * kBN + lrow) * kLd + lchunk * 4, b_src + k0);
        cp_async_commit();
    };

    float acc[KC][4][4];
#pragma unroll
    for (int c = 0; c < KC; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.f;

    const int nk = Kp / kBK;
    load_tile(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { load_tile(buf ^ 1, (kt + 1) * kBK); cp_async_wait<1>(); }
        else { cp_async_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = *reinterpret_cast<const float4*>(Bs + (buf * kBN + tx + 16 * j) * kLd + kk * 4);
#pragma unroll
            for (int c = 0; c < KC; ++c) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 a = *reinterpret_cast<const float4*>(As + ((buf * KC + c) * kBM + ty + 16 * i) * kLd + kk * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[c][i][j] = fmaf(a.x, b[j].x, acc[c][i][j]);
                        acc[c][i][j] = fmaf(a.y, b[j].y, acc[c][i][j]);
                        acc[c][i][j] = fmaf(a.z, b[j].z, acc[c][i][j]);
                        acc[c][i][j] = fmaf(a.w, b[j].w, acc[c][i][j]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // epilogue: skip-connection (x_rel columns + per-vertex latent/bias term) and jet activation
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx + 16 * j;
        if (n >= NpOut) continue;
        float wx[kMaxDim];
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) wx[k] = (k < dim && n < N) ? Wx[n * dim + k] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 16 * i;
            if (r >= rows) continue;
            float o[KC];
            if (n < N) {
                float z = acc[0][i][j] + Vb[(int64_t)vtx[r] * ncat + cat_off + n];
#pragma unroll
                for (int k = 0; k < kMaxDim; ++k)
                    if (k < dim) z = fmaf(wx[k], xrel[(int64_t)k * rows + r], z);
                float s0, s1, s2;
                act_jet(act, beta, z, s0, s1, s2);
                float zt[KC];
                zt[0] = z;
#pragma unroll
                for (int c = 1; c < KC; ++c)
                    zt[c] = (c <= spec.n_first) ? acc[c][i][j] + wx[spec.first_dirs[c - 1]] : acc[c][i][j];
                o[0] = s0;
#pragma unroll
                for (int c = 1; c < KC; ++c) {
                    if (c <= spec.n_first) o[c] = s1 * zt[c];
                    else {
                        int s = c - 1 - spec.n_first;
                        float za = 0.f, zb = 0.f;
#pragma unroll
                        for (int cc = 1; cc < KC; ++cc) {  // static indexing keeps zt[] in registers
                            if (cc == spec.sec_a[s]) za = zt[cc];
                            if (cc == spec.sec_b[s]) zb = zt[cc];
                        }
                        o[c] = fmaf(s2 * za, zb, s1 * zt[c]);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < KC; ++c) o[c] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < KC; ++c) out[((int64_t)c * rows + r) * NpOut + n] = o[c];
        }
    }
}

// Last linear layer for the nvec = KC * 128 (component, local row) vectors of a CTA.  Each warp takes batches of
// RB = 32 / OP vectors; a lane accumulates the OP partial dot products over its float4 slices of K (coalesced
// 512 B loads), then a transpose-reduce over the warp (31 shuffles for 32 values) leaves lane L with the finished
// sum of (vector L / OP, output L % OP).
template <int OP>
__device__ __forceinline__ void final_gemv(int nvec, int row0, int rows, int Kp, int O, const float* __restrict__ actIn,
                                           const float* Ws, const float* __restrict__ blast, float* outc) {
    constexpr int RB = 32 / OP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b0 = warp * RB; b0 < nvec; b0 += 8 * RB) {
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int vec = b0 + j;
            const int c = vec / 128, lr = vec % 128;
            const int r = row0 + lr;
            if (vec < nvec && r < rows) {
                const float4* a4 = reinterpret_cast<const float4*>(actIn + ((int64_t)c * rows + r) * Kp);
                for (int q4 = lane; q4 < Kp / 4; q4 += 32) {
                    const float4 a = __ldg(a4 + q4);
#pragma unroll
                    for (int o = 0; o < OP; ++o) {
                        if (o < O) {
                            const float4 w = *reinterpret_cast<const float4*>(Ws + o * Kp + q4 * 4);
                            v[j * OP + o] += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j < sft) {
                    const bool up = (lane & sft) != 0;
                    const float send = up ? v[j] : v[j + sft];
                    const float keep = up ? v[j + sft] : v[j];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                }
            }
        }
        const int vec = b0 + lane / OP, o = lane % OP;
        if (vec < nvec && o < O) {
            const int c = vec / 128, lr = vec % 128;
            outc[(c * 128 + lr) * O + o] = v[0] + (c == 0 ? blast[o] : 0.f);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// final_blend: last linear layer (no activation) + blend over the 2^d corners with product rule.
// One CTA = 128 (point, corner) rows.
// ----------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(256) final_blend_kernel(JetSpec spec, int dim, int rows, int pc, int64_t total_pts,
                                                          int64_t p0, int Kp, int O,
                                                          const float* __restrict__ actIn,  // [KC][rows][Kp]
                                                          const float* __restrict__ Wlast,  // [O][Kp] (padded)
                                                          const float* __restrict__ blast,  // [O]
                                                          ChunkBuffers cb, float* __restrict__ y,
                                                          float* __restrict__ jets) {
    extern __shared__ float sm[];
    float* Ws = sm;                    // [O][Kp]
    float* outc = sm + O * Kp;         // [KC][128][O]
    const int ncorner = 1 << dim;
    const int row0 = blockIdx.x * 128;
    for (int e = threadIdx.x; e < O * Kp; e += blockDim.x) Ws[e] = Wlast[e];
    __syncthreads();
    switch (O <= 1 ? 1 : O <= 2 ? 2 : O <= 4 ? 4 : 8) {
        case 1: final_gemv<1>(KC * 128, row0, rows, Kp, O, actIn, Ws, blast, outc); break;
        case 2: final_gemv<2>(KC * 128, row0, rows, Kp, O, actIn, Ws, blast, outc); break;
        case 4: final_gemv<4>(KC * 128, row0, rows, Kp, O, actIn, Ws, blast, outc); break;
        default: final_gemv<8>(KC * 128, row0, rows, Kp, O, actIn, Ws, blast, outc); break;
    }
    __syncthreads();
    // blend over the 2^d corners with the product rule: one thread per (point, output, jet component)
    const int pts_per_cta = 128 / ncorner;
    for (int e = threadIdx.x; e < pts_per_cta * O * KC; e += blockDim.x) {
        const int c = e % KC, o = (e / KC) % O, lp = e / (KC * O);
        const int i = row0 / ncorner + lp;  // point index within the chunk
        const int64_t gp = p0 + i;
        if (i >= pc || gp >= total_pts) continue;
        float f[kMaxDim][2], df[kMaxDim][2], dx[kMaxDim];
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) {
            f[k][0] = f[k][1] = 1.f; df[k][0] = df[k][1] = 0.f; dx[k] = 0.f;
            if (k < dim) {
                f[k][0] = cb.wfac[(k * 2 + 0) * cb.pc + i]; f[k][1] = cb.wfac[(k * 2 + 1) * cb.pc + i];
                df[k][0] = cb.dfac[(k * 2 + 0) * cb.pc + i]; df[k][1] = cb.dfac[(k * 2 + 1) * cb.pc + i];
                dx[k] = cb.dxr[k * cb.pc + i];
            }
        }
        const int kind = spec.kind[c];
        const int ca = kind == 2 ? spec.pa[c] : c, cbi = kind == 2 ? spec.pb[c] : c;   // parent components
        const int a = kind == 1 ? spec.dir[c] : spec.dir[ca], b = spec.dir[cbi];      // directions
        float dxa = 0.f, dxb = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) { if (k == a) dxa = dx[k]; if (k == b) dxb = dx[k]; }
        float res = 0.f;
        for (int j = 0; j < ncorner; ++j) {
            const int lr = lp * ncorner + j;
            const float o0 = outc[lr * O + o];                       // value component of this corner
            const float oc = outc[(c * 128 + lr) * O + o];
            float w = 1.f, wa = 1.f, wb = 1.f, wab = 1.f;            // weight and its derivatives along a, b, (a,b)
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) {
                if (k < dim) {
                    const int bit = (j >> (dim - 1 - k)) & 1;
                    const float fk = f[k][bit], dk = df[k][bit];
                    w = (k == 0) ? fk : __fmul_rn(w, fk);            // torch.prod order
                    wa *= (k == a) ? dk : fk;
                    wb *= (k == b) ? dk : fk;
                    wab *= (k == a || k == b) ? dk : fk;
                }
            }
            if (kind == 0) {
                res = __fadd_rn(res, __fmul_rn(o0, w));              // torch.sum(output * weights) in corner order
            } else if (kind == 1) {
                res += wa * o0 + w * dxa * oc;
            } else {
                const float oa = outc[(ca * 128 + lr) * O + o], ob = outc[(cbi * 128 + lr) * O + o];
                res += (a == b ? 0.f : wab) * o0 + wa * dxb * ob + wb * dxa * oa + w * dxa * dxb * oc;
            }
        }
        if (c == 0) y[gp * O + o] = res;
        else jets[((int64_t)(c - 1) * total_pts + gp) * O + o] = res;
    }
}

// ----------------------------------------------------------------------------------------------
// residual programs (postfix) evaluated per point
// ----------------------------------------------------------------------------------------------
__global__ void residual_kernel(ResidualProgram prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                                const float* __restrict__ q, int64_t qs0, int64_t qs1, int64_t qs2,
                                const float* __restrict__ y, const float* __restrict__ jets,
                                float* __restrict__ residuals) {
    for (int64_t gp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gp < total_pts;
         gp += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(gp / npts), p = (int)(gp % npts);
        float st[16];
        int sp = 0, eq = 0;
        for (int w = 0; w < prog.n_words; w += 2) {
            const int op = prog.words[w], arg = prog.words[w + 1];
            switch (op) {
                case 0: st[sp++] = prog.consts[arg]; break;
                case 1: st[sp++] = q[b * qs0 + p * qs1 + arg * qs2]; break;
                case 2: st[sp++] = y[gp * O + arg]; break;
                case 3: st[sp++] = jets[((int64_t)(arg / O) * total_pts + gp) * O + (arg % O)]; break;
                case 4: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
                case 5: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
                case 6: st[sp - 1] = -st[sp - 1]; break;
                case 7: {
                    float base = st[sp - 1], r = 1.f;
                    int n = arg < 0 ? -arg : arg;
                    for (int t = 0; t < n; ++t) r *= base;
                    st[sp - 1] = arg < 0 ? 1.f / r : r;
                    break;
                }
                default:  // END
                    residuals[(int64_t)eq * total_pts + gp] = st[0];
                    sp = 0; ++eq;
                    break;
            }
        }
    }
}

// Reverse mode of the residual programs: the host differentiates the equations symbolically and sends ONE postfix
// program per output symbol s (values y_0..y_{o-1}, then every jet entry), each evaluating
//     sum_e gres[e] * d residual_e / d s        (opcode 9 = PUSH_GRES e)
// so the kernel is the same interpreter with a different store: gy [b*p][O], gjets [n_jet][b*p][O].
__global__ void residual_backward_kernel(ResidualProgramBig prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                                         int n_eq, const float* __restrict__ q, int64_t qs0, int64_t qs1, int64_t qs2,
                                         const float* __restrict__ y, const float* __restrict__ jets,
                                         const float* __restrict__ gres, float* __restrict__ gy,
                                         float* __restrict__ gjets) {
    for (int64_t gp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gp < total_pts;
         gp += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(gp / npts), p = (int)(gp % npts);
        float st[16];
        int sp = 0, out = 0;
        for (int w = 0; w < prog.n_words; w += 2) {
            const int op = prog.words[w], arg = prog.words[w + 1];
            switch (op) {
                case 0: st[sp++] = prog.consts[arg]; break;
                case 1: st[sp++] = q[b * qs0 + p * qs1 + arg * qs2]; break;
                case 2: st[sp++] = y[gp * O + arg]; break;
                case 3: st[sp++] = jets[((int64_t)(arg / O) * total_pts + gp) * O + (arg % O)]; break;
                case 4: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
                case 5: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
                case 6: st[sp - 1] = -st[sp - 1]; break;
                case 7: {
                    float base = st[sp - 1], r = 1.f;
                    int n = arg < 0 ? -arg : arg;
                    for (int t = 0; t < n; ++t) r *= base;
                    st[sp - 1] = arg < 0 ? 1.f / r : r;
                    break;
                }
                case 9: st[sp++] = gres[(int64_t)arg * total_pts + gp]; break;
                default:  // END of the program of output symbol `out`
                    if (out < O) gy[gp * O + out] = st[0];
                    else gjets[((int64_t)((out - O) / O) * total_pts + gp) * O + ((out - O) % O)] = st[0];
                    sp = 0; ++out;
                    break;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Fused residual + loss reduction (reference experiments/rb2d/train.py:70-75):
//   reg = sum over (point, output) of l(y - target),   pde = sum over (equation, point) of l(residual)
// with l = |d| (l1), d^2 (l2) or smooth-l1 with beta 1 (huber).  The residuals never reach memory: every CTA emits one
// (reg, pde) pair of partial sums; the caller adds the <= 4736 pairs (and divides by the counts for the mean losses).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float loss_value(int kind, float d) {
    const float a = fabsf(d);
    return kind == 0 ? a : kind == 1 ? d * d : (a < 1.f ? 0.5f * d * d : a - 0.5f);
}
__device__ __forceinline__ float loss_grad(int kind, float d) {
    const float sg = d > 0.f ? 1.f : d < 0.f ? -1.f : 0.f;            // torch: sign(0) = 0
    return kind == 0 ? sg : kind == 1 ? 2.f * d : (fabsf(d) < 1.f ? d : sg);
}

// postfix interpreter shared by the fused-loss kernels: evaluates every equation of `prog` at point gp into res[]
template <class Prog>
__device__ __forceinline__ int eval_program(const Prog& prog, int b, int p, int64_t gp, int64_t total_pts, int O,
                                            const float* __restrict__ q, int64_t qs0, int64_t qs1, int64_t qs2,
                                            const float* __restrict__ y, const float* __restrict__ jets,
                                            const float* cot, float* res) {
    float st[16];
    int sp = 0, eq = 0;
    for (int w = 0; w < prog.n_words; w += 2) {
        const int op = prog.words[w], arg = prog.words[w + 1];
        switch (op) {
            case 0: st[sp++] = prog.consts[arg]; break;
            case 1: st[sp++] = q[b * qs0 + p * qs1 + arg * qs2]; break;
            case 2: st[sp++] = y[gp * O + arg]; break;
            case 3: st[sp++] = jets[((int64_t)(arg / O) * total_pts + gp) * O + (arg % O)]; break;
            case 4: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
            case 5: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
            case 6: st[sp - 1] = -st[sp - 1]; break;
            case 7: {
                float base = st[sp - 1], r = 1.f;
                int n = arg < 0 ? -arg : arg;
                for (int t = 0; t < n; ++t) r *= base;
                st[sp - 1] = arg < 0 ? 1.f / r : r;
                break;
            }
            case 9: st[sp++] = cot[arg]; break;           // adjoint programs: cotangent of residual `arg`
            default:  // END
                res[eq] = st[0];
                sp = 0; ++eq;
                break;
        }
    }
    return eq;
}

constexpr int kMaxEq = 16;

__global__ void __launch_bounds__(256) residual_loss_kernel(ResidualProgram prog, int npts, int64_t total_pts, int O, int n_eq,
                                                            int loss_kind, const float* __restrict__ q, int64_t qs0,
                                                            int64_t qs1, int64_t qs2, const float* __restrict__ y,
                                                            const float* __restrict__ jets,
                                                            const float* __restrict__ target, float* __restrict__ partial) {
    float reg = 0.f, pde = 0.f;
    for (int64_t gp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gp < total_pts;
         gp += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(gp / npts), p = (int)(gp % npts);
        float res[kMaxEq];
        eval_program(prog, b, p, gp, total_pts, O, q, qs0, qs1, qs2, y, jets, nullptr, res);
        for (int e = 0; e < n_eq; ++e) pde += loss_value(loss_kind, res[e]);
        for (int o = 0; o < O; ++o) reg += loss_value(loss_kind, y[gp * O + o] - (target ? target[gp * O + o] : 0.f));
    }
    __shared__ float red[2][8];
    for (int off = 16; off > 0; off >>= 1) {
        reg += __shfl_xor_sync(0xffffffffu, reg, off);
        pde += __shfl_xor_sync(0xffffffffu, pde, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = reg; red[1][threadIdx.x >> 5] = pde; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        partial[blockIdx.x * 2 + threadIdx.x] = t;
    }
}

// Reverse mode of the fused loss: the residuals are recomputed, their cotangents g_pde * l'(r_e) stay in registers and
// feed the adjoint programs (one per output symbol, as residual_backward_kernel); gy also receives g_reg * l'(y - target).
__global__ void __launch_bounds__(256) residual_loss_backward_kernel(ResidualProgram fwd, ResidualProgramBig adj, int npts,
                                                                     int64_t total_pts, int O, int n_jet, int n_eq, int loss_kind,
                                                                     const float* __restrict__ q, int64_t qs0, int64_t qs1,
                                                                     int64_t qs2, const float* __restrict__ y,
                                                                     const float* __restrict__ jets,
                                                                     const float* __restrict__ target,
                                                                     const float* __restrict__ g_sums, float* __restrict__ gy,
                                                                     float* __restrict__ gjets) {
    const float g_reg = g_sums[0], g_pde = g_sums[1];
    for (int64_t gp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gp < total_pts;
         gp += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(gp / npts), p = (int)(gp % npts);
        float cot[kMaxEq];
        eval_program(fwd, b, p, gp, total_pts, O, q, qs0, qs1, qs2, y, jets, nullptr, cot);
        for (int e = 0; e < n_eq; ++e) cot[e] = g_pde * loss_grad(loss_kind, cot[e]);
        // adjoint programs: output symbols y_0..y_{O-1}, then the jet entries (plane-major)
        float st[16];
        int sp = 0, out = 0;
        for (int w = 0; w < adj.n_words; w += 2) {
            const int op = adj.words[w], arg = adj.words[w + 1];
            switch (op) {
                case 0: st[sp++] = adj.consts[arg]; break;
                case 1: st[sp++] = q[b * qs0 + p * qs1 + arg * qs2]; break;
                case 2: st[sp++] = y[gp * O + arg]; break;
                case 3: st[sp++] = jets[((int64_t)(arg / O) * total_pts + gp) * O + (arg % O)]; break;
                case 4: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
                case 5: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
                case 6: st[sp - 1] = -st[sp - 1]; break;
                case 7: {
                    float base = st[sp - 1], r = 1.f;
                    int n = arg < 0 ? -arg : arg;
                    for (int t = 0; t < n; ++t) r *= base;
                    st[sp - 1] = arg < 0 ? 1.f / r : r;
                    break;
                }
                case 9: st[sp++] = cot[arg]; break;
                default:
                    if (out < O) {
                        const float d = y[gp * O + out] - (target ? target[gp * O + out] : 0.f);
                        gy[gp * O + out] = st[0] + g_reg * loss_grad(loss_kind, d);
                    } else {
                        gjets[((int64_t)((out - O) / O) * total_pts + gp) * O + ((out - O) % O)] = st[0];
                    }
                    sp = 0; ++out;
                    break;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host-side launchers
// ----------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 32) g = 148 * 32;  // grid-stride loops: a few waves of the 148 SMs
    return (int)(g < 1 ? 1 : g);
}

void launch_interp_coeff(const GridGeom& g, int batch, int npts, const float* grid, const float* q,
                         float* cv, float* w, float* xr, int* status, cudaStream_t st) {
    int64_t total = (int64_t)batch * npts * (1 << g.dim) * g.channels;
    if (total == 0) return;
    interp_coeff_kernel<<<grid_for(total, 256), 256, 0, st>>>(g, batch, npts, grid, q, cv, w, xr, status);
}

void launch_interp(const GridGeom& g, int batch, int npts, const float* grid, const float* q, float* out,
                   int* status, cudaStream_t st) {
    int64_t total = (int64_t)batch * npts * g.channels;
    if (total == 0) return;
    interp_kernel<<<grid_for(total, 256), 256, 0, st>>>(g, batch, npts, grid, q, out, status);
}

void launch_prep_points(const GridGeom& g, int npts, int64_t total_pts, int64_t p0, const ChunkBuffers& cb,
                        const float* q, int* status, cudaStream_t st) {
    prep_points_kernel<<<(cb.pc + 127) / 128, 128, 0, st>>>(g, npts, total_pts, p0, cb, q, status);
}

void launch_vertex_bias(const GridGeom& g, int nvert_total, const NetDesc& net, const float* grid, float* Vb,
                        cudaStream_t st) {
    dim3 grid_dim((net.ncat + 127) / 128, (nvert_total + 15) / 16);
    vertex_bias_kernel<<<grid_dim, 128, 16 * g.channels * sizeof(float), st>>>(g, nvert_total, net, grid, Vb);
}

void launch_pack_weights(const float* W, int N, int in_features, int kh, int dim, int Np, int Kp, float* Wh,
                         float* Wx, cudaStream_t st) {
    int64_t total = (int64_t)Np * Kp;
    if (total < (int64_t)N * dim) total = (int64_t)N * dim;  // layer 0 has no activation columns, only Wx
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(W, N, in_features, kh, dim, Np, Kp, Wh, Wx);
}

template <int KC>
static void launch_layer0_t(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np,
                            const int* vtx, const float* xrel, const float* Wx, const float* Vb, int ncat,
                            float* out, cudaStream_t st) {
    dim3 grid_dim((Np + 255) / 256, (rows + 15) / 16);
    layer0_jets_kernel<KC><<<grid_dim, 256, 0, st>>>(spec, dim, act, beta, rows, N, Np, vtx, xrel, Wx, Vb, ncat, out);
}

template <int KC>
static void launch_gemm_t(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np, int Kp,
                          int NpOut, const float* actIn, const float* Wh, const float* Wx, const float* Vb,
                          int ncat, int cat_off, const int* vtx, const float* xrel, float* out, cudaStream_t st) {
    size_t smem = (size_t)(2 * KC * kBM * kLd + 2 * kBN * kLd) * sizeof(float);
    static DeviceOnce configured;
    if (configured.first_use()) {
        cudaFuncSetAttribute(layer_gemm_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured.mark();
    }
    dim3 grid_dim((NpOut + kBN - 1) / kBN, (rows + kBM - 1) / kBM);
    layer_gemm_kernel<KC><<<grid_dim, 256, smem, st>>>(spec, dim, act, beta, rows, N, Np, Kp, NpOut, actIn, Wh, Wx,
                                                       Vb, ncat, cat_off, vtx, xrel, out);
}

template <int KC>
static void launch_final_t(const JetSpec& spec, int dim, int rows, int pc, int64_t total_pts, int64_t p0, int Kp,
                           int O, const float* actIn, const float* Wlast, const float* blast,
                           const ChunkBuffers& cb, float* y, float* jets, cudaStream_t st) {
    size_t smem = (size_t)(O * Kp + KC * 128 * O) * sizeof(float);
    static DeviceOnce configured;
    if (configured.first_use()) {
        cudaFuncSetAttribute(final_blend_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.mark();
    }
    final_blend_kernel<KC><<<(rows + 127) / 128, 256, smem, st>>>(spec, dim, rows, pc, total_pts, p0, Kp, O, actIn,
                                                                  Wlast, blast, cb, y, jets);
}

#define STPDE_DISPATCH_KC(kc, CALL)                                                            \
    switch (kc) {                                                                              \
        case 1: { constexpr int KC = 1; CALL; } break;                                         \
        case 2: { constexpr int KC = 2; CALL; } break;                                         \
        case 3: { constexpr int KC = 3; CALL; } break;                                         \
        case 4: { constexpr int KC = 4; CALL; } break;                                         \
        case 5: { constexpr int KC = 5; CALL; } break;                                         \
        case 6: { constexpr int KC = 6; CALL; } break;                                         \
        case 7: { constexpr int KC = 7; CALL; } break;                                         \
        case 8: { constexpr int KC = 8; CALL; } break;                                         \
        case 9: { constexpr int KC = 9; CALL; } break;                                         \
        default: { constexpr int KC = 10; CALL; } break;                                       \
    }

void launch_layer0(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np, const int* vtx,
                   const float* xrel, const float* Wx, const float* Vb, int ncat, float* out, cudaStream_t st) {
    STPDE_DISPATCH_KC(spec.kc, launch_layer0_t<KC>(spec, dim, act, beta, rows, N, Np, vtx, xrel, Wx, Vb, ncat, out, st));
}

void launch_layer_gemm(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np, int Kp, int NpOut,
                       const float* actIn, const float* Wh, const float* Wx, const float* Vb, int ncat, int cat_off,
                       const int* vtx, const float* xrel, float* out, cudaStream_t st) {
    STPDE_DISPATCH_KC(spec.kc, launch_gemm_t<KC>(spec, dim, act, beta, rows, N, Np, Kp, NpOut, actIn, Wh, Wx, Vb, ncat,
                                                 cat_off, vtx, xrel, out, st));
}

void launch_final_blend(const JetSpec& spec, int dim, int rows, int pc, int64_t total_pts, int64_t p0, int Kp, int O,
                        const float* actIn, const float* Wlast, const float* blast, const ChunkBuffers& cb, float* y,
                        float* jets, cudaStream_t st) {
    STPDE_DISPATCH_KC(spec.kc, launch_final_t<KC>(spec, dim, rows, pc, total_pts, p0, Kp, O, actIn, Wlast, blast, cb, y,
                                                  jets, st));
}

void launch_residuals_backward(const ResidualProgramBig& prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                               int n_eq, const float* q, const int64_t* qs, const float* y, const float* jets,
                               const float* gres, float* gy, float* gjets, cudaStream_t st) {
    if (total_pts == 0) return;
    residual_backward_kernel<<<grid_for(total_pts, 256), 256, 0, st>>>(prog, npts, total_pts, dim, O, n_jet, n_eq, q, qs[0],
                                                                        qs[1], qs[2], y, jets, gres, gy, gjets);
}

int residual_loss_blocks(int64_t total_pts) { return grid_for(total_pts, 256); }

void launch_residual_loss(const ResidualProgram& prog, int npts, int64_t total_pts, int O, int n_eq, int loss_kind,
                          const float* q, const int64_t* qs, const float* y, const float* jets, const float* target,
                          float* partial, cudaStream_t st) {
    residual_loss_kernel<<<residual_loss_blocks(total_pts), 256, 0, st>>>(prog, npts, total_pts, O, n_eq, loss_kind, q, qs[0],
                                                                          qs[1], qs[2], y, jets, target, partial);
}

void launch_residual_loss_backward(const ResidualProgram& fwd, const ResidualProgramBig& adj, int npts, int64_t total_pts,
                                   int O, int n_jet, int n_eq, int loss_kind, const float* q, const int64_t* qs,
                                   const float* y, const float* jets, const float* target, const float* g_sums, float* gy,
                                   float* gjets, cudaStream_t st) {
    if (total_pts == 0) return;
    residual_loss_backward_kernel<<<grid_for(total_pts, 256), 256, 0, st>>>(fwd, adj, npts, total_pts, O, n_jet, n_eq, loss_kind,
                                                                            q, qs[0], qs[1], qs[2], y, jets, target, g_sums,
                                                                            gy, gjets);
}

void launch_residuals(const ResidualProgram& prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                      const float* q, const int64_t* qs, const float* y, const float* jets, float* residuals,
                      cudaStream_t st) {
    if (total_pts == 0) return;
    residual_kernel<<<grid_for(total_pts, 256), 256, 0, st>>>(prog, npts, total_pts, dim, O, n_jet, q, qs[0], qs[1],
                                                               qs[2], y, jets, residuals);
}

}  // namespace stpde
