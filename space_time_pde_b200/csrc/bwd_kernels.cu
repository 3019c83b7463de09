// CUDA-core kernels of the reverse-mode (training) sweep: everything that is not a big contraction.
//
//   grad_scale      : power-of-two scale S for the adjoints (they travel through fp16 hi/lo operand planes)
//   blend_backward  : transpose of final_blend (product rule over the 2^d corners) + last linear layer
//                     + reverse jet activation of the last hidden layer
//   vertex_backward : adjoint of the per-vertex precompute Vb = b + W[:, latent cols] . latent
//   buffer_list     : zero fill / final 1/S of all gradient buffers in one launch
// Mirrors the forward kernels of simt_kernels.cu; see DESIGN.md "Reverse mode".
#include <cuda_fp16.h>

#include "bwd_kernels.h"
#include "common.cuh"
#include "kernels.h"

namespace stpde {

// ----------------------------------------------------------------------------------------------
// adjoint scale
// ----------------------------------------------------------------------------------------------
// maxes[plane] = max |g| over plane 0 = gy, plane 1.. = gjets planes (non-negative floats order like uints)
__global__ void absmax_planes_kernel(const float* __restrict__ gy, const float* __restrict__ gjets, int64_t plane_elems,
                                     int n_planes, unsigned* __restrict__ maxes) {
    const int pl = blockIdx.y;
    if (pl >= n_planes) return;
    const float* src = pl == 0 ? gy : gjets + (int64_t)(pl - 1) * plane_elems;
    float m = 0.f;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < plane_elems; e += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(src[e]));
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(maxes + pl, __float_as_uint(m));
}

// Bound on |d loss / d out_c| of any corner: the blend multiplies the incoming adjoints by weights <= 1 and by
// d x_rel / d q = 1 / cubesize per derivative order.  scale[0] = S = 2^(target_exp - ceil(log2 bound)), scale[1] = 1 / S.
// The adjoints live in fp16 hi/lo planes (|x| < 65504, lo goes subnormal below 2^-14): a large target keeps small
// adjoints precise, the headroom 2^(16 - target_exp) absorbs their growth through the transposed weights.
__global__ void grad_scale_kernel(JetSpec spec, GridGeom g, const unsigned* __restrict__ maxes, float* __restrict__ scale,
                                  int target_exp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float bound = __uint_as_float(maxes[0]);
    for (int c = 1; c < spec.kc; ++c) {
        float m = __uint_as_float(maxes[c]);
        if (spec.kind[c] == 1) m *= 1.f / g.cubesize[spec.dir[c]];
        else m *= (1.f / g.cubesize[spec.dir[spec.pa[c]]]) * (1.f / g.cubesize[spec.dir[spec.pb[c]]]);
        bound += m;
    }
    float S = 1.f;
    if (bound > 0.f && bound < 3.0e38f) {
        int e2 = 0;
        frexpf(bound, &e2);             // bound = m * 2^e2, m in [0.5, 1)
        int sh = target_exp - e2;
        sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
        S = ldexpf(1.f, sh);
    }
    scale[0] = S;
    scale[1] = 1.f / S;
}

// one launch for every gradient buffer of a call (blockIdx.y = buffer): zero fill before the sweep, final 1/S after it
// (the reference-size training step is launch-bound: 2 * n_layers + 2 separate memsets / scale kernels otherwise)
template <bool SCALE>
__global__ void buffer_list_kernel(const __grid_constant__ BufferList bl, const float* __restrict__ scale) {
    float* __restrict__ buf = bl.ptr[blockIdx.y];
    const int64_t n = bl.n[blockIdx.y];
    const float s = SCALE ? scale[1] : 0.f;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        buf[e] = SCALE ? buf[e] * s : 0.f;
}

// W^T hi/lo [fp rows (f)][pack * ldz (g)] = fp16 split of W[g][f] * 2^sw (the forward's per-layer scale, from its absmax).
// pack > 1 (narrow dgrad, tc::LayerArgs::pack): block diagonal - copy q sits in rows q * 128/pack ..., columns q * ldz ...
__global__ void split_weights_t_kernel(const float* __restrict__ W, int N, int in_features, int kh, int fp, int ldz, int pack,
                                       const unsigned* __restrict__ absmax, __half* __restrict__ hi, __half* __restrict__ lo) {
    const float amax = __uint_as_float(*absmax);
    int e2 = 0;
    if (amax > 0.f) frexpf(amax, &e2);
    const int sw = amax > 0.f ? 14 - e2 : 0;
    const float up = ldexpf(1.f, sw);
    const int ldk = pack * ldz, lanes = pack > 1 ? 128 / pack : fp;
    const int64_t total = (int64_t)fp * ldk;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / ldk), col = (int)(e % ldk);
        const int q = col / ldz, g = col - q * ldz, f = row - q * lanes;
        const float x = (f >= 0 && f < lanes && g < N && f < kh) ? W[(int64_t)g * in_features + f] * up : 0.f;
        const __half h = __float2half_rn(x);
        hi[e] = h;
        lo[e] = __float2half_rn(x - __half2float(h));
    }
}

// ----------------------------------------------------------------------------------------------
// blend_backward: one CTA = 128 (point, corner) rows, like final_blend.
//   phase 1  ob[c][row][o] = d loss / d out_c  (transpose of the product-rule blend), scaled by S
//   phase 2  abar_c[f] = sum_o Wlast[o][f] ob_c[o] -> reverse jet activation with the saved z of the last hidden
//            layer -> zbar planes (fp16 hi/lo, the dgrad / wgrad operands), adjoints of Vb, of the coordinate
//            columns, of the last layer's weight and bias
// ----------------------------------------------------------------------------------------------
// RB2: the Rayleigh-Benard jet set [value | d0, d1, d2 | d11, d22] is known at compile time (straight-line reverse jet
// activation instead of the selector loops of the generic JetSpec); OM: outputs padded to 4 or 8 (ob rows are float4s,
// the last layer's column of a feature and its adjoint live in registers).
// Phase 2 walks (32-feature group, row): a lane owns ONE feature for 16 rows, so the adjoints of the last layer's weight
// and of the coordinate columns accumulate in registers and reach shared memory once per feature group (they were 7
// shared-memory read-modify-writes per (row, feature) in round 1).
template <int KC, bool RB2, int OM>
__global__ void __launch_bounds__(256) blend_backward_kernel(JetSpec spec, BlendBwdArgs a) {
    extern __shared__ float sm[];
    const int O = a.O, Kp = a.Kp, F = a.n_feat, ldo = a.ld_out, dim = a.dim;
    float* Ws = sm;                          // [O][Kp]
    float* ob = Ws + O * Kp;                 // [KC][128][OM]  (outputs >= O stay zero)
    // partial sums of the last layer's weight adjoint [O][ldo] and of the coordinate-column adjoint [ldo][dim]: one
    // private copy per warp when shared memory allows (plain read-modify-write, a lane owns its features), else one
    // shared copy updated with atomics
    const int copies = a.acc_copies, acc_stride = O * ldo + ldo * dim;
    float* gWl = ob + KC * 128 * OM;
    float* gB = gWl + copies * acc_stride;   // [O]
    const int ncorner = 1 << dim;
    const int row0 = blockIdx.x * 128;
    const float S = a.scale[0];
    for (int e = threadIdx.x; e < O * Kp; e += blockDim.x) Ws[e] = a.Wlast[e];
    for (int e = threadIdx.x; e < KC * 128 * OM; e += blockDim.x) ob[e] = 0.f;
    for (int e = threadIdx.x; e < copies * acc_stride + O; e += blockDim.x) gWl[e] = 0.f;
    __syncthreads();

    // ---- phase 1: one thread per (local row, output); it owns ob[.][lr][o] ----
    for (int e = threadIdx.x; e < 128 * O; e += blockDim.x) {
        const int o = e % O, lr = e / O;
        const int lp = lr / ncorner, j = lr % ncorner;
        const int i = row0 / ncorner + lp;
        const int64_t gp = a.p0 + i;
        if (i >= a.cb.pc || gp >= a.total_pts) continue;
        float w = 1.f;
        float fk[kMaxDim], dk[kMaxDim], dx[kMaxDim];
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) {
            fk[k] = 1.f; dk[k] = 0.f; dx[k] = 0.f;
            if (k < dim) {
                const int bit = (j >> (dim - 1 - k)) & 1;
                fk[k] = a.cb.wfac[(k * 2 + bit) * a.cb.pc + i];
                dk[k] = a.cb.dfac[(k * 2 + bit) * a.cb.pc + i];
                dx[k] = a.cb.dxr[k * a.cb.pc + i];
                w *= fk[k];
            }
        }
        for (int c = 0; c < KC; ++c) {
            const float gin = S * (c == 0 ? a.gy[gp * O + o] : a.gjets[((int64_t)(c - 1) * a.total_pts + gp) * O + o]);
            const int kind = spec.kind[c];
            if (kind == 0) { ob[lr * OM + o] += w * gin; continue; }
            const int ca = kind == 2 ? spec.pa[c] : c, cbi = kind == 2 ? spec.pb[c] : c;
            const int da = kind == 1 ? spec.dir[c] : spec.dir[ca], db = spec.dir[cbi];
            float wa = 1.f, wb = 1.f, wab = 1.f, dxa = 0.f, dxb = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) {
                if (k < dim) {
                    wa *= (k == da) ? dk[k] : fk[k];
                    wb *= (k == db) ? dk[k] : fk[k];
                    wab *= (k == da || k == db) ? dk[k] : fk[k];
                }
                if (k == da) dxa = dx[k];
                if (k == db) dxb = dx[k];
            }
            if (kind == 1) {
                ob[lr * OM + o] += wa * gin;
                ob[(c * 128 + lr) * OM + o] += w * dxa * gin;
            } else {
                ob[lr * OM + o] += (da == db ? 0.f : wab) * gin;
                ob[(cbi * 128 + lr) * OM + o] += wa * dxb * gin;
                ob[(ca * 128 + lr) * OM + o] += wb * dxa * gin;
                ob[(c * 128 + lr) * OM + o] += w * dxa * dxb * gin;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: (feature group, local row) items ----
    const int64_t zplane = (int64_t)a.rows * a.ldz, aplane = (int64_t)a.rows * Kp, oplane = (int64_t)a.rows * ldo;
    const bool swish_beta = a.act == STPDE_ACT_SWISH && a.g_beta != nullptr;
    float amax = 0.f, bsum = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* myWl = gWl + (copies > 1 ? warp : 0) * acc_stride;
    float* myWx = myWl + O * ldo;
    dispatch_act(a.act, [&](auto act_c) {
    constexpr int kAct = decltype(act_c)::value;
    for (int f = lane; f < ldo; f += 32) {               // ldo is a multiple of 32: warp-uniform trip count
        const bool fok = f < F;
        float wl[OM], accW[OM], accX[kMaxDim];
#pragma unroll
        for (int o = 0; o < OM; ++o) { wl[o] = (fok && o < O) ? Ws[o * Kp + f] : 0.f; accW[o] = 0.f; }
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) accX[k] = 0.f;
#pragma unroll 2
        for (int lr = warp; lr < 128; lr += 8) {
            const int r = row0 + lr;
            if (r >= a.rows) break;
            float zb[KC];
            if (fok) {
                float z[KC], ab[KC], av[KC];
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const int64_t ze = (int64_t)c * zplane + (int64_t)r * a.ldz + f;
                    z[c] = a.z_half ? __half2float(reinterpret_cast<const __half*>(a.z_in)[ze]) : a.z_in[ze];
                    av[c] = a.act_last[(int64_t)c * aplane + (int64_t)r * Kp + f];
                }
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    float obv[OM];
#pragma unroll
                    for (int o4 = 0; o4 < OM; o4 += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(ob + (c * 128 + lr) * OM + o4);
                        obv[o4] = t.x; obv[o4 + 1] = t.y; obv[o4 + 2] = t.z; obv[o4 + 3] = t.w;
                    }
                    float sacc = 0.f;
#pragma unroll
                    for (int o = 0; o < OM; ++o) {
                        sacc = fmaf(wl[o], obv[o], sacc);
                        accW[o] = fmaf(obv[o], av[c], accW[o]);
                    }
                    ab[c] = sacc;
                }
                float s1, s2, s3;
                act_d123_fast(kAct, a.beta, z[0], s1, s2, s3);
                if constexpr (RB2) {
                    // o_0 = s(z_0), o_c = s' z_c (c = 1..3), o_4 = s'' z_2^2 + s' z_4, o_5 = s'' z_3^2 + s' z_5
                    float z0b = s1 * ab[0];
                    z0b = fmaf(s2, fmaf(ab[1], z[1], fmaf(ab[2], z[2], ab[3] * z[3])), z0b);
                    z0b = fmaf(ab[4], fmaf(s3 * z[2], z[2], s2 * z[4]), z0b);
                    z0b = fmaf(ab[5], fmaf(s3 * z[3], z[3], s2 * z[5]), z0b);
                    zb[0] = z0b;
                    zb[1] = s1 * ab[1];
                    zb[2] = fmaf(2.f * s2 * z[2], ab[4], s1 * ab[2]);
                    zb[3] = fmaf(2.f * s2 * z[3], ab[5], s1 * ab[3]);
                    zb[4] = s1 * ab[4];
                    zb[5] = s1 * ab[5];
                } else {
                    jet_act_backward<KC>(spec, s1, s2, s3, z, ab, zb);
                }
                if (kAct == STPDE_ACT_SWISH && swish_beta) {
                    float u = 0.f, w3 = 0.f, sb0, sb1, sb2;
#pragma unroll
                    for (int c = 1; c < KC; ++c) {
                        u = fmaf(ab[c], z[c], u);
                        float za = 0.f, zp = 0.f;
#pragma unroll
                        for (int k = 0; k < STPDE_MAX_FIRST; ++k) {
                            if (1 + k < KC) {
                                za = fmaf(spec.sel_a[c][k], z[1 + k], za);
                                zp = fmaf(spec.sel_b[c][k], z[1 + k], zp);
                            }
                        }
                        w3 = fmaf(ab[c] * za, zp, w3);
                    }
                    swish_dbeta(a.beta, z[0], sb0, sb1, sb2);
                    bsum += fmaf(ab[0], sb0, fmaf(sb1, u, sb2 * w3));
                }
                const int vrow = a.cb.vtx[r];
                atomicAdd(a.g_vb + (int64_t)vrow * a.ncat + a.cat_off + f, zb[0]);
#pragma unroll
                for (int k = 0; k < kMaxDim; ++k) {
                    if (k < dim) {
                        float sx = zb[0] * a.cb.xrel[(int64_t)k * a.rows + r];
                        if constexpr (RB2) {
                            if (k < 3) sx += zb[1 + k];
                        } else {
#pragma unroll
                            for (int c = 1; c < KC; ++c)
                                if (spec.kind[c] == 1 && spec.dir[c] == k) sx += zb[c];
                        }
                        accX[k] += sx;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < KC; ++c) zb[c] = 0.f;
            }
            const int64_t off = (int64_t)r * ldo + f;
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const float xs = zb[c];
                amax = fmaxf(amax, fabsf(xs));
                const __half hi = __float2half_rn(xs);
                a.out_hi[(int64_t)c * oplane + off] = hi;
                if (a.three) a.out_lo[(int64_t)c * oplane + off] = __float2half_rn(xs - __half2float(hi));
            }
        }
        if (fok) {                                          // one shared-memory update per (feature, warp)
#pragma unroll
            for (int o = 0; o < OM; ++o) {
                if (o < O) {
                    if (copies > 1) myWl[o * ldo + f] += accW[o];
                    else atomicAdd(myWl + o * ldo + f, accW[o]);
                }
            }
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) {
                if (k < dim) {
                    if (copies > 1) myWx[f * dim + k] += accX[k];
                    else atomicAdd(myWx + f * dim + k, accX[k]);
                }
            }
        }
    }
    });
    if (!(amax < 65000.f)) atomicOr(a.status, kStatusRange);
    if (swish_beta) {
        for (int off = 16; off > 0; off >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.g_beta, bsum);
    }
    for (int e = threadIdx.x; e < 128 * O; e += blockDim.x) {
        if (row0 + e / O < a.rows) atomicAdd(gB + e % O, ob[(e / O) * OM + e % O]);   // component 0 only: the bias enters the value
    }
    __syncthreads();

    // ---- flush the CTA's partial sums ----
    for (int e = threadIdx.x; e < O * ldo; e += blockDim.x) {
        const int o = e / ldo, f = e % ldo;
        float v = 0.f;
        for (int cp = 0; cp < copies; ++cp) v += gWl[cp * acc_stride + e];
        if (f < F) atomicAdd(a.g_wlast + (int64_t)o * F + f, v);
    }
    for (int e = threadIdx.x; e < F * dim; e += blockDim.x) {
        float v = 0.f;
        for (int cp = 0; cp < copies; ++cp) v += gWl[cp * acc_stride + O * ldo + e];
        atomicAdd(a.g_wx + (int64_t)(e / dim) * a.g_wx_ld + e % dim, v);
    }
    for (int e = threadIdx.x; e < O; e += blockDim.x) atomicAdd(a.g_blast + e, gB[e]);
}

// ----------------------------------------------------------------------------------------------
// vertex_backward (weights): gb_l[n] = sum_v gVb[v][cat],  gW_l[n][kh + dim + ch] = sum_v gVb[v][cat] latent[v][ch]
// block = 128 cats x one slice of kVbSlice vertices; channel tiles of 32 keep the accumulators in registers
// ----------------------------------------------------------------------------------------------
// vertices per block: a multiple of 16 sized so that ~4 blocks per SM exist - every block ends with one atomicAdd per
// (cat, channel), so smaller slices only multiply the atomics on the same ncat x c addresses (10 M of them at 32
// vertices per block for the reference-size step: 0.3 of its 3.3 ms)
static int vb_slice(int nvert_total, int cat_blocks) {
    const int want_slices = (148 * 4 + cat_blocks - 1) / cat_blocks;
    int slice = (nvert_total + want_slices - 1) / want_slices;
    slice = (slice + 15) / 16 * 16;
    return slice < 32 ? 32 : slice;
}

__global__ void __launch_bounds__(128) vertex_backward_w_kernel(GridGeom g, int nvert_total, VertexBwdArgs a, int kVbSlice) {
    __shared__ __align__(16) float lat[16][32];
    const int cat = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = cat < a.ncat;
    int l = 0;
    if (ok) while (l + 1 < a.n_layers - 1 && cat >= a.cat_off[l + 1]) ++l;
    const int n = cat - a.cat_off[l];
    const int v_begin = blockIdx.y * kVbSlice, v_end = min(v_begin + kVbSlice, nvert_total);
    const int c = g.channels;
    float bsum = 0.f;
    for (int ct = 0; ct < c; ct += 32) {
        float acc[32];
#pragma unroll
        for (int ch = 0; ch < 32; ++ch) acc[ch] = 0.f;
        for (int v0 = v_begin; v0 < v_end; v0 += 16) {
            __syncthreads();
            for (int e = threadIdx.x; e < 16 * 32; e += blockDim.x) {
                const int vi = e / 32, ch = ct + e % 32;
                const int v = v0 + vi;
                float val = 0.f;
                if (v < v_end && ch < c) {
                    const int b = v / g.nvert;
                    int rem = v % g.nvert;
                    int64_t off = b * g.gstride[0] + ch * g.gstride[g.dim + 1];
                    for (int k = g.dim - 1; k >= 0; --k) {
                        off += (rem % g.size[k]) * g.gstride[1 + k];
                        rem /= g.size[k];
                    }
                    val = a.grid[off];
                }
                lat[vi][e % 32] = val;
            }
            __syncthreads();
            if (ok) {
#pragma unroll 4
                for (int vi = 0; vi < 16; ++vi) {
                    if (v0 + vi >= v_end) break;
                    const float gv = a.g_vb[(int64_t)(v0 + vi) * a.ncat + cat];
                    if (ct == 0) bsum += gv;
#pragma unroll
                    for (int ch = 0; ch < 32; ch += 4) {               // (all threads read the same 16 bytes: broadcast)
                        const float4 q = *reinterpret_cast<const float4*>(&lat[vi][ch]);
                        acc[ch] = fmaf(gv, q.x, acc[ch]);
                        acc[ch + 1] = fmaf(gv, q.y, acc[ch + 1]);
                        acc[ch + 2] = fmaf(gv, q.z, acc[ch + 2]);
                        acc[ch + 3] = fmaf(gv, q.w, acc[ch + 3]);
                    }
                }
            }
        }
        if (ok) {
            float* dst = a.gW[l] + (int64_t)n * a.in_features[l] + a.kh[l] + g.dim + ct;
#pragma unroll
            for (int ch = 0; ch < 32; ++ch)
                if (ct + ch < c) atomicAdd(dst + ch, acc[ch]);
        }
    }
    if (ok) atomicAdd(a.gB[l] + n, bsum);
}

// vertex_backward (latent grid): ggrid[v][ch] += (1/S) sum_{cat in slice} gVb[v][cat] * W_l(cat)[n(cat)][kh + dim + ch]
// A small GEMM [nvert x ncat] . [ncat x c]: block = 32 vertices x 32 channels x one slice of kGridCatSlice cats
// (blockIdx.z); tiles of 64 cats of both operands are staged in shared memory, a thread owns 4 vertices of one channel
// (one conflict-free W read and four broadcast 16-byte gVb reads per 4 cats and 16 FMAs).  ggrid is zeroed by the caller
// and the slices are added with atomics.  (The first version re-read W from global memory for every 8 vertices: 0.25 ms
// of the 3.2 ms reference-size training step.)
constexpr int kGridCatSlice = 512;
constexpr int kGridVerts = 32, kGridCats = 64;

__global__ void __launch_bounds__(256) vertex_backward_grid_kernel(GridGeom g, int nvert_total, VertexBwdArgs a,
                                                                   float* __restrict__ ggrid) {
    __shared__ __align__(16) float gvt[kGridVerts][kGridCats];
    __shared__ float wt[kGridCats][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.y * 32 + lane;
    const int v0 = blockIdx.x * kGridVerts;
    const int c = g.channels;
    const int cat_begin = blockIdx.z * kGridCatSlice, cat_end = min(cat_begin + kGridCatSlice, a.ncat);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int cat0 = cat_begin; cat0 < cat_end; cat0 += kGridCats) {
        __syncthreads();
        for (int e = threadIdx.x; e < kGridVerts * kGridCats; e += blockDim.x) {
            const int vi = e / kGridCats, t = e % kGridCats;
            const int v = v0 + vi, cat = cat0 + t;
            gvt[vi][t] = (v < nvert_total && cat < cat_end) ? a.g_vb[(int64_t)v * a.ncat + cat] : 0.f;
        }
        for (int t = warp; t < kGridCats; t += 8) {
            const int cat = cat0 + t;
            float w = 0.f;
            if (cat < cat_end && ch < c) {
                int l = 0;
                while (l + 1 < a.n_layers - 1 && cat >= a.cat_off[l + 1]) ++l;
                w = __ldg(a.W[l] + (int64_t)(cat - a.cat_off[l]) * a.in_features[l] + a.kh[l] + g.dim + ch);
            }
            wt[t][lane] = w;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < kGridCats; t += 4) {
            const float w0 = wt[t][lane], w1 = wt[t + 1][lane], w2 = wt[t + 2][lane], w3 = wt[t + 3][lane];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = *reinterpret_cast<const float4*>(&gvt[warp * 4 + i][t]);
                acc[i] = fmaf(q.x, w0, fmaf(q.y, w1, fmaf(q.z, w2, fmaf(q.w, w3, acc[i]))));
            }
        }
    }
    const float inv_s = a.scale[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = v0 + warp * 4 + i;
        if (v < nvert_total && ch < c) atomicAdd(ggrid + (int64_t)v * c + ch, acc[i] * inv_s);
    }
}

// ----------------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

void launch_grad_scale(const JetSpec& spec, const GridGeom& g, const float* gy, const float* gjets, int64_t plane_elems,
                       unsigned* maxes, float* scale, int target_exp, cudaStream_t st) {
    cudaMemsetAsync(maxes, 0, sizeof(unsigned) * kMaxComp, st);
    const int n_planes = gjets ? spec.kc : 1;
    dim3 grid(grid_for(plane_elems, 256), n_planes);
    absmax_planes_kernel<<<grid, 256, 0, st>>>(gy, gjets, plane_elems, n_planes, maxes);
    grad_scale_kernel<<<1, 32, 0, st>>>(spec, g, maxes, scale, target_exp);
}

static dim3 buffer_list_grid(const BufferList& bl) {
    int64_t most = 1;
    for (int i = 0; i < bl.count; ++i) most = bl.n[i] > most ? bl.n[i] : most;
    return dim3((unsigned)grid_for(most, 256), (unsigned)bl.count);
}

void launch_zero_buffers(const BufferList& bl, cudaStream_t st) {
    if (bl.count > 0) buffer_list_kernel<false><<<buffer_list_grid(bl), 256, 0, st>>>(bl, nullptr);
}

void launch_scale_buffers(const BufferList& bl, const float* scale, cudaStream_t st) {
    if (bl.count > 0) buffer_list_kernel<true><<<buffer_list_grid(bl), 256, 0, st>>>(bl, scale);
}

void launch_split_weights_t(const float* W, int N, int in_features, int kh, int fp, int ldz, int pack, const unsigned* absmax,
                            __half* hi, __half* lo, cudaStream_t st) {
    split_weights_t_kernel<<<148 * 4, 256, 0, st>>>(W, N, in_features, kh, fp, ldz, pack, absmax, hi, lo);
}

size_t blend_backward_smem(int kc, int O, int Kp, int ldo, int dim, int copies) {
    const int om = O <= 4 ? 4 : 8;         // ob rows are padded to float4s
    return (size_t)(O * Kp + kc * 128 * om + copies * (O * ldo + ldo * dim) + O) * sizeof(float);
}

template <int KC, bool RB2, int OM>
static int launch_blend_backward_t(const JetSpec& spec, const BlendBwdArgs& a_in, cudaStream_t st) {
    BlendBwdArgs a = a_in;
    a.acc_copies = blend_backward_smem(KC, a.O, a.Kp, a.ld_out, a.dim, 8) <= 96 * 1024 ? 8 : 1;   // 2 CTAs / SM stay resident
    const size_t smem = blend_backward_smem(KC, a.O, a.Kp, a.ld_out, a.dim, a.acc_copies);
    if (smem > 200 * 1024) return STPDE_EUNSUPPORTED;
    static DeviceOnce configured;
    if (configured.first_use()) {
        cudaFuncSetAttribute(blend_backward_kernel<KC, RB2, OM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.mark();
    }
    blend_backward_kernel<KC, RB2, OM><<<(a.rows + 127) / 128, 256, smem, st>>>(spec, a);
    return STPDE_OK;
}

template <int KC>
static int launch_blend_backward_k(const JetSpec& spec, const BlendBwdArgs& a, cudaStream_t st) {
    return a.O <= 4 ? launch_blend_backward_t<KC, false, 4>(spec, a, st) : launch_blend_backward_t<KC, false, 8>(spec, a, st);
}

int launch_blend_backward(const JetSpec& spec, const BlendBwdArgs& a, cudaStream_t st) {
    if (a.ld_out % 32 || a.O > kMaxOut) return STPDE_EUNSUPPORTED;
    // the Rayleigh-Benard jet set [value | d0, d1, d2 | d11, d22] runs the compile-time specialisation
    const bool rb2 = spec.kc == 6 && spec.n_first == 3 && spec.n_second == 2 && spec.dir[1] == 0 && spec.dir[2] == 1 &&
                     spec.dir[3] == 2 && spec.pa[4] == 2 && spec.pb[4] == 2 && spec.pa[5] == 3 && spec.pb[5] == 3;
    if (rb2) return a.O <= 4 ? launch_blend_backward_t<6, true, 4>(spec, a, st) : launch_blend_backward_t<6, true, 8>(spec, a, st);
    int rc = STPDE_OK;
    switch (spec.kc) {
        case 1: rc = launch_blend_backward_k<1>(spec, a, st); break;
        case 2: rc = launch_blend_backward_k<2>(spec, a, st); break;
        case 3: rc = launch_blend_backward_k<3>(spec, a, st); break;
        case 4: rc = launch_blend_backward_k<4>(spec, a, st); break;
        case 5: rc = launch_blend_backward_k<5>(spec, a, st); break;
        case 6: rc = launch_blend_backward_k<6>(spec, a, st); break;
        case 7: rc = launch_blend_backward_k<7>(spec, a, st); break;
        case 8: rc = launch_blend_backward_k<8>(spec, a, st); break;
        case 9: rc = launch_blend_backward_k<9>(spec, a, st); break;
        default: rc = launch_blend_backward_k<10>(spec, a, st); break;
    }
    return rc;
}

void launch_vertex_backward(const GridGeom& g, int nvert_total, const VertexBwdArgs& a, float* ggrid, cudaStream_t st) {
    const int cat_blocks = (a.ncat + 127) / 128, slice = vb_slice(nvert_total, cat_blocks);
    dim3 gw(cat_blocks, (nvert_total + slice - 1) / slice);
    vertex_backward_w_kernel<<<gw, 128, 0, st>>>(g, nvert_total, a, slice);
    if (ggrid) {
        dim3 gg((nvert_total + kGridVerts - 1) / kGridVerts, (g.channels + 31) / 32, (a.ncat + kGridCatSlice - 1) / kGridCatSlice);
        vertex_backward_grid_kernel<<<gg, 256, 0, st>>>(g, nvert_total, a, ggrid);
    }
}

}  // namespace stpde
