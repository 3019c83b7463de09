// Shared device/host definitions for the stpde kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/stpde.h"

namespace stpde {

constexpr int kMaxDim = STPDE_MAX_DIM;
constexpr int kMaxComp = STPDE_MAX_COMPONENTS;
constexpr int kMaxLayers = STPDE_MAX_LAYERS;
constexpr int kMaxOut = STPDE_MAX_OUT;

// status bits written by kernels
constexpr int kStatusIndex = 1;  // cell index outside the grid (reference raises IndexError)
constexpr int kStatusRange = 2;  // split-precision operand outside fp16 range

// Jet components propagated through the MLP.  Component 0 is the value; components
// 1..n_first are d/dxrel_{first_dirs[c-1]}; components 1+n_first.. are second derivatives whose
// parents are the first-order components sec_a / sec_b.
struct JetSpec {
    int kc;
    int n_first;
    int n_second;
    int first_dirs[STPDE_MAX_FIRST];
    int sec_a[STPDE_MAX_SECOND];  // component index of d/dx_i
    int sec_b[STPDE_MAX_SECOND];  // component index of d/dx_j
    // per-component view (indexed by the component, so fully unrolled kernels read these straight
    // from the constant bank): kind 0 = value, 1 = first order along dir, 2 = second order with
    // first-order parent components pa, pb
    int kind[STPDE_MAX_COMPONENTS];
    int dir[STPDE_MAX_COMPONENTS];
    int pa[STPDE_MAX_COMPONENTS];
    int pb[STPDE_MAX_COMPONENTS];
    // one-hot selectors of the parents over the first-order components 1..4 (all zero unless kind == 2):
    // z_pa = sum_k sel_a[c][k] * z[1 + k] is 4 FFMAs with constant-bank operands instead of compare/select chains
    float sel_a[STPDE_MAX_COMPONENTS][STPDE_MAX_FIRST];
    float sel_b[STPDE_MAX_COMPONENTS][STPDE_MAX_FIRST];
};

// Geometry of the latent grid and the clip / cell arithmetic, all float32 exactly as the
// reference forms it (regular_nd_grid_interpolation.py:37-51).
struct GridGeom {
    int dim;
    int size[kMaxDim];
    int channels;
    float lo[kMaxDim];        // xmin + eps
    float hi[kMaxDim];        // xmax - eps
    float cubesize[kMaxDim];  // (xmax - xmin) / (size - 1)
    int64_t gstride[kMaxDim + 2];  // element strides of grid [b, n_1..n_d, c]
    int64_t qstride[3];            // element strides of q [b, p, d]
    int nvert;                     // prod(size)
};

// sigma, sigma', sigma'' with torch autograd conventions (reference src/nonlinearities.py:15-22).
__device__ __forceinline__ void act_jet(int kind, float beta, float z, float& s0, float& s1, float& s2) {
    switch (kind) {
        case STPDE_ACT_TANH: {
            float t = tanhf(z);
            float u = 1.f - t * t;
            s0 = t; s1 = u; s2 = -2.f * t * u;
            break;
        }
        case STPDE_ACT_RELU: {
            bool p = z > 0.f;
            s0 = p ? z : 0.f; s1 = p ? 1.f : 0.f; s2 = 0.f;
            break;
        }
        case STPDE_ACT_LEAKYRELU: {
            bool p = z > 0.f;
            s0 = p ? z : 0.01f * z; s1 = p ? 1.f : 0.01f; s2 = 0.f;
            break;
        }
        case STPDE_ACT_SOFTPLUS: {
            if (z > 20.f) { s0 = z; s1 = 1.f; s2 = 0.f; }
            else {
                float e = expf(z);
                float s = e / (1.f + e);           // torch softplus_backward: z / (z + 1)
                s0 = log1pf(e); s1 = s; s2 = s * (1.f - s);
            }
            break;
        }
        case STPDE_ACT_ELU: {
            if (z <= 0.f) { float e = expf(z); s0 = e - 1.f; s1 = e; s2 = e; }
            else { s0 = z; s1 = 1.f; s2 = 0.f; }
            break;
        }
        default: {  // STPDE_ACT_SWISH
            float bz = beta * z;
            float s = 1.f / (1.f + expf(-bz));
            float ds = s * (1.f - s);
            s0 = z * s; s1 = s + bz * ds; s2 = beta * ds * (2.f + bz * (1.f - 2.f * s));
            break;
        }
    }
}

// Same functions with MUFU-based exp / log / reciprocal (absolute error ~1e-7 .. 4e-7, i.e. the size of the
// 2^-22 operand rounding the split-precision tensor-core path already carries).  Used by the tensor-core
// kernels only; the FP32 path keeps the libdevice-accurate versions above.
__device__ __forceinline__ void act_jet_fast(int kind, float beta, float z, float& s0, float& s1, float& s2) {
    switch (kind) {
        case STPDE_ACT_TANH: {
            const float e = __expf(2.f * z);                 // tanh z = 1 - 2 / (1 + e^{2z})
            const float t = 1.f - __fdividef(2.f, 1.f + e);
            const float u = 1.f - t * t;
            s0 = t; s1 = u; s2 = -2.f * t * u;
            break;
        }
        case STPDE_ACT_RELU: {
            const bool p = z > 0.f;
            s0 = p ? z : 0.f; s1 = p ? 1.f : 0.f; s2 = 0.f;
            break;
        }
        case STPDE_ACT_LEAKYRELU: {
            const bool p = z > 0.f;
            s0 = p ? z : 0.01f * z; s1 = p ? 1.f : 0.01f; s2 = 0.f;
            break;
        }
        case STPDE_ACT_SOFTPLUS: {
            const bool lin = z > 20.f;
            const float e = __expf(fminf(z, 20.f));
            const float w = 1.f + e;
            const float r = __fdividef(1.f, w);
            const float s = e * r;
            s0 = lin ? z : __logf(w);
            s1 = lin ? 1.f : s;
            s2 = lin ? 0.f : s * r;                           // s (1 - s) = s / (1 + e)
            break;
        }
        case STPDE_ACT_ELU: {
            const bool neg = z <= 0.f;
            const float e = __expf(fminf(z, 0.f));
            s0 = neg ? e - 1.f : z; s1 = neg ? e : 1.f; s2 = neg ? e : 0.f;
            break;
        }
        default: {  // STPDE_ACT_SWISH
            const float bz = beta * z;
            const float s = __fdividef(1.f, 1.f + __expf(-bz));
            const float ds = s * (1.f - s);
            s0 = z * s; s1 = s + bz * ds; s2 = beta * ds * (2.f + bz * (1.f - 2.f * s));
            break;
        }
    }
}

// sigma', sigma'', sigma''' (MUFU-based): the reverse-mode sweep through a second-order jet needs the third
// derivative (d a_kl / d z_0 = sigma''' z_k z_l + sigma'' z_kl).  Same branch conventions as act_jet_fast.
__device__ __forceinline__ void act_d123_fast(int kind, float beta, float z, float& s1, float& s2, float& s3) {
    switch (kind) {
        case STPDE_ACT_TANH: {
            const float e = __expf(2.f * z);
            const float t = 1.f - __fdividef(2.f, 1.f + e);
            const float u = 1.f - t * t;
            s1 = u; s2 = -2.f * t * u; s3 = -2.f * u * (1.f - 3.f * t * t);
            break;
        }
        case STPDE_ACT_RELU: {
            s1 = z > 0.f ? 1.f : 0.f; s2 = 0.f; s3 = 0.f;
            break;
        }
        case STPDE_ACT_LEAKYRELU: {
            s1 = z > 0.f ? 1.f : 0.01f; s2 = 0.f; s3 = 0.f;
            break;
        }
        case STPDE_ACT_SOFTPLUS: {
            const bool lin = z > 20.f;
            const float e = __expf(fminf(z, 20.f));
            const float r = __fdividef(1.f, 1.f + e);
            const float s = e * r;
            s1 = lin ? 1.f : s;
            s2 = lin ? 0.f : s * r;
            s3 = lin ? 0.f : s * r * (1.f - 2.f * s);
            break;
        }
        case STPDE_ACT_ELU: {
            const bool neg = z <= 0.f;
            const float e = __expf(fminf(z, 0.f));
            s1 = neg ? e : 1.f; s2 = neg ? e : 0.f; s3 = neg ? e : 0.f;
            break;
        }
        default: {  // STPDE_ACT_SWISH
            const float bz = beta * z;
            const float s = __fdividef(1.f, 1.f + __expf(-bz));
            const float ds = s * (1.f - s);
            const float m = 1.f - 2.f * s;
            s1 = s + bz * ds;
            s2 = beta * ds * (2.f + bz * m);
            s3 = beta * beta * ds * (m * (3.f + bz * m) - 2.f * bz * ds);
            break;
        }
    }
}

// Swish x * sigmoid(beta x): partial derivatives of sigma, sigma', sigma'' w.r.t. the learnable beta at fixed z
// (reference src/nonlinearities.py:5-13 makes beta an nn.Parameter, so loss.backward() needs its gradient):
//   d loss / d beta += ab_0 sb0 + sb1 sum_{c>=1} ab_c z_c + sb2 sum_{second order c} ab_c z_a z_b
__device__ __forceinline__ void swish_dbeta(float beta, float z, float& sb0, float& sb1, float& sb2) {
    const float bz = beta * z;
    const float s = __fdividef(1.f, 1.f + __expf(-bz));
    const float ds = s * (1.f - s);
    const float m = 1.f - 2.f * s;
    const float t = 2.f + bz * m;
    sb0 = z * z * ds;
    sb1 = z * ds * t;
    sb2 = ds * (t * (1.f + bz * m) + bz * m - 2.f * bz * bz * ds);
}

// Reverse-mode sweep through the jet activation of one (row, feature):
//   forward   o_0 = s0(z_0),  o_k = s1 z_k (first order),  o_c = s2 z_a z_b + s1 z_c (second order, parents a, b)
//   backward  zb_c = d loss / d z_c  from  ob_c = d loss / d o_c
// sel_a / sel_b are the one-hot parent selectors of JetSpec (all zero for first-order components).
template <int KC>
__device__ __forceinline__ void jet_act_backward(const JetSpec& spec, float s1, float s2, float s3, const float* z,
                                                 const float* ob, float* zb) {
    float z0b = s1 * ob[0];
    float cross[STPDE_MAX_FIRST];
#pragma unroll
    for (int k = 0; k < STPDE_MAX_FIRST; ++k) cross[k] = 0.f;
#pragma unroll
    for (int c = 1; c < KC; ++c) {
        float za = 0.f, zbp = 0.f;
#pragma unroll
        for (int k = 0; k < STPDE_MAX_FIRST; ++k) {
            if (1 + k < KC) {
                za = fmaf(spec.sel_a[c][k], z[1 + k], za);
                zbp = fmaf(spec.sel_b[c][k], z[1 + k], zbp);
            }
        }
        z0b = fmaf(ob[c], fmaf(s3 * za, zbp, s2 * z[c]), z0b);
        zb[c] = s1 * ob[c];
#pragma unroll
        for (int k = 0; k < STPDE_MAX_FIRST; ++k)
            if (1 + k < KC) cross[k] = fmaf(ob[c], fmaf(spec.sel_a[c][k], zbp, spec.sel_b[c][k] * za), cross[k]);
    }
    zb[0] = z0b;
#pragma unroll
    for (int k = 0; k < STPDE_MAX_FIRST; ++k)
        if (1 + k < KC) zb[1 + k] = fmaf(s2, cross[k], zb[1 + k]);
}

// Hoists the activation switch OUT of an unrolled per-row loop: f is called once with a compile-time activation tag, so
// its body (the rows of an epilogue block) is straight-line code and the compiler can overlap the MUFU / FMA chains of
// different rows.  With the switch inside the loop every row was its own chain of basic blocks (all six activations
// inlined per row) and the epilogue ran latency-bound at ~0.04 IPC per warp.
template <class F>
__device__ __forceinline__ void dispatch_act(int act, F&& f) {
    switch (act) {
        case STPDE_ACT_TANH: f(std::integral_constant<int, STPDE_ACT_TANH>{}); break;
        case STPDE_ACT_RELU: f(std::integral_constant<int, STPDE_ACT_RELU>{}); break;
        case STPDE_ACT_SOFTPLUS: f(std::integral_constant<int, STPDE_ACT_SOFTPLUS>{}); break;
        case STPDE_ACT_ELU: f(std::integral_constant<int, STPDE_ACT_ELU>{}); break;
        case STPDE_ACT_LEAKYRELU: f(std::integral_constant<int, STPDE_ACT_LEAKYRELU>{}); break;
        default: f(std::integral_constant<int, STPDE_ACT_SWISH>{}); break;
    }
}

// runtime switch used by the tensor-core kernels: accurate (libdevice) in the fp32-parity mode fp16x3 unless
// STPDE_TC_FAST_ACT=1, MUFU-based in the relaxed single-pass fp16 mode
__device__ __forceinline__ void act_jet_sel(bool fast, int kind, float beta, float z, float& s0, float& s1, float& s2) {
    if (fast) act_jet_fast(kind, beta, z, s0, s1, s2);
    else act_jet(kind, beta, z, s0, s1, s2);
}

}  // namespace stpde
