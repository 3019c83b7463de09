// Host-visible declarations of the kernel launchers (internal; the public surface is include/stpde.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

#include <atomic>

namespace stpde {

// One-time per-device configuration flag (cudaFuncSetAttribute is per device).  The mask is atomic: launch wrappers may
// be entered from several host threads (one per device / stream); a duplicated attribute call is harmless.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool first_use() const {
        int d = 0;
        cudaGetDevice(&d);
        return !((mask.load(std::memory_order_acquire) >> (d & 63)) & 1ull);
    }
    void mark() {
        int d = 0;
        cudaGetDevice(&d);
        mask.fetch_or(1ull << (d & 63), std::memory_order_release);
    }
};


// Per-chunk scratch arrays (device pointers) produced by prep_points.
struct ChunkBuffers {
    int pc;        // points in the chunk (multiple of 128)
    int rows;      // pc * 2^d
    int* vtx;      // [rows]      flattened (batch, vertex) index into Vb
    float* xrel;   // [d][rows]   relative coordinate of the point w.r.t. the row's corner
    float* wfac;   // [d][2][pc]  blend factor per dimension and corner bit
    float* dfac;   // [d][2][pc]  d(blend factor)/dq  (sign * clipgrad / cubesize)
    float* dxr;    // [d][pc]     d(x_rel)/dq = clipgrad / cubesize
};

// Decoder description for vertex_bias (device pointers to the caller's weights).
struct NetDesc {
    int n_layers;
    int ncat;                        // sum of widths[0..n_layers-2]
    int cat_off[kMaxLayers];         // offset of layer l inside a Vb row
    int in_features[kMaxLayers];     // row stride of W[l]
    int kh[kMaxLayers];              // number of activation columns (0 for layer 0)
    const float* W[kMaxLayers];
    const float* B[kMaxLayers];
};

struct ResidualProgram {
    int n_words;
    int words[640];
    float consts[128];
};

void launch_interp_coeff(const GridGeom& g, int batch, int npts, const float* grid, const float* q, float* cv,
                         float* w, float* xr, int* status, cudaStream_t st);
void launch_interp(const GridGeom& g, int batch, int npts, const float* grid, const float* q, float* out,
                   int* status, cudaStream_t st);
void launch_prep_points(const GridGeom& g, int npts, int64_t total_pts, int64_t p0, const ChunkBuffers& cb,
                        const float* q, int* status, cudaStream_t st);
void launch_vertex_bias(const GridGeom& g, int nvert_total, const NetDesc& net, const float* grid, float* Vb,
                        cudaStream_t st);
void launch_pack_weights(const float* W, int N, int in_features, int kh, int dim, int Np, int Kp, float* Wh,
                         float* Wx, cudaStream_t st);
void launch_layer0(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np, const int* vtx,
                   const float* xrel, const float* Wx, const float* Vb, int ncat, float* out, cudaStream_t st);
void launch_layer_gemm(const JetSpec& spec, int dim, int act, float beta, int rows, int N, int Np, int Kp, int NpOut,
                       const float* actIn, const float* Wh, const float* Wx, const float* Vb, int ncat, int cat_off,
                       const int* vtx, const float* xrel, float* out, cudaStream_t st);
void launch_final_blend(const JetSpec& spec, int dim, int rows, int pc, int64_t total_pts, int64_t p0, int Kp, int O,
                        const float* actIn, const float* Wlast, const float* blast, const ChunkBuffers& cb, float* y,
                        float* jets, cudaStream_t st);
// adjoint programs (one per output symbol: o values, then n_jet * o jet entries) can be longer
struct ResidualProgramBig {
    int n_words;
    int words[2048];
    float consts[256];
};
void launch_residuals_backward(const ResidualProgramBig& prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                               int n_eq, const float* q, const int64_t* qs, const float* y, const float* jets,
                               const float* gres, float* gy, float* gjets, cudaStream_t st);
void launch_residuals(const ResidualProgram& prog, int npts, int64_t total_pts, int dim, int O, int n_jet,
                      const float* q, const int64_t* qs, const float* y, const float* jets, float* residuals,
                      cudaStream_t st);

int residual_loss_blocks(int64_t total_pts);
void launch_residual_loss(const ResidualProgram& prog, int npts, int64_t total_pts, int O, int n_eq, int loss_kind,
                          const float* q, const int64_t* qs, const float* y, const float* jets, const float* target,
                          float* partial, cudaStream_t st);
void launch_residual_loss_backward(const ResidualProgram& fwd, const ResidualProgramBig& adj, int npts, int64_t total_pts,
                                   int O, int n_jet, int n_eq, int loss_kind, const float* q, const int64_t* qs,
                                   const float* y, const float* jets, const float* target, const float* g_sums, float* gy,
                                   float* gjets, cudaStream_t st);

}  // namespace stpde
