/*
 * stpde.h - C ABI of the B200-native decode + PDE-residual hot path.
 *
 * The reference (maxjiang93/space_time_pde) is pure Python and has no FFI layer; the
 * drop-in boundary is its Python call surface.  Every entry point below
 * replaces one reference function on that surface and is what a ctypes binding (see
 * INTEGRATION.md) calls from the reference-side modules:
 *
 *   stpde_interp_coefficients  <- src/regular_nd_grid_interpolation.py:14-78
 *                                 regular_nd_grid_interpolation_coefficients()
 *   stpde_interp               <- src/regular_nd_grid_interpolation.py:81-104
 *                                 regular_nd_grid_interpolation()
 *   stpde_jet_forward          <- src/local_implicit_grid.py:10-61 query_local_implicit_grid()
 *                                 + src/implicit_net.py:40-54 ImNet.forward()
 *                                 + src/pde.py:8-9,115-143 PDELayer.__call__ / torch_diff
 *                                 (values and every partial derivative the equation strings need,
 *                                  in ONE pass: forward-mode jets instead of autograd.grad per dif())
 *   stpde_jet_backward         <- the loss.backward() sweep through the three items above
 *                                 (experiments/rb2d/train.py:77; src/pde.py:8 create_graph=True makes the
 *                                 reference differentiate THROUGH every autograd.grad call): gradients w.r.t. the
 *                                 decoder weights / biases and the latent grid in one reverse sweep over the jets
 *   stpde_residuals            <- src/pde.py:139-142 (evaluation of the lambdified equations)
 *   stpde_residual_loss        <- the same + experiments/rb2d/train.py:70-75 (l1 / l2 / huber loss reductions)
 *   stpde_jet_forward_host     <- same as stpde_jet_forward with HOST buffers (copies inside)
 *
 * Conventions
 *   - plain C, no torch types; all sizes explicit; returns 0 on success or a negative
 *     STPDE_E* code, never throws.  stpde_last_error() returns a thread-local message.
 *   - device entry points take caller-owned DEVICE pointers, allocate nothing, and are
 *     asynchronous on the given cudaStream_t (passed as void*).  Scratch memory is supplied by
 *     the caller (stpde_workspace_bytes()).
 *   - arithmetic is float32 end to end (cell indices int32); strides are in ELEMENTS.
 *   - the kernels reproduce the reference's quirks: ind0 = floor(q/cubesize) ignores xmin
 *     (python negative-index wrap, out-of-range -> status flag), clip ties have gradient 0.5.
 */
#ifndef STPDE_H_
#define STPDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STPDE_VERSION 200 /* major*10000 + minor*100 + patch */

#define STPDE_MAX_DIM 4
#define STPDE_MAX_LAYERS 8
#define STPDE_MAX_FIRST 4
#define STPDE_MAX_SECOND 10
#define STPDE_MAX_COMPONENTS 10 /* value + first + second components propagated by one launch */
#define STPDE_MAX_OUT 8

/* error codes */
#define STPDE_OK 0
#define STPDE_EINVAL (-1)     /* bad descriptor / argument */
#define STPDE_ENOMEM (-2)     /* workspace too small */
#define STPDE_ECUDA (-3)      /* CUDA runtime error (message in stpde_last_error) */
#define STPDE_EINDEX (-4)     /* a query point addressed a cell outside the grid (reference: IndexError) */
#define STPDE_EUNSUPPORTED (-5)
#define STPDE_ERANGE (-6)     /* split-precision operand left the fp16 range (tensor-core path) */

/* activations: reference src/nonlinearities.py:15-22 */
enum {
    STPDE_ACT_TANH = 0,
    STPDE_ACT_RELU = 1,
    STPDE_ACT_SOFTPLUS = 2, /* torch.nn.Softplus(beta=1, threshold=20) */
    STPDE_ACT_ELU = 3,      /* alpha = 1 */
    STPDE_ACT_SWISH = 4,    /* x * sigmoid(act_param * x), act_param = learnable beta */
    STPDE_ACT_LEAKYRELU = 5 /* negative_slope = 0.01 */
};

/* arithmetic of the MLP contractions */
enum {
    STPDE_PREC_FP32 = 0,   /* FP32 FFMA (CUDA cores): reference-exact arithmetic */
    STPDE_PREC_FP16X3 = 1, /* tcgen05 tensor cores, fp16 hi/lo split operands, 3 MMAs per product,
                              fp32 accumulate: ~2^-22 operand precision (fp32 parity mode) */
    STPDE_PREC_FP16 = 2    /* tcgen05, single fp16 pass (relaxed parity, BASELINE config 3) */
};

/*
 * Problem descriptor.
 *
 * The decoder is the reference's skip-MLP (src/implicit_net.py:31-36) generalised to n_layers
 * linear layers of output width widths[l]:  D = dim + channels;
 *   layer 0        : in = D
 *   layer 1..n-2   : in = widths[l-1] + D      (input re-concatenated AFTER the activation)
 *   layer n-1      : in = widths[n-2]          (no activation; widths[n-1] = out_features)
 * ImNet(dim, c, o, nf) is n_layers = 6, widths = {16nf, 8nf, 4nf, 2nf, nf, o}.
 * Weight l is row-major [widths[l], in_l] (torch nn.Linear layout); input columns are ordered
 * [activations(widths[l-1]), x_relative(dim), latent(channels)].
 *
 * Jet specification: first_dirs[] lists the coordinate columns whose first derivative is
 * needed; second_pairs[][2] lists (i, j) column pairs whose second derivative is needed - both i
 * and j must appear in first_dirs.  1 + n_first + n_second <= STPDE_MAX_COMPONENTS per call.
 */
typedef struct stpde_desc {
    int32_t batch;                      /* b */
    int32_t npts;                       /* p, query points per batch element */
    int32_t dim;                        /* d, 1..4 */
    int32_t grid_size[STPDE_MAX_DIM];   /* n_1..n_d */
    int32_t channels;                   /* c */
    int32_t n_layers;
    int32_t widths[STPDE_MAX_LAYERS];
    int32_t act_kind;
    float act_param;
    int32_t n_first;
    int32_t first_dirs[STPDE_MAX_FIRST];
    int32_t n_second;
    int32_t second_pairs[STPDE_MAX_SECOND][2];
    int32_t precision;
    float xmin[STPDE_MAX_DIM];          /* float32 bounds exactly as the reference forms them */
    float xmax[STPDE_MAX_DIM];
    int32_t reserved[8];                /* [1]: stpde_jet_forward / stpde_jet_forward_train: 1 = the call-invariant setup in the
                                           workspace (split weights, per-vertex table) was built by the previous call for exactly
                                           these grid / weight values and is reused
                                           [2]: stpde_jet_backward with reuse_forward = 1: set to 1 when the forward that left
                                           the planes ran with precision STPDE_PREC_FP16 (its pre-activation planes are fp16)
                                           [0]: backward only, extra headroom bits of the adjoint scale (0 = default)
                                         * [1]: stpde_jet_forward only, 1 = the call-invariant part of the workspace (packed /
                                         *      split weights, per-vertex latent + bias table) is still valid from the
                                         *      previous call on the SAME workspace with the same decoder weights, latent
                                         *      grid, shapes and precision: the per-call setup kernels are skipped (the
                                         *      caller guarantees it; evaluation loops over pseudo-batches) */
} stpde_desc_t;

int stpde_version(void);
const char *stpde_last_error(void);
/* sizeof(stpde_desc_t) as compiled into the library (binding layout check). */
size_t stpde_desc_size(void);

/* Number of SMs / name of the device the library would run on (diagnostics; -1 if no device). */
int stpde_device_sm_count(void);

/* Scratch bytes stpde_jet_forward needs for this descriptor (0 on invalid descriptor). */
size_t stpde_workspace_bytes(const stpde_desc_t *desc);

/*
 * corner_values [b,p,2^d,c], weights [b,p,2^d], x_relative [b,p,2^d,d]  (all contiguous f32).
 * grid is [b, n_1..n_d, c] with element strides grid_strides[d+2]; q is [b,p,d] with q_strides[3].
 * status: device int32[1], OR-ed with 1 if any point addressed a cell outside the grid.
 */
int stpde_interp_coefficients(int32_t batch, int32_t npts, int32_t dim, const int32_t *grid_size,
                              int32_t channels, const float *grid, const int64_t *grid_strides,
                              const float *q, const int64_t *q_strides, const float *xmin,
                              const float *xmax, float *corner_values, float *weights,
                              float *x_relative, int32_t *status, void *stream);

/* out [b,p,c] = sum_j corner_j * weight_j */
int stpde_interp(int32_t batch, int32_t npts, int32_t dim, const int32_t *grid_size, int32_t channels,
                 const float *grid, const int64_t *grid_strides, const float *q,
                 const int64_t *q_strides, const float *xmin, const float *xmax, float *out,
                 int32_t *status, void *stream);

/*
 * Fused decode + jets.
 *   W[l], B[l] : device pointers of layer l's weight / bias (host array of n_layers pointers)
 *   y          : [b,p,o] values
 *   jets       : [n_first + n_second, b, p, o] partial derivatives w.r.t. the query coordinates
 *                (first-order planes in first_dirs order, then second-order planes); may be NULL
 *                when n_first == n_second == 0
 *   status     : device int32[1] (bit 0: index out of range, bit 1: fp16 range exceeded)
 */
int stpde_jet_forward(const stpde_desc_t *desc, const float *grid, const int64_t *grid_strides,
                      const float *q, const int64_t *q_strides, const float *const *W,
                      const float *const *B, float *y, float *jets, void *workspace,
                      size_t workspace_bytes, int32_t *status, void *stream);

/*
 * Reverse mode of stpde_jet_forward: given gy = d loss / d y [b,p,o] and gjets = d loss / d jets
 * [n_first + n_second, b, p, o] (contiguous; gjets may be NULL when no derivatives were requested), writes
 *   gW[l] : [widths[l], in_l]  gradient of layer l's weight      (host array of n_layers device pointers)
 *   gB[l] : [widths[l]]        gradient of layer l's bias
 *   ggrid : [b, n_1..n_d, c]   gradient of the latent grid, CONTIGUOUS (may be NULL)
 *   gbeta : [1]                gradient of act_param (the learnable beta of STPDE_ACT_SWISH, reference
 *                              src/nonlinearities.py:5-13); may be NULL, untouched for other activations
 * All outputs are overwritten.  The forward is recomputed per chunk of points (no activations are kept between
 * the forward and the backward call); the contractions run on the tensor cores with the fp16 hi/lo split
 * (STPDE_PREC_FP32 and STPDE_PREC_FP16X3: 3 passes, STPDE_PREC_FP16: 1 pass).  Needs n_layers >= 3.
 * Gradients w.r.t. the query points are not produced.
 * status bit 1 reports an adjoint that left the fp16 range.  The adjoints are rescaled by a power of two S derived
 * from max|gy|, max|gjets| so that the bound on the blended adjoints sits at 2^(10 - desc->reserved[0]); when the
 * flag comes back the caller repeats the call with reserved[0] += 6 (more headroom, less precision for tiny
 * adjoints).
 */
size_t stpde_backward_workspace_bytes(const stpde_desc_t *desc);
/* points per chunk the reverse-mode layout gets out of a workspace of that size (0: too small / unsupported) */
int64_t stpde_backward_chunk_points(const stpde_desc_t *desc, size_t workspace_bytes);
/*
 * reuse_forward = 1: the workspace still holds the planes of a stpde_jet_forward_train call with the SAME
 * descriptor, inputs and workspace (nothing else may have written to it in between); the recompute is skipped.
 */
int stpde_jet_backward(const stpde_desc_t *desc, const float *grid, const int64_t *grid_strides,
                       const float *q, const int64_t *q_strides, const float *const *W,
                       const float *const *B, const float *gy, const float *gjets, float *const *gW,
                       float *const *gB, float *ggrid, float *gbeta, void *workspace,
                       size_t workspace_bytes, int32_t reuse_forward, int32_t *status, void *stream);

/*
 * Training forward: same outputs as stpde_jet_forward (tensor-core arithmetic), but every operand plane and
 * pre-activation of the batch is left in the workspace (reverse-mode layout, stpde_backward_workspace_bytes) so that
 * stpde_jet_backward(reuse_forward = 1) does not have to recompute the forward.  The whole batch must fit ONE chunk
 * (stpde_backward_chunk_points(desc, workspace_bytes) >= batch * npts), otherwise STPDE_ENOMEM: use
 * stpde_jet_forward and let the backward recompute chunk by chunk.  This is the path a training step of the
 * reference's size takes (experiments/rb2d/train.py: a few thousand query points per step).
 */
int stpde_jet_forward_train(const stpde_desc_t *desc, const float *grid, const int64_t *grid_strides,
                            const float *q, const int64_t *q_strides, const float *const *W,
                            const float *const *B, float *y, float *jets, void *workspace,
                            size_t workspace_bytes, int32_t *status, void *stream);

/*
 * Same computation with HOST buffers (grid, q, weights, y, jets all in host memory; grid and q
 * contiguous).  Allocates device memory internally, copies in, runs, copies out, synchronises.
 * This is the end-to-end entry point a non-torch caller would use.
 */
int stpde_jet_forward_host(const stpde_desc_t *desc, const float *grid, const float *q,
                           const float *const *W, const float *const *B, float *y, float *jets);

/*
 * Residual evaluation: a postfix program per equation over the symbols
 *   [ q_0..q_{d-1} | y_0..y_{o-1} | jets plane 0..n_jet-1 for every output ]  evaluated per point.
 * prog is int32 words: (opcode, operand) pairs; consts are float32 literals.
 *   opcodes: 0 PUSH_CONST c   1 PUSH_Q k   2 PUSH_Y i   3 PUSH_JET (plane*o + i)
 *            4 ADD  5 MUL  6 NEG  7 POWI n  8 END (equation separator)
 * residuals: [n_eq, b, p].
 */
int stpde_residuals(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                    const float *q, const int64_t *q_strides, const float *y, const float *jets,
                    const int32_t *prog, int32_t prog_words, const float *consts, int32_t n_consts,
                    int32_t n_eq, float *residuals, void *stream);

/*
 * Reverse mode of stpde_residuals (what autograd does for the lambdified equation arithmetic of src/pde.py:139-142
 * inside loss.backward()).  The caller differentiates the equations symbolically and passes ONE postfix program per
 * output symbol - y_0..y_{o-1}, then every (jet plane, output) entry in plane-major order - evaluating
 *     sum_e gres[e] * d residual_e / d symbol ,   extra opcode  9 PUSH_GRES e   (gres: [n_eq, b, p])
 * Outputs: gy [b,p,o], gjets [n_jet,b,p,o] (all entries written).  Limits: 2048 words, 256 constants.
 */
int stpde_residuals_backward(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                             const float *q, const int64_t *q_strides, const float *y, const float *jets,
                             const int32_t *prog, int32_t prog_words, const float *consts, int32_t n_consts,
                             int32_t n_eq, const float *gres, float *gy, float *gjets, void *stream);

/*
 * Fused residual + loss reduction - replaces, for training, the lambdified residual arithmetic of
 * src/pde.py:139-142 TOGETHER with the loss reductions of experiments/rb2d/train.py:70-75
 *     reg_loss = loss_func(pred_value, point_value);  pde_loss = loss_func(stack(residues), 0)
 * loss_kind: 0 = l1 (F.l1_loss), 1 = l2 (F.mse_loss), 2 = huber (F.smooth_l1_loss, beta 1); train.py:31-39.
 * The residuals never reach memory: every CTA writes one pair of partial SUMS
 *     partial[block] = { sum l(y - target), sum over equations l(residual) }      (target may be NULL = zeros)
 * for stpde_residual_loss_blocks(batch * npts) blocks; the caller adds them (mean loss = sum / count, which is also
 * what makes one all-reduce of [sums | counts | gradients] equal the single-process means).
 */
int32_t stpde_residual_loss_blocks(int64_t total_points);
int stpde_residual_loss(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                        const float *q, const int64_t *q_strides, const float *y, const float *jets,
                        const float *target, const int32_t *prog, int32_t prog_words, const float *consts,
                        int32_t n_consts, int32_t n_eq, int32_t loss_kind, float *partial, void *stream);

/*
 * Reverse mode of stpde_residual_loss (the part of loss.backward(), train.py:77, between the two loss scalars and the
 * decoder outputs): g_sums = { d loss / d reg_sum, d loss / d pde_sum } (device pointer, 2 floats); the residuals are
 * recomputed, their cotangents g_pde * l'(r_e) stay in registers and feed the adjoint programs of
 * stpde_residuals_backward; gy additionally receives g_reg * l'(y - target).  Outputs gy [b,p,o], gjets [n_jet,b,p,o].
 */
int stpde_residual_loss_backward(int32_t batch, int32_t npts, int32_t dim, int32_t out_features, int32_t n_jet,
                                 const float *q, const int64_t *q_strides, const float *y, const float *jets,
                                 const float *target, const int32_t *prog, int32_t prog_words, const float *consts,
                                 int32_t n_consts, const int32_t *adj_prog, int32_t adj_words,
                                 const float *adj_consts, int32_t n_adj_consts, int32_t n_eq, int32_t loss_kind,
                                 const float *g_sums, float *gy, float *gjets, void *stream);

/*
 * Instrumentation (bench.py): every kernel launch is counted per slot; with profiling enabled each
 * launch group is additionally bracketed by CUDA events on the launching stream.
 * stpde_profile_read synchronises the device, fills elapsed milliseconds and launch counts per
 * slot since the previous read, resets them and returns the number of slots.
 */
int stpde_profile_enable(int on);
int stpde_profile_read(double *ms_by_slot, int64_t *launches_by_slot, int n_slots);
const char *stpde_profile_slot_name(int slot);

#ifdef __cplusplus
}
#endif
#endif /* STPDE_H_ */
