// tcgen05 / TMA / mbarrier primitives (inline PTX, sm_100a) and the tensor-core layer kernel.
//
// Layer l (1 <= l <= n-2) computes, for every jet component c and every (point, corner) row r,
//     z[c][r][g] = sum_f W_l[g][f] * a_{l-1}[c][r][f]
// as D[M = 128 features, N = KC * NR rows] = W_tile[128 x K] * Act_tile[N x K]^T on the 5th-gen
// tensor cores: the WEIGHTS are the M-side operand, so a TMEM lane is one output feature and the
// KC jet components of a row sit in neighbouring TMEM columns of the same lane - exactly what the
// jet-activation epilogue needs in one thread.
//
// Precision: operands are fp16 "hi + lo" pairs (x ~= hi + lo, 22 significant bits); a product is
// hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM (3 MMAs, STPDE_PREC_FP16X3) or hi*hi only
// (STPDE_PREC_FP16).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace stpde {
namespace tc {

constexpr int kBlockK = 64;     // fp16 elements per K block = 128 B = one swizzle-128B span
constexpr int kTileF = 128;     // features per CTA tile (UMMA M)
constexpr int kStages = 2;      // smem ring depth of the single-CTA kernel in the 3-pass mode (hi + lo planes per stage)
constexpr int kSingleMaxStages = 4;   // ... and in the single-pass mode (hi planes only: half the bytes per stage)
constexpr int kEpiWarps = 16;   // epilogue warps (4 per TMEM lane quarter, interleaved over the 8-row blocks)
constexpr int kEpiPerQuarter = kEpiWarps / 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, then epilogue
constexpr int kActScaleLog2 = 4;  // activations are stored as fp16(a * 2^4) (+ residual)
// Register budget: 18 warps = 5 warps on two of the four SM sub-partitions (16 K registers each) -> 96 registers per
// thread is the hardware limit of this block size (112 compiles but cannot launch: "too many resources").

// Output staging of the epilogue: every epilogue warp owns one smem buffer holding stage_rows(KC) rows x KC
// components x 32 features of BOTH fp16 planes (or of the fp32 plane of the last hidden layer) - 128 bytes per
// (row, component) - which leaves the SM through one cp.async.bulk.tensor store per plane.  Sized so that the
// operand ring of the single-CTA kernel (2 stages) and the staging of its 16 epilogue warps fit 227 KB.
#ifndef STPDE_EPI_NBUF
#define STPDE_EPI_NBUF 1        // staging buffers per epilogue warp (2 half-size buffers measured the same as 1)
#endif
#ifndef STPDE_EPI_SR_DIV
#define STPDE_EPI_SR_DIV 1      // divides the rows per staging pass (with STPDE_EPI_NBUF=2)
#endif
constexpr int kEpiBuffers = STPDE_EPI_NBUF;
constexpr uint32_t kRowScratch = 384;   // per epilogue warp: 2 slots of 8 rows x (x_0..x_3) + 8 vertex indices (160 B, padded to 192)
// single-CTA kernel (may fuse the final layer + blend): 2 slots of 384 B (+ blend factors, corner weights) + 2 x 96 B of
// per-point partial outputs exchanged between the four quarter-warps
constexpr uint32_t kRowScratchFused = 1024;
__host__ __device__ constexpr int stage_rows_base(int kc) {
    return kc == 1 ? 8 : (kc == 2 || kc == 3 || kc == 6 || kc == 9) ? 4 : kc == 8 ? 1 : 2;
}
__host__ __device__ constexpr int stage_rows(int kc) {
    return stage_rows_base(kc) / STPDE_EPI_SR_DIV > 0 ? stage_rows_base(kc) / STPDE_EPI_SR_DIV : 1;
}
// reverse mode: same staging granularity (a TMEM chunk never spans two staging passes: TL = min(rows per chunk, SR))
// K = 6 (the Rayleigh-Benard jet): 2 rows per pass, which leaves room for the z staging below
__host__ __device__ constexpr int stage_rows_bwd(int kc) { return kc == 6 ? 2 : stage_rows(kc); }
__host__ __device__ constexpr uint32_t epi_stage_bytes_bwd(int kc) { return (uint32_t)kc * stage_rows_bwd(kc) * 128u; }
// reverse mode, K = 6: the saved z planes of a 2-row chunk (K x 2 rows x 32 features, fp32) are copied into shared memory
// with cp.async while the PREVIOUS chunk is evaluated - loaded straight into registers, their L2 latency (~600 cycles,
// four times per 8-row item, with only 4 warps per scheduler to hide it) was the top stall of the dgrad kernels (ncu r02)
__host__ __device__ constexpr uint32_t z_stage_bytes(int kc) { return kc == 6 ? 6u * 2u * 128u : 0u; }
__host__ __device__ constexpr uint32_t epi_stage_bytes(int kc) {
    return (uint32_t)kc * stage_rows(kc) * 128u * kEpiBuffers;
}

// rows (point, corner) per tile for KC jet components: NR % 16 == 0 and KC * NR <= 256
__host__ __device__ constexpr int rows_per_tile(int kc) {
    return kc == 1 ? 256 : kc == 2 ? 128 : kc == 3 ? 80 : kc == 4 ? 64 : kc == 5 ? 48 : kc <= 8 ? 32 : 16;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait for a barrier phase.  try_wait carries a suspend-time hint so the warp SLEEPS in hardware until the
// phase completes (a hot spin loop here steals issue slots from the single TMA / MMA issuing threads).
// Bounded: a protocol bug must surface as a trapped kernel (error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* status, uint32_t hint_ns = 0x989680u) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(hint_ns)
            : "memory");
        if (done) return;
        if ((spin & 63u) == 63u) {          // wall-clock bound (profilers / time slicing can stretch a wait a lot)
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) break;   // 20 s
        }
    }
    atomicOr(status, 0x100);
    __trap();
}

// One lane of a fully converged warp (always the same one).  The single-thread tcgen05 / TMA instructions are
// issued under this predicate while the surrounding control flow stays warp-uniform, so descriptors and
// addresses live in uniform registers (a divergent "if (lane == 0)" region costs ~250 scalar fix-up
// instructions per K block and throttles the MMA issue rate).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Programmatic dependent launch (the host launches the tensor-core kernels with
// cudaLaunchAttributeProgrammaticStreamSerialization, tc_launch_ex): griddep_wait() returns when the previous kernel in the
// stream has completed and its writes are visible - everything a kernel does before it (barrier init, TMEM allocation,
// descriptor prefetch, cluster sync) overlaps that kernel's tail; griddep_launch_dependents() lets the NEXT kernel's CTAs
// take the SMs this grid's CTAs leave.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// smem -> global tensor store (bulk async group of the issuing thread); the box is clipped at the tensor bounds
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// shared-space accesses of the staging buffers (a generic pointer makes the compiler emit generic ST / LD)
__device__ __forceinline__ void sts_b16(uint32_t addr, __half v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// the same with a compile-time byte offset inside the instruction (st.shared [base + imm])
template <int OFF>
__device__ __forceinline__ void sts_b16_o(uint32_t base, __half v) {
    asm volatile("st.shared.b16 [%0+%2], %1;" ::"r"(base), "h"(__half_as_ushort(v)), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts_f32_o(uint32_t base, float v) {
    asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(base), "f"(v), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ uint4 lds_v4_o(uint32_t base) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base), "n"(OFF) : "memory");
    return v;
}
// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{})
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 operands, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
// 4-byte asynchronous copies global -> shared (LDGSTS): the reverse epilogue stages the z planes of its NEXT chunk
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* v) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8, "columns per load");
    if constexpr (N == 8) tmem_ld_x8(taddr, v);
    else if constexpr (N == 4) tmem_ld_x4(taddr, v);
    else if constexpr (N == 2) tmem_ld_x2(taddr, v);
    else tmem_ld_x1(taddr, v);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms of 1024 B
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) with SWIZZLE_128B = 2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=F32, A=B=F16, both K-major, dense
__host__ __device__ constexpr uint32_t make_instr_desc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct LayerArgs {
    int rows;          // (point, corner) rows in the chunk
    int n_feat;        // true output width of the layer
    int kp_in;         // padded K (multiple of 64)
    int ld_out;        // row stride of the output planes (fp16: next layer's kp_in; fp32: np)
    int n_store;       // features >= n_store are not written; [n_feat, n_store) are written as zeros
    int last;          // 1: write fp32 activations for final_blend, 0: write fp16 hi/lo planes
    int passes;        // 3 = hi/lo split, 1 = single fp16 pass
    int pack;          // single-CTA kernel, narrow layers: G = 2 / 4 row groups share one 128-lane tile (block-diagonal
                       // weight operand [128][G * kp_in]: group q's features sit in lanes q * 128/G ..., its K blocks in
                       // columns q * kp_in ...), so every TMEM lane quarter has an epilogue to run; 0 / 1 = off
    int dim, act, ncat, cat_off;
    float beta;
    const float* wscale;   // device: 2^-(sw_l + sa) for this layer
    const float* Wx;       // [n_feat][dim]
    const float* Vb;       // [nvert][ncat]
    const int* vtx;        // [rows]
    const float* xrel;     // [dim][rows]
    __half* out_hi;        // [KC][rows][ld_out]
    __half* out_lo;
    float* out_f32;        // [KC][rows][ld_out]
    int* status;
    // ---- training (save) / reverse-mode fields (tc_layer_pair_kernel MODE 1..3) ----
    float* z_out;          // MODE 1: pre-activations of this layer, fp32 [KC][rows][ldz]
    const float* z_in;     // MODE 2: pre-activations of the layer whose adjoints this launch produces
    int ldz;               // row stride of z_out / z_in
    int z_half;            // 1: z_out / z_in planes hold fp16 (training entirely in the single-pass mode: the saved
                           // pre-activations are 30 % of that step's HBM bytes as fp32), 0: fp32
    float* g_vb;           // MODE 2/3: adjoint of Vb, [nvert][ncat] (atomics)
    float* g_wx;           // MODE 2/3: adjoint of the coordinate columns, &gW[0][kh], row stride g_wx_ld (atomics)
    int g_wx_ld;
    float* g_beta;         // MODE 2/3: adjoint of the Swish beta (atomics), may be null
    uint32_t wait_ns;      // suspend-time hint of the mbarrier waits (pair kernel)
    // ---- last hidden layer of an inference call with the final linear layer + multilinear blend fused in (OUTK 3) ----
    int fuse_final;        // 1: this launch writes y / jets itself (no fp32 plane, no final_blend launch)
    int n_out;             // output features O (<= 4)
    int ldw_last;          // row stride of w_last
    int pc;                // points in the chunk
    long long p0, total_pts;   // first point of the chunk, points of the call
    const float* w_last;   // [O][ldw_last] last linear layer (padded rows)
    const float* b_last;   // [O]
    const float* wfac;     // [d][2][pc] blend factors, dfac [d][2][pc] their derivatives, dxr [d][pc] (ChunkBuffers)
    const float* dfac;
    const float* dxr;
    float* y;              // [total_pts][O]
    float* jets;           // [n_jet][total_pts][O]
    // TMA store maps of the output planes (dims (ld_out, rows, KC), box 32 features x stage_rows(KC) rows x KC):
    // [0] = fp16 hi plane, or the fp32 plane of the last hidden layer; [1] = fp16 lo plane (3-pass mode only)
    CUtensorMap out_map[2];
};

// Static jet specifications (template parameter SPEC): 0 = generic (runtime JetSpec), 1 = the Rayleigh-Benard set
// K = 6 with components [value | d/dq_0, d/dq_1, d/dq_2 | d2/dq_1^2, d2/dq_2^2] (parents of component 4 / 5 are
// component 2 / 3).  With the structure known at compile time the per-row epilogue code is branch-free, so the
// compiler interleaves the dependent chains of the 8 rows instead of running them one after the other.
constexpr int kSpecGeneric = 0;
constexpr int kSpecRb2 = 1;

// kernel modes of tc_layer_pair_kernel
constexpr int kModeFwd = 0;       // forward layer (inference)
constexpr int kModeFwdSave = 1;   // forward layer that also stores its pre-activations (recompute pass of the backward)
constexpr int kModeBwd = 2;       // dgrad of layer l (W_l^T . zbar_l) + reverse jet activation of layer l-1 >= 1
constexpr int kModeBwd0 = 3;      // dgrad of layer 1 + reverse of the closed-form layer 0

// Forward epilogue of one epilogue warp over ALL its tiles (shared by the CTA-pair and the single-CTA kernel):
// TMEM -> skip term + jet activation -> next layer's operand planes.
//   * thread = output feature g (TMEM lane); the warp's work items are the 8-row blocks rb = sub, sub + EPI_PQ, ... of
//     every tile of its CTA, in order;
//   * SOFTWARE PIPELINE over the items: the skip term of item n+1 (a gather of Vb[vertex] rows - the table does not fit
//     L2 for large latent grids, so its latency is a DRAM latency) is issued before item n is computed, and the
//     per-row operands of item n+2 (vertex index, cell-local coordinates: one coalesced load per lane) are in flight
//     at the same time; they pass through a 2 x 160-byte smem scratch and are read back as broadcasts.  Without this
//     every tile exposed one global-load latency to every warp (24 % of the stall samples);
//   * the accumulator buffer is handed back right after the last tcgen05.wait::ld of the warp - before any math;
//   * results go to the warp's smem staging buffer with immediate-offset stores (st.shared [base + imm]: no address
//     arithmetic per element), SR rows at a time (2 SR when only the hi plane is written), and leave through
//     cp.async.bulk.tensor (TMA) stores; the tensor map clips rows / features at the plane bounds.  (Measured against
//     two alternatives in round 2 - the warp draining its buffer with 16-byte copies, and direct stores: within 3 %.)
// OUTK: 0 = fp16 hi + lo planes (3-pass mode), 1 = fp16 hi plane only (single pass), 2 = fp32 plane (last hidden layer).
// tile(it, f0, r0): first feature / first row of this CTA's it-th tile, false past the end.
// hand_back(buf): gives TMEM accumulator buffer `buf` back to the MMA issuer.
template <int KC, int MODE, int SPEC, int NRB, int EPI_PQ, int OUTK, class TileFn, class HandBack>
__device__ __forceinline__ void fwd_epilogue_loop(const JetSpec& spec, const LayerArgs& args, uint32_t stg_addr,
                                                  uint32_t row_addr, int quarter, int sub, int lane, uint32_t tmem_q,
                                                  int n_cols, uint32_t tfull_addr0, TileFn&& tile, HandBack&& hand_back) {
    // scratch slot of one item: [0,128) x rows, [128,160) vertex indices, fused final only: [160,224) the point's blend
    // factors, [224,352) corner weights [8][w, dw/dq0, dw/dq1, dw/dq2]
    constexpr bool kFuse = OUTK >= 3;                 // 3: fused final layer only, 4: fused final layer AND the fp32 plane
    constexpr bool kF32 = OUTK == 2 || OUTK == 4;     // the staged output is the fp32 plane of the last hidden layer
    constexpr bool kStage = OUTK != 3;
    constexpr uint32_t kSlot = kFuse ? 384 : 192;
    constexpr int SRB = stage_rows(KC);
    constexpr int SR = (OUTK == 1 && SRB < 8) ? 2 * SRB : SRB;     // one plane only: twice the rows fit the buffer
    constexpr int NBUF = kEpiBuffers;
    constexpr int NPASS = 8 / SR;
    const bool has_blocks = sub < NRB;
    const float scale = __ldg(args.wscale);
    const int n_first = spec.n_first;

    // ---- per-feature constants, reloaded when the feature tile changes ----
    float w5[4] = {0.f, 0.f, 0.f, 0.f};                              // fused final layer: this feature's column of W_last
    int cur_f0 = -1, g = 0;
    bool g_ok = false, live = false, fuse_live = false;
    constexpr bool kRb2 = SPEC == kSpecRb2 && KC == 6;               // first-order components 1..3 <-> directions 0..2
    float sm = 0.f, wx[kMaxDim], wxc[kRb2 ? 1 : KC];                // wxc: constant tangent seed of first-order components
    auto load_feature_constants = [&](int f0) {
        cur_f0 = f0;
        const int fw = f0 + quarter * 32;
        g = fw + lane;
        g_ok = g < args.n_feat;
        live = fw < (OUTK == 3 ? args.n_feat : args.n_store) && has_blocks;
        fuse_live = fw < args.n_feat;
        if constexpr (kFuse) {
#pragma unroll
            for (int o = 0; o < 4; ++o) w5[o] = (g_ok && o < args.n_out) ? __ldg(args.w_last + o * args.ldw_last + g) : 0.f;
        }
        // pad features [n_feat, n_store) are written as zeros; fp16 planes carry a * 2^4
        sm = g_ok ? (OUTK >= 2 ? 1.f : (float)(1 << kActScaleLog2)) : 0.f;
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) wx[k] = (k < args.dim && g_ok) ? __ldg(args.Wx + g * args.dim + k) : 0.f;
        if constexpr (!kRb2) {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                wxc[c] = 0.f;
#pragma unroll
                for (int k = 0; k < kMaxDim; ++k)
                    if (spec.kind[c] == 1 && spec.dir[c] == k) wxc[c] = wx[k];
            }
        }
    };

    // ---- item sequence: (tile it, block rb) ----
    struct Item { int it, rb, f0, r0; bool ok; };
    auto first_item = [&]() {
        Item n{0, sub, 0, 0, false};
        n.ok = tile(0, n.f0, n.r0);
        return n;
    };
    auto next_item = [&](const Item& c) {
        Item n = c;
        if (has_blocks && c.rb + EPI_PQ < NRB) { n.rb = c.rb + EPI_PQ; return n; }
        n.it = c.it + 1;
        n.rb = sub;
        n.ok = tile(n.it, n.f0, n.r0);
        return n;
    };
    // row operands of an item: lane -> (coordinate plane k = lane / 8, row i = lane % 8); planes >= dim are zero
    auto load_rows = [&](const Item& m, float& xv, int& vv, float& pf) {
        const int rr = min(m.r0 + m.rb * 8 + (lane & 7), args.rows - 1);
        xv = __ldg(args.xrel + (int64_t)(lane >> 3) * args.rows + rr);
        vv = __ldg(args.vtx + rr);
        if constexpr (kFuse) {            // blend factors of the item's point: lanes 0..5 wfac, 6..11 dfac, 12..14 dxr
            const int ip = min((m.r0 + m.rb * 8) >> 3, args.pc - 1);
            const float* src = lane < 6 ? args.wfac + (int64_t)lane * args.pc
                             : lane < 12 ? args.dfac + (int64_t)(lane - 6) * args.pc
                                         : args.dxr + (int64_t)(lane - 12) * args.pc;
            pf = lane < 15 ? __ldg(src + ip) : 0.f;
        }
    };
    auto store_rows = [&](int slot, float xv, int vv, float pf) {  // -> scratch [8 rows][x_0..x_3] + [8] vertex indices
        const uint32_t a = row_addr + slot * kSlot;
        sts_f32(a + (lane & 7) * 16 + (lane >> 3) * 4, xv);
        if (lane < 8) sts_f32(a + 128 + lane * 4, __int_as_float(vv));
        if constexpr (kFuse) { if (lane < 16) sts_f32(a + 160 + lane * 4, pf); }
    };
    // Vb gather of an item (its rows are in scratch slot `slot`) for THIS thread's feature of that item's tile
    auto gather_vb = [&](const Item& m, int slot, float* zraw) {
        const int gm = m.f0 + quarter * 32 + lane;
        const bool ok = gm < args.n_feat;
        const float* vb = args.Vb + args.cat_off + (ok ? gm : 0);
        const uint32_t a = row_addr + slot * kSlot + 128;
        static_for<8>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const int vt = (int)lds_b32(a + i * 4);
            zraw[i] = ok ? __ldg(vb + (int64_t)vt * args.ncat) : 0.f;
        });
    };

    Item cur = first_item();
    if (!cur.ok) return;
    if (!has_blocks) {                                            // (K >= 9: fewer 8-row blocks than warps per quarter)
        for (int it = 0; cur.ok; ++it, cur.ok = tile(it, cur.f0, cur.r0)) {
            mbar_wait(tfull_addr0 + (it & 1) * 8, (it >> 1) & 1, args.status, args.wait_ns);
            hand_back(it & 1);
        }
        return;
    }
    // prologue: rows of item 0 -> slot 0, its gather; rows of item 1 -> slot 1
    float zs[8], zn[8];
    {
        float xv, pf = 0.f; int vv;
        load_rows(cur, xv, vv, pf);
        store_rows(0, xv, vv, pf);
    }
    Item nxt = next_item(cur);
    {
        float xv = 0.f, pf = 0.f; int vv = 0;
        if (nxt.ok) load_rows(nxt, xv, vv, pf);
        store_rows(1, xv, vv, pf);
    }
    __syncwarp();
    gather_vb(cur, 0, zn);
    float amax = 0.f;
    int slot = 0;                                                   // scratch slot holding the rows of `cur`
    int n_items = 0;

    while (cur.ok) {
        if (cur.f0 != cur_f0) load_feature_constants(cur.f0);
        // skip connection + per-vertex latent/bias term of the 8 rows of `cur` (gather issued one item ago)
        static_for<8>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const uint4 x = lds_v4(row_addr + slot * kSlot + i * 16);
            float z = zn[i];
            z = fmaf(wx[0], __uint_as_float(x.x), z);
            z = fmaf(wx[1], __uint_as_float(x.y), z);
            z = fmaf(wx[2], __uint_as_float(x.z), z);
            z = fmaf(wx[3], __uint_as_float(x.w), z);
            zs[i] = z;
        });
        // prefetch: gather of the next item (rows already in the other slot), row operands of the item after it
        float xv2 = 0.f, pf2 = 0.f; int vv2 = 0;
        if (nxt.ok) {
            gather_vb(nxt, slot ^ 1, zn);
            const Item nn = next_item(nxt);
            if (nn.ok) load_rows(nn, xv2, vv2, pf2);
        }

        const int buf = cur.it & 1;
        const uint32_t taddr = tmem_q + buf * n_cols;
        const bool first_of_tile = cur.rb == sub;
        const bool last_of_tile = !(has_blocks && cur.rb + EPI_PQ < NRB);
        if (first_of_tile) {
            mbar_wait(tfull_addr0 + buf * 8, (cur.it >> 1) & 1, args.status, args.wait_ns);
            tc_fence_after();
        }
        if (!live) {                                              // warp-uniform: no feature of this tile is ours
            if (last_of_tile) hand_back(buf);
        } else {
        const int rbase = cur.r0 + cur.rb * 8;
        const int fw = cur.f0 + quarter * 32;
        // fused final layer (OUTK 3): blend accumulators of this feature over the 8 corners of the item's point
        //   acc[0] = sum w a_0            acc[1..3] = sum dw/dq_k a_0      acc[4..6] = sum w a_{1+k}
        //   acc[7], acc[8] = sum dw/dq_1 a_2, sum dw/dq_2 a_3           acc[9], acc[10] = sum w a_4, sum w a_5
        // and the (lane-uniform) sums of the corner weights sw[0] = sum w, sw[1..3] = sum dw/dq_k for the bias terms
        float acc[11], sw[4];
        if constexpr (kFuse) {
#pragma unroll
            for (int e = 0; e < 11; ++e) acc[e] = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) sw[e] = 0.f;
            // corner weights of the point: lane -> (corner j = lane % 8, which = lane / 8: 0 weight, 1 + k d/dq_k)
            const uint32_t pa = row_addr + slot * kSlot + 160;
            const int j = lane & 7, which = lane >> 3;
            float wgt = 1.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = k * 2 + ((j >> (2 - k)) & 1);
                const float fk = __uint_as_float(lds_b32(pa + (which == k + 1 ? 24 : 0) + idx * 4));
                wgt = k == 0 ? fk : __fmul_rn(wgt, fk);           // torch.prod order
            }
            sts_f32(row_addr + slot * kSlot + 224 + (j * 4 + which) * 4, wgt);
            __syncwarp();
        }
        dispatch_act(args.act, [&](auto act_c) {
        constexpr int kAct = decltype(act_c)::value;
        static_for<NPASS>([&](auto PS) {
            constexpr int ps = decltype(PS)::value;
            constexpr uint32_t kBufOff = NBUF > 1 ? (ps % NBUF) * (KC * SRB * 128) : 0;
            [[maybe_unused]] const uint32_t sl = stg_addr + kBufOff + lane * (kF32 ? 4 : 2);
            // TL rows at a time leave TMEM (a whole 8-row block in registers does not fit the 96-register budget next to
            // the prefetched skip terms: spilling those made the spill store wait for the very load it was hiding)
            constexpr int TL = SR < 4 ? SR : 4;
            static_for<SR / TL>([&](auto CH) {
                constexpr int ch = decltype(CH)::value;
                uint32_t v[KC][TL];                            // row i of component c sits in accumulator column c * 8 + i
#pragma unroll
                for (int c = 0; c < KC; ++c) tmem_ld_n<TL>(taddr + cur.rb * (8 * KC) + c * 8 + ps * SR + ch * TL, v[c]);
                tmem_wait_ld();
                if (ps == NPASS - 1 && ch == SR / TL - 1 && last_of_tile) hand_back(buf);   // all values are in registers
                static_for<TL>([&](auto IT) {
                    constexpr int it_ = decltype(IT)::value;
                    constexpr int ir = ch * TL + it_;          // row inside the staging pass
                    constexpr int i = ps * SR + ir;            // row inside the 8-row block
                    float zt[KC], o[KC];
#pragma unroll
                    for (int c = 0; c < KC; ++c) {
                        float seed;
                        if constexpr (kRb2) seed = (c >= 1 && c <= 3) ? wx[c - 1] : 0.f;
                        else seed = wxc[c];
                        zt[c] = fmaf(__uint_as_float(v[c][it_]), scale, seed);
                    }
                    const float z0 = zt[0] + zs[i];
                    float s0, s1, s2;
                    act_jet_fast(kAct, args.beta, z0, s0, s1, s2);
                    if constexpr (MODE == kModeFwdSave) {
                        const int r = rbase + i;
                        if (g_ok && r < args.rows) {           // pre-activations for the reverse sweep
                            const int64_t zplane = (int64_t)args.rows * args.ldz;
                            if (args.z_half) {                 // (warp-uniform)
                                __half* pz = reinterpret_cast<__half*>(args.z_out) + (int64_t)r * args.ldz + g;
                                *pz = __float2half_rn(z0);
                                amax = fmaxf(amax, fabsf(z0));
#pragma unroll
                                for (int c = 1; c < KC; ++c) {
                                    pz += zplane;
                                    *pz = __float2half_rn(zt[c]);
                                    amax = fmaxf(amax, fabsf(zt[c]));      // range flag: the caller retries in fp16x3
                                }
                            } else {
                                float* pz = args.z_out + (int64_t)r * args.ldz + g;
                                *pz = z0;
#pragma unroll
                                for (int c = 1; c < KC; ++c) { pz += zplane; *pz = zt[c]; }
                            }
                        }
                    }
                    s0 *= sm; s1 *= sm; s2 *= sm;
                    o[0] = s0;
                    if constexpr (SPEC == kSpecRb2 && KC == 6) {
                        o[1] = s1 * zt[1]; o[2] = s1 * zt[2]; o[3] = s1 * zt[3];
                        o[4] = fmaf(s2 * zt[2], zt[2], s1 * zt[4]);
                        o[5] = fmaf(s2 * zt[3], zt[3], s1 * zt[5]);
                    } else {
#pragma unroll
                        for (int c = 1; c < KC; ++c) {
                            float oc = s1 * zt[c];
                            if (c > n_first) {               // second order (warp-uniform): parents za, zb
                                float za = 0.f, zb = 0.f;
#pragma unroll
                                for (int k = 0; k < STPDE_MAX_FIRST; ++k) {
                                    if (1 + k < KC) {
                                        za = fmaf(spec.sel_a[c][k], zt[1 + k], za);
                                        zb = fmaf(spec.sel_b[c][k], zt[1 + k], zb);
                                    }
                                }
                                oc = fmaf(s2 * za, zb, oc);
                            }
                            o[c] = oc;
                        }
                    }
                    if constexpr (kFuse) {
                        const uint4 cw = lds_v4(row_addr + slot * kSlot + 224 + i * 16);
                        const float w = __uint_as_float(cw.x), d0 = __uint_as_float(cw.y), d1 = __uint_as_float(cw.z),
                                    d2 = __uint_as_float(cw.w);
                        sw[0] += w; sw[1] += d0; sw[2] += d1; sw[3] += d2;
                        acc[0] = fmaf(w, o[0], acc[0]);
                        acc[1] = fmaf(d0, o[0], acc[1]); acc[2] = fmaf(d1, o[0], acc[2]); acc[3] = fmaf(d2, o[0], acc[3]);
                        acc[4] = fmaf(w, o[1], acc[4]); acc[5] = fmaf(w, o[2], acc[5]); acc[6] = fmaf(w, o[3], acc[6]);
                        acc[7] = fmaf(d1, o[2], acc[7]); acc[8] = fmaf(d2, o[3], acc[8]);
                        acc[9] = fmaf(w, o[4], acc[9]); acc[10] = fmaf(w, o[5], acc[10]);
                    }
                    if constexpr (kStage) {
                    if constexpr (ir == 0) {
                        // the TMA engine must have read the buffer's previous contents (a tile ago when the block is one pass)
                        if (lane == 0) { if constexpr (NBUF > 1) bulk_wait_read1(); else bulk_wait_read0(); }
                        __syncwarp();
                    }
                    static_for<KC>([&](auto C) {
                        constexpr int c = decltype(C)::value;
                        const float xs = o[c];
                        if constexpr (kF32) {
                            sts_f32_o<(c * SR + ir) * 128>(sl, xs);
                        } else {
                            amax = fmaxf(amax, fabsf(xs));
                            const __half hi = __float2half_rn(xs);
                            sts_b16_o<(c * SR + ir) * 64>(sl, hi);
                            if constexpr (OUTK == 0)
                                sts_b16_o<(KC * SR + c * SR + ir) * 64>(sl, __float2half_rn(xs - __half2float(hi)));
                        }
                    });
                    }
                });
            });
            if constexpr (kStage) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const uint32_t sb = stg_addr + kBufOff;
                    tma_store_3d(&args.out_map[0], sb, fw, rbase + ps * SR, 0);
                    if constexpr (OUTK == 0) tma_store_3d(&args.out_map[1], sb + KC * SR * 64, fw, rbase + ps * SR, 0);
                    bulk_commit();
                }
            }
        });
        });
        if constexpr (kFuse) {
            if (fuse_live) {
            // ---- product rule of the blend (reference local_implicit_grid.py:57-61 + the jets), per feature ----
            const uint32_t pa = row_addr + slot * kSlot + 160;
            const float dx0 = __uint_as_float(lds_b32(pa + 48)), dx1 = __uint_as_float(lds_b32(pa + 52)),
                        dx2 = __uint_as_float(lds_b32(pa + 56));
            float bl[6];
            bl[0] = acc[0];
            bl[1] = fmaf(dx0, acc[4], acc[1]);
            bl[2] = fmaf(dx1, acc[5], acc[2]);
            bl[3] = fmaf(dx2, acc[6], acc[3]);
            bl[4] = fmaf(2.f * dx1, acc[7], dx1 * dx1 * acc[9]);
            bl[5] = fmaf(2.f * dx2, acc[8], dx2 * dx2 * acc[10]);
            // ---- last linear layer: out[c][o] = sum_g W5[o][g] bl[c][g]: 24 values reduced over the 32 lanes so that
            //      lane L ends with value L = c * 4 + o (multi-value butterfly: 31 shuffles) ----
            float pv[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) pv[e] = e < 24 ? bl[e >> 2] * w5[e & 3] : 0.f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int e = 0; e < off; ++e) {
                    const float send = up ? pv[e] : pv[e + off];
                    const float keep = up ? pv[e + off] : pv[e];
                    pv[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            float total = pv[0];
            // quarters (32-feature groups) that hold real features add up in a fixed order through smem
            const int nq = min(4, (args.n_feat + 31) >> 5);
            if (nq > 1) {
                const uint32_t mine = row_addr + 768 + (n_items & 1) * 96;
                if (lane < 24) sts_f32(mine + lane * 4, total);
                named_bar_sync(1 + sub, 32 * nq);
                if (quarter == 0 && lane < 24) {
                    total = 0.f;
                    // the warps of one `sub` sit at scratch indices 4 sub + ((quarter + 2) & 3); this is quarter 0 (index + 2)
                    for (int qq = 0; qq < nq; ++qq)
                        total += __uint_as_float(lds_b32(mine + (((qq + 2) & 3) - 2) * (int)kRowScratchFused + lane * 4));
                }
            }
            const int ip = rbase >> 3;
            const long long gp = args.p0 + ip;
            const int c = lane >> 2, o = lane & 3;
            if (quarter == 0 && lane < 24 && o < args.n_out && ip < args.pc && gp < args.total_pts) {
                if (c < 4) total = fmaf(__ldg(args.b_last + o), sw[c], total);   // bias * sum of (d)weights
                if (c == 0) args.y[gp * args.n_out + o] = total;
                else args.jets[((long long)(c - 1) * args.total_pts + gp) * args.n_out + o] = total;
            }
            }
        }
        }
        // rows of the item after next -> the slot `cur` just vacated (every lane has read its x values above)
        __syncwarp();
        store_rows(slot, xv2, vv2, pf2);
        __syncwarp();
        slot ^= 1;
        ++n_items;
        cur = nxt;
        if (cur.ok) nxt = next_item(cur);
    }
    if (!(amax < 65000.f)) atomicOr(args.status, kStatusRange);
    if (lane == 0) bulk_wait0();                                    // staging must stay valid until the last store has read it
}

template <int KC, int MODE, int SPEC, int NRB, int EPI_PQ, bool CAN_FUSE, class TileFn, class HandBack>
__device__ __forceinline__ void fwd_epilogue(const JetSpec& spec, const LayerArgs& args, uint32_t stg, uint32_t rowbuf,
                                             int quarter, int sub, int lane, uint32_t tmem_q, int n_cols,
                                             uint32_t tfull_addr0, TileFn&& tile, HandBack&& hand_back) {
    if constexpr (CAN_FUSE) {
        if (args.last && args.fuse_final) {     // training forward (MODE 1): the reverse sweep still needs the fp32 plane
            fwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, MODE == kModeFwdSave ? 4 : 3>(spec, args, stg, rowbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
            return;
        }
    }
    if (args.last)
        fwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, 2>(spec, args, stg, rowbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
    else if (args.passes == 3)
        fwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, 0>(spec, args, stg, rowbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
    else
        fwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, 1>(spec, args, stg, rowbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
}

// Reverse-mode epilogue of one epilogue warp over ALL its tiles (MODE 2 / 3), same structure as fwd_epilogue_loop:
// software-pipelined row operands through a 2-slot smem scratch, the Vb gather of the NEXT item in flight (MODE 3
// recomputes z_0 of layer 0 from it), accumulator hand-back right after the last tcgen05.wait::ld, and - MODE 2 - the
// zbar planes staged in shared memory with immediate-offset stores and written by TMA.
// The accumulator holds S * 2^sw * abar[c][r][g]: the adjoint of the activations of the layer BELOW the contraction
// (feature g = TMEM lane).  MODE 2 turns it into the adjoint of that layer's pre-activations with the saved z planes
// and writes the next dgrad / wgrad operand planes; MODE 3 does the same for the closed-form layer 0 (nothing below it:
// only the parameter adjoints are accumulated).  The per-thread partial sums of the coordinate-column adjoint (G, A)
// and of the Swish beta adjoint live in registers across ALL tiles of a feature tile and are flushed with atomics once.
// OUTK: 0 = zbar hi + lo planes, 1 = hi plane only (single-pass reverse sweep), ignored in MODE 3.
template <int KC, int MODE, int SPEC, int NRB, int EPI_PQ, int OUTK, class TileFn, class HandBack>
__device__ __forceinline__ void bwd_epilogue_loop(const JetSpec& spec, const LayerArgs& args, uint32_t stg_addr,
                                                  uint32_t row_addr, uint32_t z_addr, int quarter, int sub, int lane, uint32_t tmem_q,
                                                  int n_cols, uint32_t tfull_addr0, TileFn&& tile, HandBack&& hand_back) {
    constexpr bool kBwd0 = MODE == kModeBwd0;
    constexpr uint32_t kSlot = 192;
    constexpr int SRB = stage_rows_bwd(KC);
    constexpr int SR = (OUTK == 1 && SRB < 8) ? 2 * SRB : SRB;      // rows per staging pass (MODE 2)
    constexpr int HMAX = (kBwd0 || KC <= 3) ? 4 : 2;               // rows per TMEM chunk (register budget)
    constexpr int TL = kBwd0 ? HMAX : (SR < HMAX ? SR : HMAX);
    constexpr int NPASS = kBwd0 ? 1 : 8 / SR;
    constexpr int CHUNKS = kBwd0 ? 8 / TL : SR / TL;               // chunks per pass
    constexpr bool ZST = !kBwd0 && z_stage_bytes(KC) > 0;          // z chunks staged through shared memory, one chunk ahead
    constexpr uint32_t kZLane = KC * TL * 4;                       // bytes per lane: the lane's K x TL values, contiguous
    static_assert(!ZST || (z_stage_bytes(KC) == 32 * kZLane && kZLane % 16 == 0), "z staging layout");
    const bool has_blocks = sub < NRB;
    const float scale = __ldg(args.wscale) * (float)(1 << kActScaleLog2);   // 2^-sw
    const int64_t zplane = (int64_t)args.rows * args.ldz;
    const int n_first = spec.n_first;
    const bool swish_beta_rt = args.act == STPDE_ACT_SWISH && args.g_beta != nullptr;

    // ---- per-feature state ----
    int cur_f0 = -1, g = 0;
    bool g_ok = false, live = false;
    float wx[kMaxDim], cf[KC];                              // MODE 3: layer-0 coordinate columns / jet coefficients
    float G[kMaxDim], A[KC], bsum = 0.f;                    // partial sums of the coordinate-column / beta adjoints
#pragma unroll
    for (int k = 0; k < kMaxDim; ++k) { G[k] = 0.f; wx[k] = 0.f; }
#pragma unroll
    for (int c = 0; c < KC; ++c) { A[c] = 0.f; cf[c] = 1.f; }
    auto flush_sums = [&]() {
        if (cur_f0 >= 0 && g_ok) {
#pragma unroll
            for (int c = 1; c < KC; ++c) {
#pragma unroll
                for (int k = 0; k < kMaxDim; ++k) {
                    if (spec.kind[c] == 1 && k == spec.dir[c]) G[k] += A[c];
                    if constexpr (kBwd0) {
                        if (spec.kind[c] == 2) {
                            const int da = spec.dir[spec.pa[c]], db = spec.dir[spec.pb[c]];
                            float wda = 0.f, wdb = 0.f;
#pragma unroll
                            for (int kk = 0; kk < kMaxDim; ++kk) { if (kk == da) wda = wx[kk]; if (kk == db) wdb = wx[kk]; }
                            if (k == da) G[k] = fmaf(A[c], wdb, G[k]);
                            if (k == db) G[k] = fmaf(A[c], wda, G[k]);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k)
                if (k < args.dim) atomicAdd(args.g_wx + (int64_t)g * args.g_wx_ld + k, G[k]);
        }
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) G[k] = 0.f;
#pragma unroll
        for (int c = 0; c < KC; ++c) A[c] = 0.f;
    };
    auto load_feature_constants = [&](int f0) {
        flush_sums();
        cur_f0 = f0;
        const int fw = f0 + quarter * 32;
        g = fw + lane;
        g_ok = g < args.n_feat;
        live = fw < args.n_store && has_blocks;
        if constexpr (kBwd0) {
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) wx[k] = (k < args.dim && g_ok) ? __ldg(args.Wx + g * args.dim + k) : 0.f;
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                float wa = 1.f, wb = 1.f;
#pragma unroll
                for (int k = 0; k < kMaxDim; ++k) {
                    if (spec.kind[c] == 1 && k == spec.dir[c]) wa = wx[k];
                    if (spec.kind[c] == 2 && k == spec.dir[spec.pa[c]]) wa = wx[k];
                    if (spec.kind[c] == 2 && k == spec.dir[spec.pb[c]]) wb = wx[k];
                }
                cf[c] = wa * wb;
            }
        }
    };

    // ---- item sequence (tile it, 8-row block rb) and the prefetch pipeline, as in fwd_epilogue_loop ----
    struct Item { int it, rb, f0, r0; bool ok; };
    auto next_item = [&](const Item& c) {
        Item n = c;
        if (has_blocks && c.rb + EPI_PQ < NRB) { n.rb = c.rb + EPI_PQ; return n; }
        n.it = c.it + 1;
        n.rb = sub;
        n.ok = tile(n.it, n.f0, n.r0);
        return n;
    };
    auto load_rows = [&](const Item& m, float& xv, int& vv) {
        const int rr = min(m.r0 + m.rb * 8 + (lane & 7), args.rows - 1);
        xv = __ldg(args.xrel + (int64_t)(lane >> 3) * args.rows + rr);
        vv = __ldg(args.vtx + rr);
    };
    auto store_rows = [&](int slot, float xv, int vv) {
        const uint32_t a = row_addr + slot * kSlot;
        sts_f32(a + (lane & 7) * 16 + (lane >> 3) * 4, xv);
        if (lane < 8) sts_f32(a + 128 + lane * 4, __int_as_float(vv));
    };
    auto gather_vb = [&](const Item& m, int slot, float* zraw) {       // MODE 3: Vb[vertex] of the item's 8 rows
        const int gm = m.f0 + quarter * 32 + lane;
        const bool ok = gm < args.n_feat;
        const float* vb = args.Vb + args.cat_off + (ok ? gm : 0);
        const uint32_t a = row_addr + slot * kSlot + 128;
        static_for<8>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const int vt = (int)lds_b32(a + i * 4);
            zraw[i] = ok ? __ldg(vb + (int64_t)vt * args.ncat) : 0.f;
        });
    };
    // MODE 2: the z planes stream from HBM (no reuse): pull the lines of an item into L2 ahead of time, one 128-byte
    // line (32 features of one row and component) per lane
    auto prefetch_z = [&](const Item& m) {
        const int gm0 = m.f0 + quarter * 32;
        if (gm0 < args.n_feat) {
            for (int idx = lane; idx < KC * 8; idx += 32) {
                const int rn = min(m.r0 + m.rb * 8 + (idx & 7), args.rows - 1);
                const int64_t e = (int64_t)(idx >> 3) * zplane + (int64_t)rn * args.ldz + gm0;
                prefetch_l2(args.z_half ? (const void*)(reinterpret_cast<const __half*>(args.z_in) + e) : (const void*)(args.z_in + e));
            }
        }
    };

    // MODE 2, K = 6: cp.async of the TL rows x K components of one chunk (first row `row0`, feature tile f0)
    // fp16 planes (args.z_half): a 4-byte copy holds the values of a lane PAIR (features 2k, 2k + 1); the pair shares one
    // kZLane-byte slot and each of its lanes issues every other (component, row) entry
    auto issue_z = [&](int f0, int row0) {
        const int gm = f0 + quarter * 32 + lane;
        if (args.z_half) {
            const int gp = gm & ~1;
            if (gp < args.n_feat) {
                const __half* src = reinterpret_cast<const __half*>(args.z_in) + gp;
                const uint32_t dst = z_addr + (lane >> 1) * kZLane;
                static_for<TL>([&](auto JT) {
                    constexpr int j = decltype(JT)::value;
                    const __half* srow = src + (int64_t)min(row0 + j, args.rows - 1) * args.ldz;
                    static_for<KC>([&](auto C) {
                        constexpr int c = decltype(C)::value;
                        if (((c * TL + j) & 1) == (lane & 1)) cp_async_4(dst + (c * TL + j) * 4, srow + (int64_t)c * zplane);
                    });
                });
            }
        } else if (gm < args.n_feat) {
            const float* src = args.z_in + gm;
            const uint32_t dst = z_addr + lane * kZLane;
            static_for<TL>([&](auto JT) {
                constexpr int j = decltype(JT)::value;
                const float* srow = src + (int64_t)min(row0 + j, args.rows - 1) * args.ldz;
                static_for<KC>([&](auto C) {
                    constexpr int c = decltype(C)::value;
                    cp_async_4(dst + (c * TL + j) * 4, srow + (int64_t)c * zplane);
                });
            });
        }
        cp_async_commit();
    };
    bool z_inflight = false;                                 // chunk 0 of `cur` was issued by the previous item

    Item cur{0, sub, 0, 0, false};
    cur.ok = tile(0, cur.f0, cur.r0);
    if (!cur.ok) return;
    if (!has_blocks) {
        for (int it = 0; cur.ok; ++it, cur.ok = tile(it, cur.f0, cur.r0)) {
            mbar_wait(tfull_addr0 + (it & 1) * 8, (it >> 1) & 1, args.status, args.wait_ns);
            hand_back(it & 1);
        }
        return;
    }
    float zn[kBwd0 ? 8 : 1], zs[kBwd0 ? 8 : 1];
    {
        float xv; int vv;
        load_rows(cur, xv, vv);
        store_rows(0, xv, vv);
    }
    Item nxt = next_item(cur);
    {
        float xv = 0.f; int vv = 0;
        if (nxt.ok) load_rows(nxt, xv, vv);
        store_rows(1, xv, vv);
    }
    __syncwarp();
    if constexpr (kBwd0) gather_vb(cur, 0, zn);
    float amax = 0.f;
    int slot = 0;

    while (cur.ok) {
        if (cur.f0 != cur_f0) load_feature_constants(cur.f0);
        const uint32_t rs = row_addr + slot * kSlot;
        if constexpr (kBwd0) {                                   // z_0 of layer 0 for the 8 rows (gather issued an item ago)
            static_for<8>([&](auto I) {
                constexpr int i = decltype(I)::value;
                const uint4 x = lds_v4(rs + i * 16);
                float z = zn[i];
                z = fmaf(wx[0], __uint_as_float(x.x), z);
                z = fmaf(wx[1], __uint_as_float(x.y), z);
                z = fmaf(wx[2], __uint_as_float(x.z), z);
                z = fmaf(wx[3], __uint_as_float(x.w), z);
                zs[i] = z;
            });
        }
        float xv2 = 0.f; int vv2 = 0;
        if (nxt.ok) {
            if constexpr (kBwd0) gather_vb(nxt, slot ^ 1, zn);
            else prefetch_z(nxt);
            const Item nn = next_item(nxt);
            if (nn.ok) load_rows(nn, xv2, vv2);
        }
        const int buf = cur.it & 1;
        const uint32_t taddr = tmem_q + buf * n_cols;
        const bool first_of_tile = cur.rb == sub;
        const bool last_of_tile = !(cur.rb + EPI_PQ < NRB);
        if (first_of_tile) {
            mbar_wait(tfull_addr0 + buf * 8, (cur.it >> 1) & 1, args.status, args.wait_ns);
            tc_fence_after();
        }
        if (!live) {
            if (last_of_tile) hand_back(buf);
        } else {
        const int rbase = cur.r0 + cur.rb * 8;
        const int fw = cur.f0 + quarter * 32;
        const int64_t zoff = (int64_t)rbase * args.ldz + (g_ok ? g : 0);
        const float* zbase = (kBwd0 || ZST) ? nullptr : args.z_in + zoff;
        const __half* zbase_h = (kBwd0 || ZST) ? nullptr : reinterpret_cast<const __half*>(args.z_in) + zoff;
        if constexpr (ZST) {
            if (!z_inflight) issue_z(cur.f0, rbase);
        }
        dispatch_act(args.act, [&](auto act_c) {
        constexpr int kAct = decltype(act_c)::value;
        static_for<NPASS>([&](auto PS) {
            constexpr int ps = decltype(PS)::value;
            const uint32_t sl = stg_addr + lane * 2;
            static_for<CHUNKS>([&](auto CH) {
                constexpr int ch = decltype(CH)::value;
                constexpr int i0 = (kBwd0 ? 0 : ps * SR) + ch * TL;       // first row of the chunk inside the 8-row block
                uint32_t v[KC][TL];
#pragma unroll
                for (int c = 0; c < KC; ++c) tmem_ld_n<TL>(taddr + cur.rb * (8 * KC) + c * 8 + i0, v[c]);
                float zc[kBwd0 ? 1 : KC][TL];
                if constexpr (ZST) {
                    cp_async_wait_all();                               // (fp32: each lane reads back only what it copied itself)
                    float zf[KC * TL];
                    if (args.z_half) {
                        __syncwarp();                                  // the pair's other lane copied half of the entries
                        static_for<KC * TL / 4>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            const uint4 t = lds_v4(z_addr + (lane >> 1) * kZLane + q * 16);
                            const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __half2 h2 = *reinterpret_cast<const __half2*>(&w4[e]);
                                zf[4 * q + e] = (lane & 1) ? __high2float(h2) : __low2float(h2);
                            }
                        });
                        __syncwarp();                                  // both lanes have read before the slot is refilled
                    } else {
                    static_for<KC * TL / 4>([&](auto Q) {
                        constexpr int q = decltype(Q)::value;
                        const uint4 t = lds_v4(z_addr + lane * kZLane + q * 16);
                        zf[4 * q] = __uint_as_float(t.x); zf[4 * q + 1] = __uint_as_float(t.y);
                        zf[4 * q + 2] = __uint_as_float(t.z); zf[4 * q + 3] = __uint_as_float(t.w);
                    });
                    }
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int j = 0; j < TL; ++j) zc[c][j] = g_ok ? zf[c * TL + j] : 0.f;
                    // the next chunk's copies go into the same buffer: they land a memory latency after the reads above
                    if constexpr (!(ps == NPASS - 1 && ch == CHUNKS - 1)) {
                        issue_z(cur.f0, rbase + i0 + TL);
                    } else {
                        z_inflight = nxt.ok && nxt.f0 + quarter * 32 < args.n_store;
                        if (z_inflight) issue_z(nxt.f0, nxt.r0 + nxt.rb * 8);
                    }
                } else if constexpr (!kBwd0) {
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int j = 0; j < TL; ++j) {
                            const int rr = min(rbase + i0 + j, args.rows - 1) - rbase;
                            const int64_t e = (int64_t)c * zplane + (int64_t)rr * args.ldz;
                            zc[c][j] = !g_ok ? 0.f : args.z_half ? __half2float(__ldg(zbase_h + e)) : __ldg(zbase + e);
                        }
                }
                tmem_wait_ld();
                if (ps == NPASS - 1 && ch == CHUNKS - 1 && last_of_tile) hand_back(buf);
                static_for<TL>([&](auto JT) {
                    constexpr int j = decltype(JT)::value;
                    constexpr int i = i0 + j;
                    constexpr int ir = i - ps * SR;                    // row inside the staging pass (MODE 2)
                    const int r = rbase + i;
                    const bool r_ok = r < args.rows;
                    const uint4 x = lds_v4(rs + i * 16);
                    const float xr[kMaxDim] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
                    const int vrow = (int)lds_b32(rs + 128 + i * 4);
                    float ab[KC];
#pragma unroll
                    for (int c = 0; c < KC; ++c) ab[c] = (g_ok && r_ok) ? __uint_as_float(v[c][j]) * scale : 0.f;
                    const float z0 = kBwd0 ? zs[kBwd0 ? i : 0] : zc[0][j];
                    float s1, s2, s3, z0b;
                    act_d123_fast(kAct, args.beta, z0, s1, s2, s3);
                    if constexpr (!kBwd0) {
                        float zb[KC];
                        float u = 0.f, w3 = 0.f;
                        if constexpr (SPEC == kSpecRb2 && KC == 6) {
#pragma unroll
                            for (int c = 1; c < KC; ++c) { u = fmaf(ab[c], zc[c][j], u); zb[c] = s1 * ab[c]; }
                            const float p4 = ab[4] * zc[2][j], p5 = ab[5] * zc[3][j];
                            w3 = fmaf(p4, zc[2][j], p5 * zc[3][j]);
                            zb[2] = fmaf(2.f * s2, p4, zb[2]);
                            zb[3] = fmaf(2.f * s2, p5, zb[3]);
                        } else {
                            float cross[STPDE_MAX_FIRST];
#pragma unroll
                            for (int k = 0; k < STPDE_MAX_FIRST; ++k) cross[k] = 0.f;
#pragma unroll
                            for (int c = 1; c < KC; ++c) {
                                u = fmaf(ab[c], zc[c][j], u);
                                zb[c] = s1 * ab[c];
                                if (c > n_first) {                 // second order (warp-uniform): parents za, zp
                                    float za = 0.f, zp = 0.f;
#pragma unroll
                                    for (int k = 0; k < STPDE_MAX_FIRST; ++k) {
                                        if (1 + k < KC) {
                                            za = fmaf(spec.sel_a[c][k], zc[1 + k][j], za);
                                            zp = fmaf(spec.sel_b[c][k], zc[1 + k][j], zp);
                                        }
                                    }
                                    w3 = fmaf(ab[c] * za, zp, w3);
#pragma unroll
                                    for (int k = 0; k < STPDE_MAX_FIRST; ++k)
                                        if (1 + k < KC) cross[k] = fmaf(ab[c], fmaf(spec.sel_a[c][k], zp, spec.sel_b[c][k] * za), cross[k]);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < STPDE_MAX_FIRST; ++k)
                                if (1 + k < KC) zb[1 + k] = fmaf(s2, cross[k], zb[1 + k]);
                        }
                        z0b = fmaf(s1, ab[0], fmaf(s2, u, s3 * w3));
                        zb[0] = z0b;
                        if (kAct == STPDE_ACT_SWISH && swish_beta_rt) {
                            float sb0, sb1, sb2;
                            swish_dbeta(args.beta, z0, sb0, sb1, sb2);
                            bsum += fmaf(ab[0], sb0, fmaf(sb1, u, sb2 * w3));
                        }
#pragma unroll
                        for (int c = 1; c < KC; ++c) A[c] += zb[c];
                        if constexpr (ir == 0) {                   // the TMA engine has read the buffer's previous contents
                            if (lane == 0) bulk_wait_read0();
                            __syncwarp();
                        }
                        static_for<KC>([&](auto C) {
                            constexpr int c = decltype(C)::value;
                            const float xs = zb[c];
                            amax = fmaxf(amax, fabsf(xs));
                            const __half hi = __float2half_rn(xs);
                            sts_b16_o<(c * SR + ir) * 64>(sl, hi);
                            if constexpr (OUTK == 0)
                                sts_b16_o<(KC * SR + c * SR + ir) * 64>(sl, __float2half_rn(xs - __half2float(hi)));
                        });
                    } else {
                        // layer 0: a_c = sigma^(order_c)(z0) * cf_c
                        float t1 = 0.f, t2 = 0.f;
                        if constexpr (SPEC == kSpecRb2 && KC == 6) {
                            t1 = fmaf(ab[1], cf[1], fmaf(ab[2], cf[2], ab[3] * cf[3]));
                            t2 = fmaf(ab[4], cf[4], ab[5] * cf[5]);
#pragma unroll
                            for (int c = 1; c < 4; ++c) A[c] = fmaf(ab[c], s1, A[c]);
                            A[4] = fmaf(ab[4], s2, A[4]);
                            A[5] = fmaf(ab[5], s2, A[5]);
                        } else {
#pragma unroll
                            for (int c = 1; c < KC; ++c) {
                                const float pc = ab[c] * cf[c];
                                if (c <= n_first) { t1 += pc; A[c] = fmaf(ab[c], s1, A[c]); }
                                else { t2 += pc; A[c] = fmaf(ab[c], s2, A[c]); }
                            }
                        }
                        z0b = fmaf(s1, ab[0], fmaf(s2, t1, s3 * t2));
                        if (kAct == STPDE_ACT_SWISH && swish_beta_rt) {
                            float sb0, sb1, sb2;
                            swish_dbeta(args.beta, z0, sb0, sb1, sb2);
                            bsum += fmaf(ab[0], sb0, fmaf(sb1, t1, sb2 * t2));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kMaxDim; ++k) G[k] = fmaf(z0b, xr[k], G[k]);
                    // (exact zeros - padding rows, dead relu units - need no atomic)
                    if (g_ok && r_ok && z0b != 0.f && args.g_vb) atomicAdd(args.g_vb + (int64_t)vrow * args.ncat + args.cat_off + g, z0b);
                });
            });
            if constexpr (!kBwd0) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&args.out_map[0], stg_addr, fw, rbase + ps * SR, 0);
                    if constexpr (OUTK == 0) tma_store_3d(&args.out_map[1], stg_addr + KC * SR * 64, fw, rbase + ps * SR, 0);
                    bulk_commit();
                }
            }
        });
        });
        }
        __syncwarp();
        store_rows(slot, xv2, vv2);
        __syncwarp();
        slot ^= 1;
        cur = nxt;
        if (cur.ok) nxt = next_item(cur);
    }
    flush_sums();
    if (swish_beta_rt) {
        for (int off = 16; off > 0; off >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, off);
        if (lane == 0) atomicAdd(args.g_beta, bsum);
    }
    if (!(amax < 65000.f)) atomicOr(args.status, kStatusRange);
    if (!kBwd0 && lane == 0) bulk_wait0();
}

template <int KC, int MODE, int SPEC, int NRB, int EPI_PQ, class TileFn, class HandBack>
__device__ __forceinline__ void bwd_epilogue(const JetSpec& spec, const LayerArgs& args, uint32_t stg, uint32_t rowbuf,
                                             uint32_t zbuf, int quarter, int sub, int lane, uint32_t tmem_q, int n_cols,
                                             uint32_t tfull_addr0, TileFn&& tile, HandBack&& hand_back) {
    if (MODE == kModeBwd0 || args.passes == 3)
        bwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, 0>(spec, args, stg, rowbuf, zbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
    else
        bwd_epilogue_loop<KC, MODE, SPEC, NRB, EPI_PQ, 1>(spec, args, stg, rowbuf, zbuf, quarter, sub, lane, tmem_q, n_cols, tfull_addr0, tile, hand_back);
}

// bytes of output staging (+ row scratch, + z staging in reverse mode) a kernel mode needs (all 16 epilogue warps)
template <int KC, int MODE, bool SINGLE = false>
__host__ __device__ constexpr uint32_t epi_staging_total() {
    return MODE < kModeBwd ? (uint32_t)kEpiWarps * (epi_stage_bytes(KC) + (SINGLE ? kRowScratchFused : kRowScratch))
           : MODE == kModeBwd ? (uint32_t)kEpiWarps * (epi_stage_bytes_bwd(KC) + kRowScratch + z_stage_bytes(KC))
                              : (uint32_t)kEpiWarps * kRowScratch;           // MODE 3 writes no planes
}

template <int KC, int SPEC = 0, int MODE = kModeFwd>
__global__ void __launch_bounds__(kThreads, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ JetSpec spec, const __grid_constant__ LayerArgs args) {
    constexpr int NR = rows_per_tile(KC);
    constexpr int N = KC * NR;
    constexpr int NRB = NR / 8;
    constexpr uint32_t kWBytes = kTileF * kBlockK * 2;        // one W plane tile
    constexpr uint32_t kABytes = N * kBlockK * 2;             // one activation plane tile
    constexpr uint32_t kRing = kStages * (2 * kWBytes + 2 * kABytes);   // operand ring: 2 stages of hi + lo planes ...
    constexpr uint32_t kStaging = epi_staging_total<KC, MODE, true>();
    constexpr uint32_t kTmemCols = 512;
    const bool three = args.passes == 3;
    // ... or 4 stages when only the hi planes are loaded (single-pass mode): twice the bytes in flight per SM
    const uint32_t stage_bytes = three ? 2 * kWBytes + 2 * kABytes : kWBytes + kABytes;
    const int n_stages = min((int)(kRing / stage_bytes), kSingleMaxStages);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* staging = smem + kRing;
    uint64_t* bars = (uint64_t*)(staging + kStaging);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kSingleMaxStages;
    uint64_t* tfull_bar = bars + 2 * kSingleMaxStages;
    uint64_t* tempty_bar = bars + 2 * kSingleMaxStages + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kSingleMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Narrow layers (forward: <= 64 features; dgrad MODE 2: <= 64 features of the layer below): G row groups of NR rows per tile.  The contraction of group q runs over K blocks
    // [q * kb_count, (q + 1) * kb_count) of the block-diagonal weight operand, whose only non-zero rows there are the
    // lanes of group q - the other groups' accumulator lanes receive exact zeros from those blocks.
    const int G = args.pack > 1 ? args.pack : 1;
    const int n_ftiles = G > 1 ? 1 : (args.n_store + kTileF - 1) / kTileF;
    const int n_rtiles = (args.rows + G * NR - 1) / (G * NR);
    const int n_tiles = n_ftiles * n_rtiles;
    const int kb_count = args.kp_in / kBlockK;
    const int kb_total = G * kb_count;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        if (MODE < kModeBwd) { tma_prefetch_desc(&args.out_map[0]); tma_prefetch_desc(&args.out_map[1]); }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kSingleMaxStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), kEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();          // no global memory is read or written above this line

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int f0 = (t % n_ftiles) * kTileF, r0 = (t / n_ftiles) * (G * NR);
            for (int kq = 0, kb = 0, rq = r0; kq < kb_total; ++kq) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, args.status);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t base = smem_u32(smem + stage * stage_bytes);
                const uint32_t a_base = base + (three ? 2 * kWBytes : kWBytes);
                if (elect_one()) {
                    mbar_expect_tx(fb, stage_bytes);
                    tma_load_2d(base, &map_w_hi, kq * kBlockK, f0, fb);
                    if (three) tma_load_2d(base + kWBytes, &map_w_lo, kq * kBlockK, f0, fb);
#pragma unroll
                    for (int rb = 0; rb < NRB; ++rb) {
                        tma_load_3d(a_base + rb * (8 * KC * 128), &map_a_hi, kb * kBlockK, rq + rb * 8, 0, fb);
                        if (three)
                            tma_load_3d(a_base + kABytes + rb * (8 * KC * 128), &map_a_lo, kb * kBlockK, rq + rb * 8, 0, fb);
                    }
                }
                if (++kb == kb_count) { kb = 0; rq += NR; }      // next row group of the tile
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        constexpr uint32_t idesc = make_instr_desc(kTileF, N);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(smem_u32(&tempty_bar[buf]), ((it >> 1) & 1) ^ 1, args.status);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * N;
            for (int kb = 0; kb < kb_total; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase, args.status);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + stage * stage_bytes);
                const uint32_t a_base = base + (three ? 2 * kWBytes : kWBytes);
                const uint64_t w_hi = make_smem_desc(base), w_lo = make_smem_desc(base + kWBytes);
                const uint64_t a_hi = make_smem_desc(a_base), a_lo = make_smem_desc(a_base + kABytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);   // +32 B along K inside the swizzle span
                        umma_f16(d_tmem, w_hi + adv, a_hi + adv, idesc, (kb | k) ? 1u : 0u);
                        if (three) {
                            umma_f16(d_tmem, w_hi + adv, a_lo + adv, idesc, 1u);
                            umma_f16(d_tmem, w_lo + adv, a_hi + adv, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));          // frees the smem stage when the MMAs retire
                    if (kb == kb_total - 1) umma_commit(smem_u32(&tfull_bar[buf]));
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        // Warp w may only touch TMEM lanes 32*(w%4)..+31; the warps of a quarter take the 8-row blocks in turn.
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int qpg = 4 / G;                               // lane quarters per row group
        const int grp = quarter / qpg, fq = quarter - grp * qpg;   // row group / 32-feature block of this warp
        const uint32_t tmem_q = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // TMEM hand-back: the tcgen05.ld results are in registers (tcgen05.wait::ld), ordered before the arrive
        auto hand_back = [&](int buf) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        };
        if constexpr (MODE >= kModeBwd) {
            auto tile = [&](int it, int& f0, int& r0) {
                const int t = blockIdx.x + it * gridDim.x;
                f0 = (t % n_ftiles) * kTileF;
                r0 = (t / n_ftiles) * (G * NR) + grp * NR;
                return t < n_tiles;
            };
            constexpr uint32_t kStg = MODE == kModeBwd ? epi_stage_bytes_bwd(KC) : 0u;
            bwd_epilogue<KC, MODE, SPEC, NRB, kEpiPerQuarter>(spec, args, smem_u32(staging) + (warp - 2) * kStg,
                                                              smem_u32(staging) + kEpiWarps * kStg + (warp - 2) * kRowScratch,
                                                              smem_u32(staging) + kEpiWarps * (kStg + kRowScratch) + (warp - 2) * z_stage_bytes(KC),
                                                              fq, sub, lane, tmem_q, N, smem_u32(&tfull_bar[0]), tile, hand_back);
        } else {
            auto tile = [&](int it, int& f0, int& r0) {
                const int t = blockIdx.x + it * gridDim.x;
                f0 = (t % n_ftiles) * kTileF;
                r0 = (t / n_ftiles) * (G * NR) + grp * NR;
                return t < n_tiles;
            };
            constexpr bool kCanFuse = MODE < kModeBwd && KC == 6 && SPEC == kSpecRb2;
            fwd_epilogue<KC, MODE, SPEC, NRB, kEpiPerQuarter, kCanFuse>(spec, args, smem_u32(staging) + (warp - 2) * epi_stage_bytes(KC),
                                                              smem_u32(staging) + kEpiWarps * epi_stage_bytes(KC) + (warp - 2) * kRowScratchFused,
                                                              fq, sub, lane, tmem_q, N, smem_u32(&tfull_bar[0]), tile, hand_back);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}


// =============================================================================================
// CTA-pair version (cta_group::2): two SMs of a cluster share one 256-feature x N tile.
//   * each CTA TMA-loads its own 128 weight rows and its own HALF of the activation rows, so the
//     per-SM operand traffic and smem footprint drop (3+ pipeline stages instead of 2);
//   * the leader CTA (cluster rank 0) issues every tcgen05.mma.cta_group::2 (M = 256); the
//     accumulator rows 0..127 land in the leader's TMEM, rows 128..255 in the peer's;
//   * full barriers live in the leader (both CTAs' TMA traffic completes on them), empty / tmem-full
//     barriers are multicast to both CTAs by tcgen05.commit, tmem-empty arrivals go to the leader.
// =============================================================================================
constexpr int kPairMaxStages = 8;
constexpr uint32_t kPairSmemBudget = 224 * 1024;            // operand ring + output staging

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// arrive on a barrier of the peer CTA with the default semantics (.release at CTA scope: no device-wide drain)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int KC, int MODE = kModeFwd, int SPEC = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_layer_pair_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ JetSpec spec, const __grid_constant__ LayerArgs args) {
    constexpr int NR = rows_per_tile(KC);
    constexpr int N = KC * NR;
    constexpr int NRB = NR / 8;
    static_assert(NRB % 2 == 0, "each CTA of the pair loads half of the 8-row blocks");
    constexpr int kTileF2 = 2 * kTileF;
    constexpr uint32_t kWBytes = kTileF * kBlockK * 2;          // one W plane tile of this CTA (128 rows)
    constexpr uint32_t kABytes = (N / 2) * kBlockK * 2;         // this CTA's half of one activation plane tile
    constexpr uint32_t kStaging = epi_staging_total<KC, MODE>();
    constexpr uint32_t kRingBudget = kPairSmemBudget - kStaging;
    constexpr uint32_t kTmemCols = 512;
    const bool three = args.passes == 3;
    const uint32_t stage_bytes = three ? 2 * kWBytes + 2 * kABytes : kWBytes + kABytes;
    const int n_stages = min((int)(kRingBudget / stage_bytes), kPairMaxStages);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* staging = smem + kRingBudget;
    uint64_t* bars = (uint64_t*)(smem + kPairSmemBudget);
    uint64_t* full_bar = bars;                                  // used in the leader only
    uint64_t* empty_bar = bars + kPairMaxStages;                // one copy per CTA
    uint64_t* tfull_bar = bars + 2 * kPairMaxStages;            // one copy per CTA
    uint64_t* tempty_bar = bars + 2 * kPairMaxStages + 2;       // used in the leader only
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kPairMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_ftiles = (args.n_store + kTileF2 - 1) / kTileF2;
    const int n_rtiles = (args.rows + NR - 1) / NR;
    const int n_tiles = n_ftiles * n_rtiles;
    const int kb_count = args.kp_in / kBlockK;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        if (MODE < kModeBwd) { tma_prefetch_desc(&args.out_map[0]); tma_prefetch_desc(&args.out_map[1]); }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kPairMaxStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 2 * kEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();          // no global memory is read or written above this line

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop; one elected lane per CTA issues) =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = pair_id; t < n_tiles; t += n_pairs) {
            const int f0 = (t % n_ftiles) * kTileF2 + (int)rank * kTileF;
            const int r0 = (t / n_ftiles) * NR + (int)rank * (NR / 2);
            for (int kb = 0; kb < kb_count; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, args.status, args.wait_ns);
                const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);   // the leader's copy of the barrier
                const uint32_t base = smem_u32(smem + stage * stage_bytes);
                const uint32_t a_base = base + (three ? 2 * kWBytes : kWBytes);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(smem_u32(&full_bar[stage]), 2 * stage_bytes);   // bytes of BOTH CTAs
                    tma_load_2d_pair(base, &map_w_hi, kb * kBlockK, f0, fb);
                    if (three) tma_load_2d_pair(base + kWBytes, &map_w_lo, kb * kBlockK, f0, fb);
#pragma unroll
                    for (int rb = 0; rb < NRB / 2; ++rb) {
                        tma_load_3d_pair(a_base + rb * (8 * KC * 128), &map_a_hi, kb * kBlockK, r0 + rb * 8, 0, fb);
                        if (three)
                            tma_load_3d_pair(a_base + kABytes + rb * (8 * KC * 128), &map_a_lo, kb * kBlockK, r0 + rb * 8, 0, fb);
                    }
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: leader CTA, warp-uniform loop, one elected lane issues =====================
        if (leader) {
            constexpr uint32_t idesc = make_instr_desc(kTileF2, N);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = pair_id; t < n_tiles; t += n_pairs, ++it) {
                const int buf = it & 1;
                mbar_wait(smem_u32(&tempty_bar[buf]), ((it >> 1) & 1) ^ 1, args.status, args.wait_ns);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * N;
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase, args.status, args.wait_ns);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + stage * stage_bytes);
                    const uint32_t a_base = base + (three ? 2 * kWBytes : kWBytes);
                    const uint64_t w_hi = make_smem_desc(base), w_lo = make_smem_desc(base + kWBytes);
                    const uint64_t a_hi = make_smem_desc(a_base), a_lo = make_smem_desc(a_base + kABytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_f16_pair(d_tmem, w_hi + adv, a_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            if (three) {
                                umma_f16_pair(d_tmem, w_hi + adv, a_lo + adv, idesc, 1u);
                                umma_f16_pair(d_tmem, w_lo + adv, a_hi + adv, idesc, 1u);
                            }
                        }
                        umma_commit_pair(smem_u32(&empty_bar[stage]));
                        if (kb == kb_count - 1) umma_commit_pair(smem_u32(&tfull_bar[buf]));
                    }
                    __syncwarp();
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs): this CTA's 128 features x all N columns =====================
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const uint32_t tmem_q = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t tempty_leader0 = map_to_cta(smem_u32(&tempty_bar[0]), 0);
        const uint32_t tempty_leader1 = map_to_cta(smem_u32(&tempty_bar[1]), 0);
        // TMEM hand-back to the leader's barrier: the tcgen05.ld results are in registers (tcgen05.wait::ld +
        // tcgen05.fence::before_thread_sync) and the arrive carries the default .release.cta semantics - the
        // cluster-scope release measured 20 % of this warp's stall samples (MEMBAR + ERRBAR drain every outstanding
        // global access, including the prefetches issued on purpose just before).
        auto hand_back = [&](int buf) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(buf ? tempty_leader1 : tempty_leader0);
        };
        auto tile = [&](int it, int& f0, int& r0) {
            const int t = pair_id + it * n_pairs;
            f0 = (t % n_ftiles) * kTileF2 + (int)rank * kTileF;
            r0 = (t / n_ftiles) * NR;
            return t < n_tiles;
        };
        if constexpr (MODE >= kModeBwd) {
            constexpr uint32_t kStg = MODE == kModeBwd ? epi_stage_bytes_bwd(KC) : 0u;
            bwd_epilogue<KC, MODE, SPEC, NRB, kEpiPerQuarter>(spec, args, smem_u32(staging) + (warp - 2) * kStg,
                                                              smem_u32(staging) + kEpiWarps * kStg + (warp - 2) * kRowScratch,
                                                              smem_u32(staging) + kEpiWarps * (kStg + kRowScratch) + (warp - 2) * z_stage_bytes(KC),
                                                              quarter, sub, lane, tmem_q, N, smem_u32(&tfull_bar[0]), tile, hand_back);
        } else {
            fwd_epilogue<KC, MODE, SPEC, NRB, kEpiPerQuarter, false>(spec, args, smem_u32(staging) + (warp - 2) * epi_stage_bytes(KC),
                                                              smem_u32(staging) + kEpiWarps * epi_stage_bytes(KC) + (warp - 2) * kRowScratch,
                                                              quarter, sub, lane, tmem_q, N, smem_u32(&tfull_bar[0]), tile, hand_back);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace stpde
