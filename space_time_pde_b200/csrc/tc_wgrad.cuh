// Weight-gradient contraction of one hidden layer on the tcgen05 tensor cores (CTA pair, cta_group::2).
//
//   gW_l[g][f] += sum_{c, r} zbar_l[c][r][g] * a_{l-1}[c][r][f]          (c = jet component, r = (point, corner) row)
//
// The contraction runs over the ROWS of the activation planes [KC][rows][features], i.e. both operands are
// "MN-major" for the MMA (the feature index is the contiguous one).  The TMA boxes are 64 features x 64 rows with
// SWIZZLE_128B - exactly the canonical MN-major SW128 atom (8 x 16-byte chunks along MN, 8 rows along K per 1024 B,
// SBO = 1024 B between 8-row groups, LBO = the distance between two 64-feature atoms) - so no transposed copy of
// the planes is ever written.
//
//   D[M = 256 features f of a_{l-1}, N = NT features g of zbar_l]   (fp32, TMEM, double buffered)
//   A = a_{l-1} planes (f is the TMEM lane => the epilogue's red.global.add is coalesced along a gW row)
//   B = zbar_l planes
//
// Work unit = (f tile, g tile, K slice); the K slices (split-K over rows x components) keep all 74 CTA pairs busy
// for the small output matrices, partial sums are added to gW with fp32 atomics.  Precision: the same fp16 hi/lo
// split as the forward (3 MMAs per product) or a single fp16 pass.
#pragma once
#include "tc_kernels.cuh"

namespace stpde {
namespace tc {

constexpr int kWgKBlock = 64;                       // rows (K) per pipeline stage
constexpr uint32_t kWgAtomBytes = kWgKBlock * 128;  // one 64-feature x 64-row box

struct WgradArgs {
    int rows, kc;          // K extent = kc * rows (rows % 64 == 0)
    int nf_a, nf_b;        // true feature counts: M extent (widths[l-1]) and N extent (widths[l])
    int nt;                // N tile: 128 or 256
    int n_ft, n_gt, n_slices;
    int tile_fastest;      // unit order: 1 = the tiles of one K slice are neighbours (concurrent CTA pairs share the slice's
                           // operand rows through L2), 0 = the slices of one tile are neighbours (round 1; STPDE_WGRAD_ORDER=0)
    int passes;
    float out_scale;       // 2^-4: the a planes are stored scaled by 2^4
    float* gW;             // [nf_b][ldw]
    int ldw;
    int* status;
};

// MN-major, 128B-swizzled operand: 64-element (128 B) rows along MN, one row per K index, 8-row atoms of 1024 B
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 64-element atoms along MN
    d |= (uint64_t)(1024 >> 4) << 32;                    // between 8-row groups along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor with both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_instr_desc_mn(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// work unit u -> (output tile, K slice)
__device__ __forceinline__ void wgrad_unit(const WgradArgs& a, int u, int& tile, int& slice) {
    const int n_tiles = a.n_ft * a.n_gt;
    if (a.tile_fastest) { tile = u % n_tiles; slice = u / n_tiles; }
    else                { slice = u % a.n_slices; tile = u / a.n_slices; }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_wgrad_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     WgradArgs args) {
    constexpr int kTileF2 = 2 * kTileF;
    constexpr uint32_t kTmemCols = 512;
    const bool three = args.passes == 3;
    const int nb_atoms = args.nt / 128;                              // 64-feature atoms of B per CTA
    const uint32_t a_bytes = 2 * kWgAtomBytes;                       // 128 features of A per CTA and plane
    const uint32_t b_bytes = (uint32_t)nb_atoms * kWgAtomBytes;
    const uint32_t stage_bytes = (three ? 2u : 1u) * (a_bytes + b_bytes);
    const int n_stages = min((int)(kPairSmemBudget / stage_bytes), kPairMaxStages);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + kPairSmemBudget);
    uint64_t* full_bar = bars;                                  // used in the leader only
    uint64_t* empty_bar = bars + kPairMaxStages;                // one copy per CTA
    uint64_t* tfull_bar = bars + 2 * kPairMaxStages;            // one copy per CTA
    uint64_t* tempty_bar = bars + 2 * kPairMaxStages + 2;       // used in the leader only
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kPairMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_units = args.n_ft * args.n_gt * args.n_slices;
    const int kb_per_comp = args.rows / kWgKBlock;
    const int64_t total_kb = (int64_t)args.kc * kb_per_comp;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
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
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int u = pair_id; u < n_units; u += n_pairs) {
            int tile, slice;
            wgrad_unit(args, u, tile, slice);
            const int f0 = (tile % args.n_ft) * kTileF2 + (int)rank * kTileF;
            const int g0 = (tile / args.n_ft) * args.nt + (int)rank * (args.nt / 2);
            const int64_t kb0 = total_kb * slice / args.n_slices, kb1 = total_kb * (slice + 1) / args.n_slices;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                const int c = (int)(kb / kb_per_comp), row0 = (int)(kb % kb_per_comp) * kWgKBlock;
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, args.status);
                const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
                const uint32_t base = smem_u32(smem + stage * stage_bytes);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(smem_u32(&full_bar[stage]), 2 * stage_bytes);
                    // layout of a stage: [A hi][A lo][B hi][B lo], each a run of 8 KB atoms (64 features each)
                    const uint32_t a_hi = base, a_lo = base + a_bytes;
                    const uint32_t b_hi = base + (three ? 2 * a_bytes : a_bytes), b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int at = 0; at < 2; ++at) {
                        tma_load_3d_pair(a_hi + at * kWgAtomBytes, &map_a_hi, f0 + at * 64, row0, c, fb);
                        if (three) tma_load_3d_pair(a_lo + at * kWgAtomBytes, &map_a_lo, f0 + at * 64, row0, c, fb);
                    }
                    for (int at = 0; at < nb_atoms; ++at) {
                        tma_load_3d_pair(b_hi + at * kWgAtomBytes, &map_b_hi, g0 + at * 64, row0, c, fb);
                        if (three) tma_load_3d_pair(b_lo + at * kWgAtomBytes, &map_b_lo, g0 + at * 64, row0, c, fb);
                    }
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            const uint32_t idesc = make_instr_desc_mn(kTileF2, args.nt);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int u = pair_id; u < n_units; u += n_pairs, ++it) {
                int tile, slice;
                wgrad_unit(args, u, tile, slice);
                const int64_t kb0 = total_kb * slice / args.n_slices, kb1 = total_kb * (slice + 1) / args.n_slices;
                const int buf = it & 1;
                mbar_wait(smem_u32(&tempty_bar[buf]), ((it >> 1) & 1) ^ 1, args.status);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int64_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase, args.status);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + stage * stage_bytes);
                    const uint32_t b_base = base + (three ? 2 * a_bytes : a_bytes);
                    const uint64_t a_hi = make_smem_desc_mn(base, kWgAtomBytes), a_lo = make_smem_desc_mn(base + a_bytes, kWgAtomBytes);
                    const uint64_t b_hi = make_smem_desc_mn(b_base, kWgAtomBytes), b_lo = make_smem_desc_mn(b_base + b_bytes, kWgAtomBytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kWgKBlock / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 16 * 128) >> 4);   // 16 K rows of 128 B
                            umma_f16_pair(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb > kb0 || k) ? 1u : 0u);
                            if (three) {
                                umma_f16_pair(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                umma_f16_pair(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                            }
                        }
                        umma_commit_pair(smem_u32(&empty_bar[stage]));
                        if (kb == kb1 - 1) umma_commit_pair(smem_u32(&tfull_bar[buf]));
                    }
                    __syncwarp();
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> scaled fp32 atomics into gW =====================
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;                         // 0..3: interleaved 8-column groups
        const uint32_t tempty_leader0 = map_to_cta(smem_u32(&tempty_bar[0]), 0);
        const uint32_t tempty_leader1 = map_to_cta(smem_u32(&tempty_bar[1]), 0);
        const int n_cg = args.nt / 8;
        int it = 0;
        for (int u = pair_id; u < n_units; u += n_pairs, ++it) {
            int tile, slice;
            wgrad_unit(args, u, tile, slice);
            const int64_t kb0 = total_kb * slice / args.n_slices, kb1 = total_kb * (slice + 1) / args.n_slices;
            const int buf = it & 1;
            const int f = (tile % args.n_ft) * kTileF2 + (int)rank * kTileF + quarter * 32 + lane;
            const int g0 = (tile / args.n_ft) * args.nt;
            mbar_wait(smem_u32(&tfull_bar[buf]), (it >> 1) & 1, args.status);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256;
            if (kb1 > kb0) {
#pragma unroll 1
                for (int cg = sub; cg < n_cg; cg += kEpiPerQuarter) {
                    uint32_t v[8];
                    tmem_ld_x8(taddr + cg * 8, v);
                    tmem_wait_ld();
                    if (f < args.nf_a) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int g = g0 + cg * 8 + j;
                            if (g < args.nf_b) atomicAdd(args.gW + (int64_t)g * args.ldw + f, __uint_as_float(v[j]) * args.out_scale);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(buf ? tempty_leader1 : tempty_leader0);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace stpde
