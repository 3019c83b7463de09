// Launch counting and optional per-kernel CUDA-event timing (bench.py's roofline leg).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stpde {

enum ProfSlot {
    kSlotSetup = 0,     // pack_weights + vertex_bias (+ tensor-core operand split)
    kSlotPrep = 1,      // prep_points
    kSlotLayer0 = 2,    // closed-form layer-0 jets
    kSlotGemm = 3,      // kSlotGemm + (l - 1) for hidden layer l = 1..7
    kSlotFinal = 10,    // last linear layer + blend
    kSlotResidual = 11, // residual programs
    kSlotBwdBlend = 12, // reverse: blend + last layer + last hidden activation
    kSlotBwdVertex = 13,// reverse: per-vertex adjoint -> latent columns, biases, grid; final 1/S
    kSlotWgrad = 16,    // reverse: weight-gradient contraction of hidden layer l at kSlotWgrad + (l - 1)
    kSlotDgrad = 24,    // reverse: activation-gradient contraction + reverse jet activation, kSlotDgrad + (l - 1)
    kNumSlots = 32
};

void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, cudaStream_t st, int n_launches);

struct ProfScope {
    int slot; cudaStream_t st; int n;
    ProfScope(int slot_, cudaStream_t st_, int n_launches = 1) : slot(slot_), st(st_), n(n_launches) { prof_begin(slot, st); }
    ~ProfScope() { prof_end(slot, st, n); }
};

}  // namespace stpde
