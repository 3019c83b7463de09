#include "profile.h"

#include <mutex>
#include <vector>

#include "../../include/stpde.h"

namespace stpde {

namespace {
struct Pair { cudaEvent_t a, b; int slot; };
std::mutex g_mu;
bool g_enabled = false;
std::vector<Pair> g_open[kNumSlots];     // begun, waiting for end
std::vector<Pair> g_done;                // recorded pairs not yet read
std::vector<Pair> g_free;
int64_t g_launches[kNumSlots] = {0};
const char* kNames[kNumSlots] = {"setup", "prep_points", "layer0_jets", "gemm_layer1", "gemm_layer2", "gemm_layer3",
                                 "gemm_layer4", "gemm_layer5", "gemm_layer6", "gemm_layer7", "final_blend",
                                 "residuals", "bwd_blend", "bwd_vertex", "", "",
                                 "bwd_wgrad_layer1", "bwd_wgrad_layer2", "bwd_wgrad_layer3", "bwd_wgrad_layer4",
                                 "bwd_wgrad_layer5", "bwd_wgrad_layer6", "bwd_wgrad_layer7", "",
                                 "bwd_dgrad_layer1", "bwd_dgrad_layer2", "bwd_dgrad_layer3", "bwd_dgrad_layer4",
                                 "bwd_dgrad_layer5", "bwd_dgrad_layer6", "bwd_dgrad_layer7", ""};
}  // namespace

void prof_begin(int slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_enabled) return;
    Pair p;
    if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
    else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    p.slot = slot;
    cudaEventRecord(p.a, st);
    g_open[slot].push_back(p);
}

void prof_end(int slot, cudaStream_t st, int n_launches) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_launches[slot] += n_launches;
    if (!g_enabled || g_open[slot].empty()) return;
    Pair p = g_open[slot].back();
    g_open[slot].pop_back();
    cudaEventRecord(p.b, st);
    g_done.push_back(p);
}

}  // namespace stpde

using namespace stpde;

extern "C" {

int stpde_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_enabled = on != 0;
    return 0;
}

int stpde_profile_read(double* ms_by_slot, int64_t* launches_by_slot, int n_slots) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(g_mu);
    for (int s = 0; s < n_slots && s < kNumSlots; ++s) {
        if (ms_by_slot) ms_by_slot[s] = 0.0;
        if (launches_by_slot) { launches_by_slot[s] = g_launches[s]; }
        g_launches[s] = 0;
    }
    for (const Pair& p : g_done) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess && ms_by_slot && p.slot < n_slots) ms_by_slot[p.slot] += ms;
        g_free.push_back(p);
    }
    g_done.clear();
    return kNumSlots;
}

const char* stpde_profile_slot_name(int slot) { return (slot >= 0 && slot < kNumSlots) ? kNames[slot] : ""; }

}  // extern "C"
