// Placeholder until the tcgen05 kernels land: the tensor-core precisions report STPDE_EUNSUPPORTED.
#include "tc_path.h"

namespace stpde {

size_t tc_fixed_bytes(const stpde_desc_t*, int, const int*, const int*) { return 0; }
size_t tc_per_point_bytes(const stpde_desc_t*, int, int, int, int) { return 0; }
int tc_prepare(TcContext&, const stpde_desc_t*, int, const int*, const int*, const int*, const int*,
               const float* const*, char*, char*, size_t, const JetSpec&, int, int, int*, cudaStream_t) {
    return STPDE_EUNSUPPORTED;
}
int tc_run_chunk(TcContext&, const JetSpec&, int, int, float, const ChunkBuffers&, const float*, const float*, int,
                 const int*, const float* const*, char*, const size_t*, float*, cudaStream_t) {
    return STPDE_EUNSUPPORTED;
}
const char* tc_last_error() { return "tensor-core path not available in this build"; }

}  // namespace stpde
