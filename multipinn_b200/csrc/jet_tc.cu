// jet_tc.cu -- tcgen05 (5th-gen tensor core) engine for the square hidden layers.
// Placeholder: reports "unsupported" until the tensor-core path lands; the FFMA engine runs.
#include "common.cuh"
#include "launch.h"

namespace mpinn {

int64_t tc_block_floats(int) { return 0; }

cudaError_t launch_pack_tc(const DevProg&, const float*, float*, cudaStream_t) { return cudaSuccess; }

bool tc_supported(const DevProg&, int, int) { return false; }

cudaError_t launch_fwd_tc(const DevProg&, const float*, const float*, const float*, int64_t, int, int,
                          float*, float*, int64_t, int, cudaStream_t) {
  return cudaErrorNotSupported;
}

cudaError_t launch_bwd_tc(const DevProg&, const float*, const float*, const float*, int64_t, int, int,
                          const float*, const float*, int64_t, float*, int, int64_t, float*, float*,
                          int, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace mpinn
