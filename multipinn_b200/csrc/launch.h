// launch.h -- internal declarations shared by the translation units of libmpinn_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace mpinn {

// FP32-FMA engine (jet_ffma.cu)
cudaError_t launch_fwd_ffma(const DevProg& prog, const float* params, const float* packed,
                            const float* x, int64_t n, int d, int order, float* out, float* stash,
                            int64_t n_pad, int sm_count, size_t smem_limit, cudaStream_t stream);
cudaError_t launch_bwd_ffma(const DevProg& prog, const float* params, const float* packed,
                            const float* x, int64_t n, int d, int order, const float* dout,
                            const float* stash, int64_t n_pad, float* part, int max_slabs,
                            int64_t n_params, float* dparams, float* dx, int sm_count,
                            size_t smem_limit, cudaStream_t stream);

// tcgen05 engine (jet_tc.cu)
bool tc_forward_supported(const DevProg& prog, int d, int order);
bool tc_backward_supported(const DevProg& prog, int d, int order);
cudaError_t launch_umma_selftest(int mode, const float* A, const float* B, float* C, cudaStream_t stream);
cudaError_t launch_fwd_tc(const DevProg& prog, const float* params, const float* packed,
                          const float* x, int64_t n, int d, int order, float* out, float* stash,
                          int64_t n_pad, int sm_count, cudaStream_t stream);
cudaError_t launch_bwd_tc(const DevProg& prog, const float* params, const float* packed,
                          const float* x, int64_t n, int d, int order, const float* dout,
                          const float* stash, int64_t n_pad, float* part, int max_slabs,
                          int64_t n_params, float* dparams, float* dx, float* skip_scratch, int sm_count,
                          cudaStream_t stream);

// tcgen05 engine for 128-wide plain FNN programs, layer-wise (jet_tc_wide.cu)
bool wide_supported(const DevProg& prog, int d, int order);
int64_t wide_block_floats();
int64_t wide_scratch_floats(const DevProg& prog, int d, int order, int64_t n_pad);
cudaError_t launch_pack_wide(const DevProg& prog, const float* params, float* packed, cudaStream_t stream);
cudaError_t launch_fwd_wide(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                            int d, int order, float* out, float* stash, int64_t n_pad, int sm_count,
                            cudaStream_t stream);
cudaError_t launch_bwd_wide(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                            int d, int order, const float* dout, float* stash, int64_t n_pad, float* part,
                            int max_slabs, int64_t n_params, float* dparams, int sm_count, cudaStream_t stream);

}  // namespace mpinn
