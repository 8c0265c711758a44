// api.cu -- the extern "C" boundary of libmpinn_b200.so (see include/mpinn_b200.h) plus the
// small kernels around the jet engines: weight packing, sum-of-squares reduction, fused Adam.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "common.cuh"
#include "launch.h"

namespace mpinn {

static thread_local std::string g_err;

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
static int fail_cuda(const char* what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 2;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// tensor-core operand block of one square hidden layer, in floats (jet_tc.cu)
int64_t tc_block_floats(int width);
cudaError_t launch_pack_tc(const DevProg& prog, const float* params, float* packed,
                           cudaStream_t stream);
cudaError_t read_tc_debug(long long* out, int n);
cudaError_t read_tc_debug_fwd(long long* out, int n);
// residual.cu
cudaError_t launch_residual(const mpinn_rprog_t& prog, const float* x, const float* data, int64_t n,
                            const float* const* jets,
                            float* const* djets, const float* weights, float* sums, float* values,
                            int values_point_major, int64_t jet_stride, const float* point_weights, float* wsums,
                            int sm_count, cudaStream_t stream);
cudaError_t launch_adam_tick(float* state, float beta1, float beta2, cudaStream_t stream);
cudaError_t launch_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* state,
                            float beta1, float beta2, float eps, float grad_scale, cudaStream_t stream);

// Translate the host program into the device description: padded dims, packed-buffer offsets,
// stash planes.  Returns an error string or nullptr.
static const char* build_prog(const mpinn_op_t* ops, int n_ops, int d, int order, DevProg* out,
                              const float* lap_coef = nullptr) {
  if (!ops || n_ops <= 0) return "empty program";
  if (n_ops > kMaxOps) return "program has more than MPINN_MAX_OPS ops";
  if (d < 1 || d > 3) return "input dimension must be 1, 2 or 3";
  if (order < 0 || order > 4) return "jet order must be 0, 1, 2, 3 (MPINN_ORDER_LAPLACE) or 4 (MPINN_ORDER_HESSIAN)";
  if (order == 4 && d < 2) return "MPINN_ORDER_HESSIAN needs at least two input dimensions (use order 2)";
  DevProg& p = *out;
  memset(&p, 0, sizeof(p));
  p.n_ops = n_ops;
  p.d = d;
  for (int i = 0; i < d; ++i) p.lap_c[i] = lap_coef ? lap_coef[i] : 1.0f;
  const int J = jet_count(d, order);
  int cur_w = d, max_w = round_up(d, 8), planes = 0;
  int64_t pk = 0;
  for (int l = 0; l < n_ops; ++l) {
    const mpinn_op_t& s = ops[l];
    DevOp& o = p.ops[l];
    o.kind = s.kind;
    o.in_dim = s.in_dim;
    o.out_dim = s.out_dim;
    o.act = s.act;
    o.flags = s.flags;
    o.w_off = s.w_off;
    o.b_off = s.b_off;
    o.stash_plane = -1;
    o.tc_off = -1;
    if (s.in_dim != cur_w) return "op in_dim does not match the previous op's out_dim";
    if (s.in_dim < 1 || s.out_dim < 1) return "bad op dims";
    if (s.kind == MPINN_OP_FOURIER) {
      if (l != 0) return "FOURIER must be the first op";
      if (s.out_dim % 2) return "FOURIER out_dim must be even";
      o.stash_plane = planes;   // z = xB duplicated for the sin and cos halves (tcgen05 engine only)
      planes += J * s.out_dim;
    } else if (s.kind == MPINN_OP_LINEAR) {
      if (s.act < 0 || s.act > 2) return "bad activation";
      if ((s.flags & MPINN_FLAG_ADD_SKIP) && s.out_dim != s.in_dim)
        return "skip connection needs equal widths";
      o.outp = round_up(s.out_dim, 8);
      o.inp = round_up(s.in_dim, 8);
      o.pk_fwd = (int)pk;
      pk += (int64_t)s.in_dim * o.outp;
      o.pk_bwd = (int)pk;
      pk += (int64_t)s.out_dim * o.inp;
      if (s.in_dim == s.out_dim && tc_block_floats(s.in_dim) > 0) {
        pk = round_up64(pk, 256);  // 1 KiB alignment for UMMA operand tiles
        o.tc_off = (int)pk;
        pk += tc_block_floats(s.in_dim);
      }
      pk = round_up64(pk, 4);
      const bool final_plain = (l == n_ops - 1) && s.act == MPINN_ACT_NONE;
      if (!final_plain) {
        o.stash_plane = planes;
        planes += J * s.out_dim;
      }
      if (s.flags & (MPINN_FLAG_SAVE_SKIP | MPINN_FLAG_ADD_SKIP)) p.has_skip = 1;
    } else {
      return "unknown op kind";
    }
    cur_w = s.out_dim;
    if (round_up(cur_w, 8) > max_w) max_w = round_up(cur_w, 8);
  }
  p.max_w = max_w;
  p.n_out = cur_w;
  p.total_planes = planes;
  p.packed_floats = pk;
  return nullptr;
}

struct DeviceInfo {
  int sm_count = 0;
  size_t smem_optin = 0;
  bool ok = false;
};
static cudaError_t device_info(DeviceInfo* info) {
  static thread_local DeviceInfo cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && cache[dev].ok) {
    *info = cache[dev];
    return cudaSuccess;
  }
  int v = 0;
  e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  info->sm_count = v;
  e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  info->smem_optin = (size_t)v;
  info->ok = true;
  if (dev >= 0 && dev < 64) cache[dev] = *info;
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------ kernels
// forward operand Wt[k][f] = W[f][k] (zero padded to outp), reverse operand Wn[o][i] = W[o][i]
// (zero padded to inp); tensor-core operand blocks are produced by launch_pack_tc (jet_tc.cu).

__global__ void pack_params_kernel(const __grid_constant__ DevProg prog,
                                   const float* __restrict__ params, float* __restrict__ packed) {
  const DevOp& op = prog.ops[blockIdx.y];
  if (op.kind != MPINN_OP_LINEAR) return;
  const float* W = params + op.w_off;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  float* fw = packed + op.pk_fwd;
  for (int i = tid; i < op.in_dim * op.outp; i += nth) {
    const int k = i / op.outp, f = i % op.outp;
    fw[i] = f < op.out_dim ? W[f * op.in_dim + k] : 0.0f;
  }
  float* bw = packed + op.pk_bwd;
  for (int i = tid; i < op.out_dim * op.inp; i += nth) {
    const int o = i / op.inp, k = i % op.inp;
    bw[i] = k < op.in_dim ? W[o * op.in_dim + k] : 0.0f;
  }
}

__global__ void sumsq_kernel(const float* __restrict__ r, int64_t n, int64_t stride,
                             float* __restrict__ accum) {
  float s = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = r[i * stride];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ float ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) ws[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < (int)(blockDim.x >> 5) ? ws[lane] : 0.0f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) atomicAdd(accum, s);
  }
}

// torch.optim.Adam's single-tensor update, restated (see oracle/jet_ref.py:adam_step)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float one_minus_b1, float b2,
                            float one_minus_b2, float step_size, float bc2_sqrt, float eps,
                            float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = fmaf(one_minus_b1, gi - m[i], m[i]);
  const float vi = fmaf(one_minus_b2 * gi, gi, v[i] * b2);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

}  // namespace mpinn

// ------------------------------------------------------------------------------------ C ABI
using namespace mpinn;

// per-CTA scratch for the jets that travel along a skip connection (tcgen05 engine): 148 CTAs x 7 channels
// x 64 features x 128 points, L2 resident; sits at the end of the stash (forward) / start of the workspace (reverse)
static const size_t kSkipScratchFloats = (size_t)148 * 7 * 64 * 128;

extern "C" {

int mpinn_abi_version(void) { return MPINN_ABI_VERSION; }

const char* mpinn_last_error(void) { return g_err.c_str(); }

size_t mpinn_stash_bytes(const mpinn_op_t* ops, int n_ops, int d, int order, int64_t n_points) {
  DevProg prog;
  if (build_prog(ops, n_ops, d, order, &prog)) return 0;
  const int64_t n_pad = round_up64(n_points, 128);
  return ((size_t)prog.total_planes * (size_t)n_pad + (prog.has_skip ? kSkipScratchFloats : 0) +
          (size_t)wide_scratch_floats(prog, d, order, n_pad)) * sizeof(float);
}

// one private slab of n_params floats per resident CTA (at most 4 per SM on a 148-SM part)
static const int kMaxSlabs = 148 * 4;

size_t mpinn_backward_workspace_bytes(const mpinn_op_t* ops, int n_ops, int64_t n_params) {
  (void)ops;
  (void)n_ops;
  return ((size_t)kMaxSlabs * (size_t)n_params + kSkipScratchFloats) * sizeof(float);
}

size_t mpinn_packed_bytes(const mpinn_op_t* ops, int n_ops) {
  DevProg prog;
  if (build_prog(ops, n_ops, ops ? ops[0].in_dim : 0, 0, &prog)) return 0;
  return (size_t)(prog.packed_floats + 256) * sizeof(float);
}

int mpinn_pack_params(const mpinn_op_t* ops, int n_ops, const float* params, void* packed,
                      void* stream) {
  DevProg prog;
  const char* err = build_prog(ops, n_ops, ops ? ops[0].in_dim : 0, 0, &prog);
  if (err) return fail(err);
  if (!params || !packed) return fail("null pointer");
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail("packed buffer must be 16-byte aligned");
  dim3 grid(16, n_ops);
  pack_params_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prog, params, (float*)packed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda("pack_params", e);
  e = launch_pack_tc(prog, params, (float*)packed, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("pack_params(tc)", e);
  e = launch_pack_wide(prog, params, (float*)packed, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("pack_params(tc, 128-wide)", e);
  return 0;
}

static bool use_tc(const DevProg& prog, int d, int order, int engine, bool backward, int* rc) {
  *rc = 0;
  if (engine == MPINN_ENGINE_FFMA) return false;
  const bool ok = (backward ? tc_backward_supported(prog, d, order) : tc_forward_supported(prog, d, order)) ||
                  wide_supported(prog, d, order);
  if (engine == MPINN_ENGINE_TCGEN05 && !ok) {
    *rc = fail("tcgen05 engine requested but the program is not supported by it");
    return false;
  }
  return ok;
}

int mpinn_jet_forward(const mpinn_op_t* ops, int n_ops, const float* params, const void* packed,
                      const float* x, int64_t n, int d, int order, const float* lap_coef_host, float* out,
                      void* stash, int engine, void* stream) {
  DevProg prog;
  if (order == MPINN_ORDER_LAPLACE && !lap_coef_host) return fail("order 3 needs lap_coef_host");
  const char* err = build_prog(ops, n_ops, d, order, &prog, lap_coef_host);
  if (err) return fail(err);
  if (ops[0].in_dim != d) return fail("first op in_dim must equal d");
  if (n <= 0) return 0;
  if (!params || !packed || !x || !out) return fail("null pointer");
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail("packed buffer must be 16-byte aligned");
  DeviceInfo info;
  cudaError_t e = device_info(&info);
  if (e != cudaSuccess) return fail_cuda("device query", e);
  const int64_t n_pad = round_up64(n, 128);
  int rc = 0;
  const bool tc = use_tc(prog, d, order, engine, false, &rc);
  if (rc) return rc;
  // stash == NULL is a value-only pass.  The tcgen05 engine has a no-stash variant for order 0 (inference) of
  // programs without skip connections (their per-CTA skip scratch lives behind the stash); otherwise ENGINE_AUTO
  // falls to the FP32-FMA engine and an explicit tcgen05 request is refused.
  const bool wide = tc && wide_supported(prog, d, order);   // layer-wise engine: the planes ARE its working storage
  const bool tc_ok = tc && (stash != nullptr || (order == 0 && !prog.has_skip && !wide));
  if (tc && !tc_ok && engine == MPINN_ENGINE_TCGEN05)
    return fail("the tcgen05 engine runs without a stash buffer only at order 0, without skip connections and for "
                "64-wide layers (pass mpinn_stash_bytes() bytes)");
  if (tc_ok && wide) {
    e = launch_fwd_wide(prog, params, (const float*)packed, x, n, d, order, out, (float*)stash, n_pad, info.sm_count,
                        (cudaStream_t)stream);
  } else if (tc_ok) {
    e = launch_fwd_tc(prog, params, (const float*)packed, x, n, d, order, out, (float*)stash, n_pad,
                      info.sm_count, (cudaStream_t)stream);
  } else {
    e = launch_fwd_ffma(prog, params, (const float*)packed, x, n, d, order, out, (float*)stash,
                        n_pad, info.sm_count, info.smem_optin, (cudaStream_t)stream);
  }
  if (e == cudaErrorInvalidConfiguration)
    return fail("layer widths too large for the shared-memory budget of this engine");
  if (e != cudaSuccess) return fail_cuda("jet_forward", e);
  return 0;
}

// bsum[o] = sum_pt dout[0][pt][o] in double precision (per-thread double sums, shuffle + shared-memory block reduction,
// one double atomic per CTA and output): the last layer's bias gradient (DevProg::bsum)
constexpr int kBsumMaxOut = 8;
__global__ void value_seed_colsum(const float* __restrict__ dout, int64_t n, int no, double* __restrict__ bsum,
                                  int single_cta) {
  double acc[kBsumMaxOut];
#pragma unroll
  for (int o = 0; o < kBsumMaxOut; ++o) acc[o] = 0.0;
  for (int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pt < n; pt += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int o = 0; o < kBsumMaxOut; ++o)
      if (o < no) acc[o] += (double)dout[pt * no + o];
  __shared__ double red[kBsumMaxOut][8];
#pragma unroll
  for (int o = 0; o < kBsumMaxOut; ++o) {
    if (o < no) {
      double v = acc[o];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31) == 0) red[o][threadIdx.x >> 5] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < no) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    if (single_cta) bsum[threadIdx.x] = v;   // small launches: one CTA, no memset node in front
    else atomicAdd(bsum + threadIdx.x, v);
  }
}

int mpinn_jet_backward(const mpinn_op_t* ops, int n_ops, const float* params, const void* packed,
                       const float* x, int64_t n, int d, int order, const float* lap_coef_host,
                       const float* dout, const void* stash, float* dparams, int64_t n_params, float* dx,
                       void* workspace, size_t workspace_bytes, int engine, void* stream) {
  DevProg prog;
  if (order == MPINN_ORDER_LAPLACE && !lap_coef_host) return fail("order 3 needs lap_coef_host");
  const char* err = build_prog(ops, n_ops, d, order, &prog, lap_coef_host);
  if (err) return fail(err);
  if (n <= 0) return 0;
  if (!params || !packed || !x || !dout || !stash || !dparams || !workspace) return fail("null pointer");
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail("packed buffer must be 16-byte aligned");
  if (workspace_bytes < kSkipScratchFloats * sizeof(float)) return fail("backward workspace too small");
  float* skip_scratch = (float*)workspace;                       // first part of the workspace
  workspace = (void*)((float*)workspace + kSkipScratchFloats);   // slabs follow
  workspace_bytes -= kSkipScratchFloats * sizeof(float);
  int slabs = (int)(workspace_bytes / ((size_t)n_params * sizeof(float)));
  if (slabs < 1) return fail("backward workspace too small");
  if (slabs > kMaxSlabs) slabs = kMaxSlabs;
  DeviceInfo info;
  cudaError_t e = device_info(&info);
  if (e != cudaSuccess) return fail_cuda("device query", e);
  const int64_t n_pad = round_up64(n, 128);
  int rc = 0;
  const bool tc = use_tc(prog, d, order, engine, true, &rc);
  if (rc) return rc;
  // the last layer's bias gradient in double precision: eight doubles at the (8-byte aligned) end of the workspace,
  // behind the slabs the engines may use
  const DevOp& last_op = prog.ops[n_ops - 1];
  if (last_op.kind == MPINN_OP_LINEAR && last_op.act == MPINN_ACT_NONE && !last_op.flags && last_op.b_off >= 0 &&
      prog.n_out <= kBsumMaxOut) {
    // slabs the sums (plus alignment slack) may overlap: more than one for very small networks
    const size_t slab_bytes = (size_t)n_params * sizeof(float);
    const int reserve = (int)((kBsumMaxOut * sizeof(double) + 8 + slab_bytes - 1) / slab_bytes) + 1;
    if (slabs < reserve + 2) return fail("backward workspace too small");
    uintptr_t end = (reinterpret_cast<uintptr_t>(workspace) + workspace_bytes) & ~uintptr_t(7);
    double* bsum = reinterpret_cast<double*>(end) - kBsumMaxOut;
    slabs -= reserve;   // the engines never touch the slabs the sums live in
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 2 * info.sm_count) blocks = 2 * info.sm_count;
    if (blocks > 1) {
      e = cudaMemsetAsync(bsum, 0, kBsumMaxOut * sizeof(double), (cudaStream_t)stream);
      if (e != cudaSuccess) return fail_cuda("jet_backward (seed sum)", e);
    }
    value_seed_colsum<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, n, prog.n_out, bsum, blocks == 1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda("jet_backward (seed sum)", e);
    prog.bsum = bsum;
    prog.bsum_off = last_op.b_off;
  }
  if (tc && dx && ops[0].kind == MPINN_OP_FOURIER)
    return fail("gradient w.r.t. the points (dx) through a Fourier-feature first op runs on the FP32-FMA engine: "
                "request MPINN_ENGINE_FFMA for the forward AND the reverse pass (the stash layouts differ)");
  const bool wide = tc && wide_supported(prog, d, order);
  if (wide && dx)
    return fail("gradient w.r.t. the points (dx) of a 128-wide program runs on the FP32-FMA engine: request "
                "MPINN_ENGINE_FFMA for the forward AND the reverse pass (the stash layouts differ)");
  if (wide) {
    // the reverse pass keeps its Zbar planes behind the stash (mpinn_stash_bytes() reserves them)
    e = launch_bwd_wide(prog, params, (const float*)packed, x, n, d, order, dout, (float*)const_cast<void*>(stash), n_pad,
                        (float*)workspace, slabs, n_params, dparams, info.sm_count, (cudaStream_t)stream);
  } else if (tc) {
    e = launch_bwd_tc(prog, params, (const float*)packed, x, n, d, order, dout, (const float*)stash,
                      n_pad, (float*)workspace, slabs, n_params, dparams, dx, skip_scratch, info.sm_count,
                      (cudaStream_t)stream);
  } else {
    e = launch_bwd_ffma(prog, params, (const float*)packed, x, n, d, order, dout,
                        (const float*)stash, n_pad, (float*)workspace, slabs, n_params, dparams, dx,
                        info.sm_count, info.smem_optin, (cudaStream_t)stream);
  }
  if (e == cudaErrorInvalidConfiguration)
    return fail("layer widths too large for the shared-memory budget of this engine");
  if (e != cudaSuccess) return fail_cuda("jet_backward", e);
  return 0;
}

int mpinn_engine_supports(const mpinn_op_t* ops, int n_ops, int d, int order, int engine) {
  DevProg prog;
  if (build_prog(ops, n_ops, d, order, &prog)) return 0;
  if (engine == MPINN_ENGINE_TCGEN05) return (tc_forward_supported(prog, d, order) || wide_supported(prog, d, order)) ? 1 : 0;
  return 1;   // the FP32-FMA engine runs every valid program (shared-memory budget is checked at launch)
}

int mpinn_residual_program(const mpinn_rprog_t* prog, const float* x, const float* data, int64_t n,
                           const float* const* jets, float* const* djets, const float* weights, float* sums, float* values,
                           int values_point_major, int64_t jet_stride, const float* point_weights, float* wsums,
                           void* stream) {
  if (!prog || !x) return fail("null pointer");
  if (prog->n_blocks > 0 && !jets) return fail("residual program: null jet table");
  if (prog->n_blocks == 0 && djets) return fail("residual program: seeds requested without model calls");
  if (jet_stride != 0 && jet_stride < n) return fail("residual program: jet_stride smaller than n");
  if (n <= 0) return 0;
  if (prog->n_instr < 1 || prog->n_instr > MPINN_RP_MAX_INSTR) return fail("residual program: bad instruction count");
  if (prog->n_out < 1 || prog->n_out > MPINN_RP_MAX_OUT) return fail("residual program: bad output count");
  if (prog->n_blocks < 0 || prog->n_blocks > MPINN_RP_MAX_BLOCKS) return fail("residual program: bad block count");
  if (prog->n_const < 0 || prog->n_const > MPINN_RP_MAX_CONST) return fail("residual program: bad constant count");
  if (prog->d < 1 || prog->d > 3) return fail("residual program: bad input dimension");
  if (prog->n_data < 0 || prog->n_data > MPINN_RP_MAX_DATA) return fail("residual program: bad data column count");
  if (prog->n_data > 0 && !data) return fail("residual program: null data block");
  for (int b = 0; b < prog->n_blocks; ++b) {
    if (!jets[b]) return fail("residual program: null jet block");
    if (prog->block_o[b] < 1 || prog->block_J[b] < 1 || prog->block_J[b] * prog->block_o[b] > MPINN_RP_MAX_JO)
      return fail("residual program: bad block shape");
  }
  // every operand must name an earlier register (SSA) and every leaf an existing column
  for (int i = 0; i < prog->n_instr; ++i) {
    const mpinn_rinstr_t& in = prog->instr[i];
    bool ok = true;
    switch (in.op) {
      case MPINN_RP_CONST: ok = in.a >= 0 && in.a < prog->n_const; break;
      case MPINN_RP_LOADX: ok = in.a >= 0 && in.a < prog->d; break;
      case MPINN_RP_LOADD: ok = in.a >= 0 && in.a < prog->n_data; break;
      case MPINN_RP_LOADJ:
        ok = in.a >= 0 && in.a < prog->n_blocks && in.b >= 0 && in.b < prog->block_J[in.a] * prog->block_o[in.a];
        break;
      case MPINN_RP_ADD: case MPINN_RP_SUB: case MPINN_RP_MUL: case MPINN_RP_DIV: case MPINN_RP_POWF:
      case MPINN_RP_MAX: case MPINN_RP_MIN: case MPINN_RP_GT: case MPINN_RP_GE:
        ok = in.a >= 0 && in.a < i && in.b >= 0 && in.b < i; break;
      case MPINN_RP_SEL:
        ok = in.a >= 0 && in.a < i && in.b >= 0 && in.b < i && in.c >= 0 && in.c < i; break;
      case MPINN_RP_NEG: case MPINN_RP_SIN: case MPINN_RP_COS: case MPINN_RP_EXP: case MPINN_RP_TANH:
      case MPINN_RP_RELU: case MPINN_RP_ABS: case MPINN_RP_SQRT: case MPINN_RP_LOG: case MPINN_RP_POWI:
        ok = in.a >= 0 && in.a < i; break;
      default: ok = false;
    }
    if (!ok) return fail("residual program: malformed instruction " + std::to_string(i));
  }
  for (int k = 0; k < prog->n_out; ++k)
    if (prog->out_reg[k] < 0 || prog->out_reg[k] >= prog->n_instr) return fail("residual program: bad output register");
  DeviceInfo info;
  cudaError_t e = device_info(&info);
  if (e != cudaSuccess) return fail_cuda("device query", e);
  e = launch_residual(*prog, x, data, n, jets, djets, weights, sums, values, values_point_major, jet_stride,
                      point_weights, wsums, info.sm_count, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("residual_program", e);
  return 0;
}

int mpinn_adam_tick(float* state, float beta1, float beta2, void* stream) {
  if (!state) return fail("null pointer");
  cudaError_t e = launch_adam_tick(state, beta1, beta2, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("adam_tick", e);
  return 0;
}

int mpinn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* state, float beta1, float beta2, float eps, float grad_scale,
                        void* stream) {
  if (n <= 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state) return fail("null pointer");
  cudaError_t e = launch_adam_dev(params, grads, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, grad_scale,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("adam_step_dev", e);
  return 0;
}

int mpinn_selftest_umma(int mode, const float* a, const float* b, float* c, void* stream) {
  if (!a || !b || !c) return fail("null pointer");
  if (mode != 0 && mode != 2 && mode != 3 && mode != 4) return fail("mode must be 0, 2, 3 or 4");
  cudaError_t e = launch_umma_selftest(mode, a, b, c, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("selftest_umma", e);
  return 0;
}

int mpinn_debug_counters(long long* out_host, int n) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = read_tc_debug(out_host, n);
  if (e == cudaSuccess && n > 32) e = read_tc_debug_fwd(out_host + 32, n - 32);   // forward kernel: counters 32..63
  if (e != cudaSuccess) return fail_cuda("debug_counters", e);
  return 0;
}

int mpinn_sumsq(const float* r, int64_t n, int64_t stride, float* accum, void* stream) {
  if (n <= 0) return 0;
  if (!r || !accum) return fail("null pointer");
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(r, n, stride, accum);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda("sumsq", e);
  return 0;
}

int mpinn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    int64_t step, float lr, float beta1, float beta2, float eps, float grad_scale,
                    void* stream) {
  if (n <= 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail("null pointer");
  if (step < 1) return fail("step must be >= 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, 1.0f - beta1, beta2, 1.0f - beta2, step_size,
      bc2_sqrt, eps, grad_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda("adam_step", e);
  return 0;
}

}  // extern "C"
