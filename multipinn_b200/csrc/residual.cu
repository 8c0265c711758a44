// residual.cu -- device evaluation of a traced residual program (see include/mpinn_b200.h, MPINN_RP_*):
// per point, the forward sweep computes every register, the residual components feed sum-of-squares
// accumulators (warp shuffle + one atomic per CTA and component), and a reverse sweep over the same tape
// produces dL/d(jet) -- the seeds of the fused reverse pass -- for  L = sum_k w_k sum_n r_k[n]^2.
//
// Replaces, for one condition: the user's residual algebra as ~30-100 eager element-wise kernels, torch.sum(r**2)
// (reference PINN/PINN.py:109-125) and the autograd graph between the residual list and the network outputs.
//
// HBM-bound by construction: 4 B x (J*o + d) read and 4 B x J*o written per point.  The register file of the
// little virtual machine lives in shared memory as [register][thread] (conflict-free, no local-memory traffic);
// the program itself sits in the kernel parameter (constant) bank and is read warp-uniformly.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpinn_b200.h"

namespace mpinn {

constexpr int kResidThreads = 128;

struct ResidIO {
  const float* jets[MPINN_RP_MAX_BLOCKS];
  float* djets[MPINN_RP_MAX_BLOCKS];
  const float* point_weights;   // [n_out][n] per-point loss weights lambda_k[n] (PointwiseLosses), or NULL
  float* wsums;                 // [n_out] += sum_n lambda_k[n] r_k[n]^2, or NULL
};

__device__ __forceinline__ float powi(float v, int k) {
  const bool neg = k < 0;
  unsigned e = neg ? (unsigned)(-k) : (unsigned)k;
  float r = 1.0f, b = v;
  while (e) {
    if (e & 1u) r *= b;
    b *= b;
    e >>= 1;
  }
  return neg ? 1.0f / r : r;
}

template <bool SEEDS>
__global__ void __launch_bounds__(kResidThreads)
residual_kernel(const __grid_constant__ mpinn_rprog_t prog, const ResidIO io, const float* __restrict__ x,
                const float* __restrict__ data, int64_t n, const float* __restrict__ weights, float* __restrict__ sums,
                float* __restrict__ values, int values_point_major, int64_t jstride) {
  extern __shared__ float rp_smem[];
  const int t = threadIdx.x;
  const int NI = prog.n_instr;
  float* v = rp_smem + t;                                  // register i at v[i * kResidThreads]
  float* adj = rp_smem + NI * kResidThreads + t;
  float acc[MPINN_RP_MAX_OUT];
#pragma unroll
  for (int k = 0; k < MPINN_RP_MAX_OUT; ++k) acc[k] = 0.0f;
  float wk[MPINN_RP_MAX_OUT], accw[MPINN_RP_MAX_OUT];
#pragma unroll
  for (int k = 0; k < MPINN_RP_MAX_OUT; ++k) {
    wk[k] = (SEEDS && k < prog.n_out) ? (weights ? __ldg(weights + k) : 1.0f) : 0.0f;
    accw[k] = 0.0f;
  }
  const float* __restrict__ pwts = io.point_weights;

  for (int64_t pt = (int64_t)blockIdx.x * kResidThreads + t; pt < n; pt += (int64_t)gridDim.x * kResidThreads) {
    // ---------------------------------------------------------------- forward sweep
    for (int i = 0; i < NI; ++i) {
      const mpinn_rinstr_t in = prog.instr[i];
      float r;
      switch (in.op) {
        case MPINN_RP_CONST: r = prog.consts[in.a]; break;
        case MPINN_RP_LOADX: r = __ldg(x + pt * prog.d + in.a); break;
        case MPINN_RP_LOADD: r = __ldg(data + pt * prog.n_data + in.a); break;
        case MPINN_RP_LOADJ: {
          const int o = prog.block_o[in.a];
          const int j = in.b / o, k = in.b - j * o;
          r = __ldg(io.jets[in.a] + ((int64_t)j * jstride + pt) * o + k);
        } break;
        case MPINN_RP_ADD: r = v[in.a * kResidThreads] + v[in.b * kResidThreads]; break;
        case MPINN_RP_SUB: r = v[in.a * kResidThreads] - v[in.b * kResidThreads]; break;
        case MPINN_RP_MUL: r = v[in.a * kResidThreads] * v[in.b * kResidThreads]; break;
        case MPINN_RP_DIV: r = v[in.a * kResidThreads] / v[in.b * kResidThreads]; break;
        case MPINN_RP_MAX: r = fmaxf(v[in.a * kResidThreads], v[in.b * kResidThreads]); break;
        case MPINN_RP_MIN: r = fminf(v[in.a * kResidThreads], v[in.b * kResidThreads]); break;
        case MPINN_RP_GT: r = v[in.a * kResidThreads] > v[in.b * kResidThreads] ? 1.0f : 0.0f; break;
        case MPINN_RP_GE: r = v[in.a * kResidThreads] >= v[in.b * kResidThreads] ? 1.0f : 0.0f; break;
        case MPINN_RP_SEL:
          r = v[in.a * kResidThreads] != 0.0f ? v[in.b * kResidThreads] : v[in.c * kResidThreads];
          break;
        case MPINN_RP_NEG: r = -v[in.a * kResidThreads]; break;
        case MPINN_RP_SIN: r = sinf(v[in.a * kResidThreads]); break;
        case MPINN_RP_COS: r = cosf(v[in.a * kResidThreads]); break;
        case MPINN_RP_EXP: r = expf(v[in.a * kResidThreads]); break;
        case MPINN_RP_TANH: r = tanhf(v[in.a * kResidThreads]); break;
        case MPINN_RP_RELU: r = fmaxf(v[in.a * kResidThreads], 0.0f); break;
        case MPINN_RP_ABS: r = fabsf(v[in.a * kResidThreads]); break;
        case MPINN_RP_SQRT: r = sqrtf(v[in.a * kResidThreads]); break;
        case MPINN_RP_LOG: r = logf(v[in.a * kResidThreads]); break;
        case MPINN_RP_POWI: r = powi(v[in.a * kResidThreads], in.b); break;
        case MPINN_RP_POWF: r = powf(v[in.a * kResidThreads], v[in.b * kResidThreads]); break;
        default: r = 0.0f; break;
      }
      v[i * kResidThreads] = r;
      if (SEEDS) adj[i * kResidThreads] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < MPINN_RP_MAX_OUT; ++k) {
      if (k < prog.n_out) {
        const int reg = prog.out_reg[k];
        const float r = v[reg * kResidThreads];
        const float pw = pwts ? __ldg(pwts + (int64_t)k * n + pt) : 1.0f;
        acc[k] = fmaf(r, r, acc[k]);
        accw[k] = fmaf(pw * r, r, accw[k]);
        if (values) values[values_point_major ? pt * prog.n_out + k : (int64_t)k * n + pt] = r;
        if (SEEDS) adj[reg * kResidThreads] += 2.0f * wk[k] * pw * r;
      }
    }
    if (!SEEDS) continue;
    // ---------------------------------------------------------------- reverse sweep over the tape
    for (int i = NI - 1; i >= 0; --i) {
      const mpinn_rinstr_t in = prog.instr[i];
      if (in.op < MPINN_RP_ADD) continue;          // leaves: read out below
      const float g = adj[i * kResidThreads];
      float* aa = adj + in.a * kResidThreads;
      const float va = v[in.a * kResidThreads];
      switch (in.op) {
        case MPINN_RP_ADD: *aa += g; adj[in.b * kResidThreads] += g; break;
        case MPINN_RP_SUB: *aa += g; adj[in.b * kResidThreads] -= g; break;
        case MPINN_RP_MUL: {
          const float vb = v[in.b * kResidThreads];
          *aa += g * vb;                             // (a == b is fine: two separate read-modify-writes)
          adj[in.b * kResidThreads] += g * va;
        } break;
        case MPINN_RP_DIV: {
          const float inv = 1.0f / v[in.b * kResidThreads];
          *aa += g * inv;
          adj[in.b * kResidThreads] -= g * v[i * kResidThreads] * inv;
        } break;
        case MPINN_RP_MAX: case MPINN_RP_MIN: {     // the gradient follows the selected operand (a on ties)
          const float vb = v[in.b * kResidThreads];
          const bool pick_a = in.op == MPINN_RP_MAX ? va >= vb : va <= vb;
          if (pick_a) *aa += g; else adj[in.b * kResidThreads] += g;
        } break;
        case MPINN_RP_GT: case MPINN_RP_GE: break;  // masks are piecewise constant
        case MPINN_RP_SEL:                          // (aa is the mask's adjoint here: untouched)
          if (va != 0.0f) adj[in.b * kResidThreads] += g; else adj[in.c * kResidThreads] += g;
          break;
        case MPINN_RP_NEG: *aa -= g; break;
        case MPINN_RP_SIN: *aa += g * cosf(va); break;
        case MPINN_RP_COS: *aa -= g * sinf(va); break;
        case MPINN_RP_EXP: *aa += g * v[i * kResidThreads]; break;
        case MPINN_RP_TANH: { const float y = v[i * kResidThreads]; *aa += g * (1.0f - y * y); } break;
        case MPINN_RP_RELU: *aa += va > 0.0f ? g : 0.0f; break;
        case MPINN_RP_ABS: *aa += va > 0.0f ? g : (va < 0.0f ? -g : 0.0f); break;
        case MPINN_RP_SQRT: *aa += g * 0.5f / v[i * kResidThreads]; break;
        case MPINN_RP_LOG: *aa += g / va; break;
        case MPINN_RP_POWI: *aa += g * (float)in.b * powi(va, in.b - 1); break;
        case MPINN_RP_POWF: { const float e = v[in.b * kResidThreads]; *aa += g * e * powf(va, e - 1.0f); } break;
        default: break;
      }
    }
    for (int b = 0; b < prog.n_blocks; ++b) {
      float* dj = io.djets[b];
      if (dj == nullptr) continue;
      const int o = prog.block_o[b], JO = prog.block_J[b] * o;
      for (int jo = 0; jo < JO; ++jo) {
        const int reg = prog.jet_reg[b][jo];
        const int j = jo / o, k = jo - j * o;
        dj[((int64_t)j * jstride + pt) * o + k] = reg >= 0 ? adj[reg * kResidThreads] : 0.0f;
      }
    }
  }

  if (io.wsums != nullptr) {      // weighted sums first (same reduction, separate scratch)
    __shared__ float redw[MPINN_RP_MAX_OUT][kResidThreads / 32];
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int k = 0; k < MPINN_RP_MAX_OUT; ++k) {
      if (k < prog.n_out) {
        float s = accw[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) redw[k][warp] = s;
      }
    }
    __syncthreads();
    if (t < prog.n_out) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kResidThreads / 32; ++w) s += redw[t][w];
      atomicAdd(io.wsums + t, s);
    }
  }
  if (sums != nullptr) {
    __shared__ float red[MPINN_RP_MAX_OUT][kResidThreads / 32];
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int k = 0; k < MPINN_RP_MAX_OUT; ++k) {
      if (k < prog.n_out) {
        float s = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) red[k][warp] = s;
      }
    }
    __syncthreads();
    if (t < prog.n_out) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kResidThreads / 32; ++w) s += red[t][w];
      atomicAdd(sums + t, s);
    }
  }
}

cudaError_t launch_residual(const mpinn_rprog_t& prog, const float* x, const float* data, int64_t n,
                            const float* const* jets,
                            float* const* djets, const float* weights, float* sums, float* values,
                            int values_point_major, int64_t jet_stride, const float* point_weights, float* wsums,
                            int sm_count, cudaStream_t stream) {
  const int64_t jstride = jet_stride > 0 ? jet_stride : n;
  ResidIO io;
  io.point_weights = point_weights;
  io.wsums = wsums;
  for (int b = 0; b < MPINN_RP_MAX_BLOCKS; ++b) {
    io.jets[b] = (b < prog.n_blocks && jets) ? jets[b] : nullptr;
    io.djets[b] = (b < prog.n_blocks && djets) ? djets[b] : nullptr;
  }
  const bool seeds = djets != nullptr;
  const size_t smem = (size_t)(seeds ? 2 : 1) * prog.n_instr * kResidThreads * sizeof(float);
  int64_t blocks = (n + kResidThreads - 1) / kResidThreads;
  const int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  cudaError_t e;
  if (seeds) {
    e = cudaFuncSetAttribute(residual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    residual_kernel<true><<<(unsigned)blocks, kResidThreads, smem, stream>>>(prog, io, x, data, n, weights, sums,
                                                                             values, values_point_major, jstride);
  } else {
    e = cudaFuncSetAttribute(residual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    residual_kernel<false><<<(unsigned)blocks, kResidThreads, smem, stream>>>(prog, io, x, data, n, weights, sums,
                                                                              values, values_point_major, jstride);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ Adam, device state
// state = {lr, step, step_size, sqrt(1 - beta2^step)}: one thread advances the step count and refreshes the bias
// corrections in double precision (torch.optim.Adam computes them on the host in double).
__global__ void adam_tick_kernel(float* __restrict__ state, float beta1, float beta2) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double step = (double)state[1] + 1.0;
  state[1] = (float)step;
  const double bc1 = 1.0 - pow((double)beta1, step);
  const double bc2 = 1.0 - pow((double)beta2, step);
  state[2] = (float)((double)state[0] / bc1);
  state[3] = (float)sqrt(bc2);
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ state,
                                float one_minus_b1, float b2, float one_minus_b2, float eps, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float step_size = __ldg(state + 2), bc2_sqrt = __ldg(state + 3);
  const float gi = g[i] * grad_scale;
  const float mi = fmaf(one_minus_b1, gi - m[i], m[i]);
  const float vi = fmaf(one_minus_b2 * gi, gi, v[i] * b2);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

cudaError_t launch_adam_tick(float* state, float beta1, float beta2, cudaStream_t stream) {
  adam_tick_kernel<<<1, 32, 0, stream>>>(state, beta1, beta2);
  return cudaGetLastError();
}

cudaError_t launch_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* state,
                            float beta1, float beta2, float eps, float grad_scale, cudaStream_t stream) {
  adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, state, 1.0f - beta1, beta2,
                                                                   1.0f - beta2, eps, grad_scale);
  return cudaGetLastError();
}

}  // namespace mpinn
