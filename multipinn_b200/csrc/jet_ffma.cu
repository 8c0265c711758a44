// jet_ffma.cu -- whole-network fused jet forward / reverse kernels on the FP32 FMA pipe.
//
// This is the general engine (any layer widths up to the shared-memory budget, any of the
// FNN / ResNet / Fourier programs).  One persistent CTA walks tiles of TP collocation points;
// for each tile the complete layer program runs with all activations resident in shared
// memory (feature-major planes, points contiguous: conflict-free for lane==point access).
// Pre-activation jets are stashed to HBM once in the forward pass (coalesced planes) and read
// back once per use in the reverse pass; parameter gradients accumulate in a per-CTA private
// slab that a final kernel folds into the caller's gradient vector.
//
// Reference behaviour implemented: FNN.forward (neural_network/feedforward_neural_network.py:50-52),
// LightResidualBlock.forward (light_residual_block.py:33-38), FourierEncoding.forward
// (fourier_features.py:37-40), grad() sweeps (condition/diff.py:47) and loss.backward()
// (trainer/trainer.py:141) -- restated as jets, see oracle/jet_ref.py.
#include "common.cuh"
#include "launch.h"

namespace mpinn {

// Threads per CTA and output features per thread-chunk.  256 / 8 is the shipped configuration; the diagnostic build
// -DMPINN_FFMA_THREADS=512 -DMPINN_FFMA_CH=4 doubles the resident warps (same shared memory) for 128-wide layers.
#ifndef MPINN_FFMA_THREADS
#define MPINN_FFMA_THREADS 256
#endif
#ifndef MPINN_FFMA_CH
#define MPINN_FFMA_CH 8
#endif
constexpr int kThreads = MPINN_FFMA_THREADS;
constexpr int kCh = MPINN_FFMA_CH;      // 4 or 8
static_assert(kCh == 4 || kCh == 8, "chunk width");

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n_floats) {
  // n_floats is a multiple of 4 and both pointers are 16-byte aligned (packed buffer layout)
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n_floats / 4; i += kThreads) d4[i] = __ldg(s4 + i);
}

// acc[f][j] += sum_k Ws[k*ldw + kCh*chunk + f] * hin[(j*W + k)*TPP + p]
template <int J, int TPP>
__device__ __forceinline__ void gemm_chunk(const float* __restrict__ hin, const float* __restrict__ Ws,
                                           int kdim, int ldw, int chunk, int W, int p,
                                           float (&acc)[kCh][J]) {
#pragma unroll
  for (int f = 0; f < kCh; ++f)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[f][j] = 0.0f;
  const float* wrow = Ws + kCh * chunk;
#pragma unroll 2
  for (int k = 0; k < kdim; ++k) {
    float w[kCh];
    const float4 w0 = *reinterpret_cast<const float4*>(wrow + k * ldw);
    w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
    if (kCh == 8) {
      const float4 w1 = *reinterpret_cast<const float4*>(wrow + k * ldw + 4);
      w[kCh - 4] = w1.x; w[kCh - 3] = w1.y; w[kCh - 2] = w1.z; w[kCh - 1] = w1.w;
    }
    float h[J];
#pragma unroll
    for (int j = 0; j < J; ++j) h[j] = hin[(j * W + k) * TPP + p];
#pragma unroll
    for (int f = 0; f < kCh; ++f)
#pragma unroll
      for (int j = 0; j < J; ++j) acc[f][j] = fmaf(w[f], h[j], acc[f][j]);
  }
}

template <int D, int ORDER, int TPP>
__device__ __forceinline__ void write_input_jets(float* buf, int W, const float* xv, int p, int g, int G) {
  constexpr int J = jet_count(D, ORDER);
  for (int k = g; k < D; k += G) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float v = 0.0f;
      if (j == 0) v = xv[k];
      else if (j <= D) v = (j - 1 == k) ? 1.0f : 0.0f;
      buf[(j * W + k) * TPP + p] = v;
    }
  }
}

// Fourier encoding jets of the raw input: out planes [0,m) = sin, [m,2m) = cos
template <int D, int ORDER, int TPP>
__device__ __forceinline__ void fourier_jets(float* buf, int W, const float* __restrict__ B, int m,
                                             const float* xv, int p, int g, int G, const float* lc) {
  constexpr int J = jet_count(D, ORDER);
  for (int f = g; f < m; f += G) {
    float z[J], sj[J], cj[J];
    z[0] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float b = __ldg(B + i * m + f);
      z[0] = fmaf(xv[i], b, z[0]);
      if (ORDER >= 1) z[1 + i] = b;
      if (ORDER == 2 || ORDER == 4) z[1 + D + i] = 0.0f;
    }
    if (ORDER == 3) z[1 + D] = 0.0f;
    if (ORDER == 4) {
#pragma unroll
      for (int m = 0; m < D * (D - 1) / 2; ++m) z[1 + 2 * D + m] = 0.0f;
    }
    sincos_jets<D, ORDER>(z, sj, cj, lc);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      buf[(j * W + f) * TPP + p] = sj[j];
      buf[(j * W + m + f) * TPP + p] = cj[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER, int TP>
__global__ void __launch_bounds__(kThreads)
jet_fwd_ffma(const __grid_constant__ DevProg prog, const float* __restrict__ params,
             const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
             float* __restrict__ out, float* __restrict__ stash, int64_t n_pad, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr int TPP = TP + 1;
  constexpr int G = kThreads / TP;
  extern __shared__ __align__(16) float smem[];
  const int W = prog.max_w;
  const int plane = J * W * TPP;
  float* bufs[3] = {smem, smem + plane, smem + 2 * plane};
  float* Ws = smem + (prog.has_skip ? 3 : 2) * plane;
  Ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Ws) + 15) & ~uintptr_t(15));
  float* bs = Ws + W * W;

  const int p = threadIdx.x % TP;
  const int g = threadIdx.x / TP;
  const int o = prog.n_out;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * TP + p;
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;

    int cur = 0, skip = -1;
    write_input_jets<D, ORDER, TPP>(bufs[cur], W, xv, p, g, G);
    __syncthreads();

    for (int l = 0; l < prog.n_ops; ++l) {
      const DevOp& op = prog.ops[l];
      int nxt = 0;
      while (nxt == cur || nxt == skip) ++nxt;
      if (op.kind == MPINN_OP_FOURIER) {
        fourier_jets<D, ORDER, TPP>(bufs[nxt], W, params + op.w_off, op.out_dim / 2, xv, p, g, G, prog.lap_c);
        __syncthreads();
        cur = nxt;
        continue;
      }
      copy_to_smem(Ws, packed + op.pk_fwd, op.in_dim * op.outp);
      for (int i = threadIdx.x; i < op.out_dim; i += kThreads) bs[i] = __ldg(params + op.b_off + i);
      if (op.flags & MPINN_FLAG_SAVE_SKIP) skip = cur;
      __syncthreads();
      const float* hin = bufs[cur];
      const float* hskip = (op.flags & MPINN_FLAG_ADD_SKIP) ? bufs[skip] : nullptr;
      float* hout = bufs[nxt];
      const bool do_stash = (stash != nullptr) && (op.stash_plane >= 0);
      for (int c = g; c < op.outp / kCh; c += G) {
        float acc[kCh][J];
        gemm_chunk<J, TPP>(hin, Ws, op.in_dim, op.outp, c, W, p, acc);
#pragma unroll
        for (int f = 0; f < kCh; ++f) {
          const int fo = kCh * c + f;
          if (fo < op.out_dim) {
            float z[J];
#pragma unroll
            for (int j = 0; j < J; ++j) z[j] = acc[f][j];
            z[0] += bs[fo];
            if (hskip) {
#pragma unroll
              for (int j = 0; j < J; ++j) z[j] += hskip[(j * W + fo) * TPP + p];
            }
            if (do_stash) {
#pragma unroll
              for (int j = 0; j < J; ++j)
                stash[(int64_t)(op.stash_plane + j * op.out_dim + fo) * n_pad + pt] = z[j];
            }
            act_jets<D, ORDER>(op.act, z, prog.lap_c);
#pragma unroll
            for (int j = 0; j < J; ++j) hout[(j * W + fo) * TPP + p] = z[j];
          }
        }
      }
      if (op.flags & MPINN_FLAG_ADD_SKIP) skip = -1;
      __syncthreads();
      cur = nxt;
    }
    // network output jets -> out[J][n][o]
    if (valid) {
      for (int idx = g; idx < J * o; idx += G) {
        const int j = idx / o, k = idx % o;
        out[((int64_t)j * n + pt) * o + k] = bufs[cur][(j * W + k) * TPP + p];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// reverse
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER, int TP>
__global__ void __launch_bounds__(kThreads)
jet_bwd_ffma(const __grid_constant__ DevProg prog, const float* __restrict__ params,
             const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
             const float* __restrict__ dout, const float* __restrict__ stash, int64_t n_pad,
             float* __restrict__ part_all, int64_t n_params, float* __restrict__ dx, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr int TPP = TP + 1;
  constexpr int G = kThreads / TP;
  extern __shared__ __align__(16) float smem[];
  const int W = prog.max_w;
  const int plane = J * W * TPP;
  float* bufG = smem;              // grad wrt current op's output (then z-bar, in place)
  float* bufC = smem + plane;      // grad wrt current op's input
  float* bufH = smem + 2 * plane;  // current op's input activations (recomputed)
  float* bufS = smem + 3 * plane;  // skip-branch gradient (only if has_skip)
  float* Ws = smem + (prog.has_skip ? 4 : 3) * plane;
  Ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Ws) + 15) & ~uintptr_t(15));

  float* part = part_all + (int64_t)blockIdx.x * n_params;
  for (int64_t i = threadIdx.x; i < n_params; i += kThreads) part[i] = 0.0f;
  __syncthreads();

  const int p = threadIdx.x % TP;
  const int g = threadIdx.x / TP;
  const int lane = threadIdx.x % 32;
  const int warp = threadIdx.x / 32;
  const int o = prog.n_out;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * TP + p;
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;

    for (int idx = g; idx < J * o; idx += G) {
      const int j = idx / o, k = idx % o;
      bufG[(j * W + k) * TPP + p] = valid ? __ldg(dout + ((int64_t)j * n + pt) * o + k) : 0.0f;
    }
    __syncthreads();

    for (int l = prog.n_ops - 1; l >= 0; --l) {
      const DevOp& op = prog.ops[l];
      if (op.kind == MPINN_OP_FOURIER) {
        // A frozen encoding needs nothing upstream.  A trainable one (reference fourier_features.py:30-32): bufG holds
        // abar for [sin(xB) | cos(xB)]; both halves flow back into z = xB, and dL/dB[i][f] = sum over points of
        // zbar_0[f] x_i + zbar_{1+i}[f]  (z's first-derivative channel i is B[i][f] itself, the second-order ones 0).
        // With dx requested the same zbar_0 also gives dL/dx_i = sum_f zbar_0[f] B[i][f] (x enters z_0 = xB only).
        const bool train = (op.flags & MPINN_FLAG_TRAINABLE) != 0;
        if (!train && dx == nullptr) break;
        const int m = op.out_dim / 2;
        const float* B = params + op.w_off;
        float dxp[D];
#pragma unroll
        for (int i = 0; i < D; ++i) dxp[i] = 0.0f;
        for (int f = g; f < m; f += G) {
          float z[J], as[J], ac[J];
          z[0] = 0.0f;
#pragma unroll
          for (int j = 1; j < J; ++j) z[j] = 0.0f;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            const float b = __ldg(B + i * m + f);
            z[0] = fmaf(xv[i], b, z[0]);
            if (ORDER >= 1) z[1 + i] = b;
          }
#pragma unroll
          for (int j = 0; j < J; ++j) {
            as[j] = bufG[(j * W + f) * TPP + p];
            ac[j] = bufG[(j * W + m + f) * TPP + p];
          }
          act_jets_bwd<D, ORDER>(MPINN_ACT_SINE, z, as, prog.lap_c);
          act_jets_bwd<D, ORDER>(kActCos, z, ac, prog.lap_c);
#pragma unroll
          for (int i = 0; i < D; ++i) dxp[i] = fmaf(as[0] + ac[0], __ldg(B + i * m + f), dxp[i]);
          if (valid && train) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
              float gb = (as[0] + ac[0]) * xv[i];
              if (ORDER >= 1) gb += as[1 + i] + ac[1 + i];
              atomicAdd(part + op.w_off + i * m + f, gb);
            }
          }
        }
        if (dx != nullptr) {     // sum the G feature groups' partials of every point (bufC is free at op 0)
#pragma unroll
          for (int i = 0; i < D; ++i) bufC[(i * G + g) * TPP + p] = dxp[i];
          __syncthreads();
          if (g < D && valid) {
            float s = 0.0f;
            for (int q = 0; q < G; ++q) s += bufC[(g * G + q) * TPP + p];
            dx[pt * D + g] = s;
          }
        }
        break;
      }

      // (1) z-bar = act'(z) applied to a-bar, in place; keep a copy for the skip branch
      if (op.act != MPINN_ACT_NONE || (op.flags & MPINN_FLAG_ADD_SKIP)) {
        for (int f = g; f < op.out_dim; f += G) {
          float ab[J];
#pragma unroll
          for (int j = 0; j < J; ++j) ab[j] = bufG[(j * W + f) * TPP + p];
          if (op.act != MPINN_ACT_NONE) {
            float z[J];
#pragma unroll
            for (int j = 0; j < J; ++j)
              z[j] = __ldg(stash + (int64_t)(op.stash_plane + j * op.out_dim + f) * n_pad + pt);
            act_jets_bwd<D, ORDER>(op.act, z, ab, prog.lap_c);
#pragma unroll
            for (int j = 0; j < J; ++j) bufG[(j * W + f) * TPP + p] = ab[j];
          }
          if (op.flags & MPINN_FLAG_ADD_SKIP) {
#pragma unroll
            for (int j = 0; j < J; ++j) bufS[(j * W + f) * TPP + p] = ab[j];
          }
        }
      }
      // (2) recompute this op's input activations into bufH
      if (l == 0) {
        write_input_jets<D, ORDER, TPP>(bufH, W, xv, p, g, G);
      } else {
        const DevOp& pv = prog.ops[l - 1];
        if (pv.kind == MPINN_OP_FOURIER) {
          fourier_jets<D, ORDER, TPP>(bufH, W, params + pv.w_off, pv.out_dim / 2, xv, p, g, G, prog.lap_c);
        } else {
          for (int f = g; f < pv.out_dim; f += G) {
            float z[J];
#pragma unroll
            for (int j = 0; j < J; ++j)
              z[j] = __ldg(stash + (int64_t)(pv.stash_plane + j * pv.out_dim + f) * n_pad + pt);
            act_jets<D, ORDER>(pv.act, z, prog.lap_c);
#pragma unroll
            for (int j = 0; j < J; ++j) bufH[(j * W + f) * TPP + p] = z[j];
          }
        }
      }
      const bool need_dinput = (l > 0) && !(prog.ops[l - 1].kind == MPINN_OP_FOURIER && dx == nullptr &&
                                            !(prog.ops[l - 1].flags & MPINN_FLAG_TRAINABLE));
      if (need_dinput) copy_to_smem(Ws, packed + op.pk_bwd, op.out_dim * op.inp);
      __syncthreads();

      // (3) dW[o][i] += sum_{j,p} zbar[j][o][p] * h[j][i][p] ; db[o] += sum_p zbar[0][o][p]
      for (int oc = warp; oc < op.outp / 8; oc += kThreads / 32) {
        for (int ib = 0; ib < op.in_dim; ib += 64) {
          const int i0 = ib + lane, i1 = ib + 32 + lane;
          const int i0c = i0 < op.in_dim ? i0 : 0, i1c = i1 < op.in_dim ? i1 : 0;
          float a0[8], a1[8];
#pragma unroll
          for (int f = 0; f < 8; ++f) a0[f] = a1[f] = 0.0f;
          for (int j = 0; j < J; ++j) {
            const float* zr = bufG + (j * W + 8 * oc) * TPP;
            const float* h0 = bufH + (j * W + i0c) * TPP;
            const float* h1 = bufH + (j * W + i1c) * TPP;
#pragma unroll 4
            for (int q = 0; q < TP; ++q) {
              const float v0 = h0[q], v1 = h1[q];
#pragma unroll
              for (int f = 0; f < 8; ++f) {
                const float zb = zr[f * TPP + q];
                a0[f] = fmaf(zb, v0, a0[f]);
                a1[f] = fmaf(zb, v1, a1[f]);
              }
            }
          }
#pragma unroll
          for (int f = 0; f < 8; ++f) {
            const int fo = 8 * oc + f;
            if (fo < op.out_dim) {
              // fire-and-forget reductions into the CTA's private slab (a load-add-store here exposed one L2
              // round trip per tile and was 29 % of this kernel's stall samples on 128-wide layers)
              if (i0 < op.in_dim) atomicAdd(part + op.w_off + fo * op.in_dim + i0, a0[f]);
              if (i1 < op.in_dim) atomicAdd(part + op.w_off + fo * op.in_dim + i1, a1[f]);
            }
          }
        }
      }
      for (int f = threadIdx.x; f < op.out_dim; f += kThreads) {
        float s = 0.0f;
        for (int q = 0; q < TP; ++q) s += bufG[f * TPP + q];
        atomicAdd(part + op.b_off + f, s);
      }
      // gradient w.r.t. the points: x enters only through the value channel of op 0 (z = W x + b), so
      // dL/dx_i = sum_o W[o][i] zbar_0[o]  (a Fourier first op: see its branch above)
      if (l == 0 && dx != nullptr && op.kind == MPINN_OP_LINEAR && g < D && pt < n) {
        float s = 0.0f;
        for (int oo = 0; oo < op.out_dim; ++oo) s = fmaf(__ldg(params + op.w_off + oo * D + g), bufG[oo * TPP + p], s);
        dx[pt * D + g] = s;
      }

      // (4) grad wrt this op's input: C[j][i][p] = sum_o W[o][i] zbar[j][o][p] (+ skip branch)
      if (need_dinput) {
        for (int c = g; c < op.inp / kCh; c += G) {
          float acc[kCh][J];
          gemm_chunk<J, TPP>(bufG, Ws, op.out_dim, op.inp, c, W, p, acc);
#pragma unroll
          for (int f = 0; f < kCh; ++f) {
            const int fi = kCh * c + f;
            if (fi < op.in_dim) {
#pragma unroll
              for (int j = 0; j < J; ++j) {
                float v = acc[f][j];
                if (op.flags & MPINN_FLAG_SAVE_SKIP) v += bufS[(j * W + fi) * TPP + p];
                bufC[(j * W + fi) * TPP + p] = v;
              }
            }
          }
        }
      }
      __syncthreads();
      if (!need_dinput) break;
      float* t = bufG;
      bufG = bufC;
      bufC = t;
    }
    __syncthreads();
  }
}

// dparams[i] += sum_c part[c][i]
__global__ void fold_partials(const float* __restrict__ part, int n_slabs, int64_t n_params,
                              float* __restrict__ dparams, const double* __restrict__ bsum, int bsum_off, int n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  float s = 0.0f;
  if (!take_bsum(bsum, bsum_off, n_out, i, &s))
    for (int c = 0; c < n_slabs; ++c) s += part[(int64_t)c * n_params + i];
  dparams[i] += s;
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static size_t ffma_smem_bytes(const DevProg& prog, int J, int TP, int n_bufs) {
  const size_t plane = (size_t)J * prog.max_w * (TP + 1);
  return (n_bufs * plane + (size_t)prog.max_w * prog.max_w + prog.max_w) * sizeof(float) + 32;
}

template <int D, int ORDER, int TP>
static cudaError_t launch_fwd_t(const DevProg& prog, const float* params, const float* packed,
                                const float* x, int64_t n, float* out, float* stash, int64_t n_pad,
                                int sm_count, cudaStream_t stream) {
  constexpr int J = jet_count(D, ORDER);
  const size_t smem = ffma_smem_bytes(prog, J, TP, prog.has_skip ? 3 : 2);
  auto kern = jet_fwd_ffma<D, ORDER, TP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  const int n_tiles = (int)((n + TP - 1) / TP);
  int grid = sm_count * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  kern<<<grid, kThreads, smem, stream>>>(prog, params, packed, x, n, out, stash, n_pad, n_tiles);
  return cudaGetLastError();
}

template <int D, int ORDER, int TP>
static cudaError_t launch_bwd_t(const DevProg& prog, const float* params, const float* packed,
                                const float* x, int64_t n, const float* dout, const float* stash,
                                int64_t n_pad, float* part, int max_slabs, int64_t n_params,
                                float* dparams, float* dx, int sm_count, cudaStream_t stream) {
  constexpr int J = jet_count(D, ORDER);
  const size_t smem = ffma_smem_bytes(prog, J, TP, prog.has_skip ? 4 : 3);
  auto kern = jet_bwd_ffma<D, ORDER, TP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  const int n_tiles = (int)((n + TP - 1) / TP);
  int grid = sm_count * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > max_slabs) grid = max_slabs;
  kern<<<grid, kThreads, smem, stream>>>(prog, params, packed, x, n, dout, stash, n_pad, part,
                                         n_params, dx, n_tiles);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int fb = 256;
  fold_partials<<<(unsigned)((n_params + fb - 1) / fb), fb, 0, stream>>>(part, grid, n_params, dparams, prog.bsum,
                                                                          prog.bsum_off, prog.n_out);
  return cudaGetLastError();
}

static int pick_tp(const DevProg& prog, int J, int n_bufs, size_t smem_limit) {
  if (ffma_smem_bytes(prog, J, 32, n_bufs) <= smem_limit) return 32;
  if (ffma_smem_bytes(prog, J, 16, n_bufs) <= smem_limit) return 16;
  return 0;
}

#define MPINN_DISPATCH(D_, O_, TP_, CALL)                                  \
  if (d == D_ && order == O_ && tp == TP_) return CALL<D_, O_, TP_>

cudaError_t launch_fwd_ffma(const DevProg& prog, const float* params, const float* packed,
                            const float* x, int64_t n, int d, int order, float* out, float* stash,
                            int64_t n_pad, int sm_count, size_t smem_limit, cudaStream_t stream) {
  const int J = jet_count(d, order);
  // the reverse pass must agree on TP (the stash is TP-independent, but keep both kernels on the
  // same tiling so their smem budgets are checked together)
  const int tp = pick_tp(prog, J, prog.has_skip ? 4 : 3, smem_limit);
  if (tp == 0) return cudaErrorInvalidConfiguration;
#define FWD_ARGS (prog, params, packed, x, n, out, stash, n_pad, sm_count, stream)
  MPINN_DISPATCH(1, 0, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 1, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 2, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 0, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 1, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 2, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 0, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 1, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 2, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 3, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 3, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 3, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 4, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 4, 32, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 0, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 1, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 2, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 0, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 1, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 2, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 0, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 1, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 2, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(1, 3, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 3, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 3, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(2, 4, 16, launch_fwd_t) FWD_ARGS;
  MPINN_DISPATCH(3, 4, 16, launch_fwd_t) FWD_ARGS;
#undef FWD_ARGS
  return cudaErrorInvalidValue;
}

cudaError_t launch_bwd_ffma(const DevProg& prog, const float* params, const float* packed,
                            const float* x, int64_t n, int d, int order, const float* dout,
                            const float* stash, int64_t n_pad, float* part, int max_slabs,
                            int64_t n_params, float* dparams, float* dx, int sm_count,
                            size_t smem_limit, cudaStream_t stream) {
  const int J = jet_count(d, order);
  const int tp = pick_tp(prog, J, prog.has_skip ? 4 : 3, smem_limit);
  if (tp == 0) return cudaErrorInvalidConfiguration;
#define BWD_ARGS \
  (prog, params, packed, x, n, dout, stash, n_pad, part, max_slabs, n_params, dparams, dx, sm_count, stream)
  MPINN_DISPATCH(1, 0, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 1, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 2, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 0, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 1, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 2, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 0, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 1, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 2, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 3, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 3, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 3, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 4, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 4, 32, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 0, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 1, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 2, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 0, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 1, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 2, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 0, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 1, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 2, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(1, 3, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 3, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 3, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(2, 4, 16, launch_bwd_t) BWD_ARGS;
  MPINN_DISPATCH(3, 4, 16, launch_bwd_t) BWD_ARGS;
#undef BWD_ARGS
  return cudaErrorInvalidValue;
}

}  // namespace mpinn
