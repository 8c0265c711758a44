// jet_tc_wide.cu -- tcgen05 engine for plain FNN programs with 128-wide hidden layers (burgers_1D, reference
// examples/burgers_1D/configs/config.yaml:7), up to four jet channels.
//
// The 64-wide engine (jet_tc.cu / jet_tc_bwd.cu) keeps a whole network on chip; at 128 features neither its
// shared-memory operand tiles nor its tensor-memory accumulators fit.  This engine is
// LAYER-WISE instead: one persistent kernel per layer and pass, pre-activation jets travel through HBM planes
// (the stash the reverse pass needs anyway), and the jet channels are stacked along the M dimension of the MMA:
//
//     tile = 32 collocation points x 4 channels = 128 rows (row 32 j + p  <->  channel j, point p)
//     forward   Z[128 x 128]  = act_jets(Z_prev)[128 x 128] * W^T        K-major operands, 48 MMAs M128 N128 K16
//     reverse   Abar[128x128] = Zbar[128 x 128] * W                      K-major operands, 48 MMAs
//               dW[128 x 128] += Zbar^T * act_jets(Z_prev)               MN-major views of the same tile layout, K = rows
//
// so one weight operand serves all channels of a tile, the activation coefficients are evaluated once per (point,
// feature, layer) with every channel in registers, and the weight-gradient GEMM sums over points AND channels in
// the tensor core.  All products are BF16x3 (three round-to-nearest bf16 pieces, six piece products, smallest
// first, fp32 accumulation in tensor memory) like the 64-wide reverse pass.  The first (d -> 128) and last
// (128 -> o) layers are HBM-bound elementwise / reduction kernels on the CUDA cores.
//
// Reference behaviour implemented: FNN.forward (neural_network/feedforward_neural_network.py:50-52), grad()
// sweeps (condition/diff.py:47), loss.backward() (trainer/trainer.py:141) -- restated as jets, oracle/jet_ref.py.
#include "jet_tc_common.cuh"

namespace mpinn {

constexpr int kWd = 128;                      // layer width of this engine
constexpr int kWdPts = 32;                    // points per tile (x 4 channel slots = 128 MMA rows)
constexpr int kWdThreads = 512;               // thread (p = tid % 32, fg = tid / 32): point p, features 8 fg .. 8 fg + 7
constexpr int kWdRegion = 128 * 128;          // [128 rows x 64 bf16], 128B swizzle: 16 KiB
constexpr int kWdTile = 6 * kWdRegion;        // [piece 0..2][feature half 0..1]: 96 KiB
constexpr int kWdSmem = 2 * kWdTile + 1024;   // two tiles (activation + weight, or Zbar + H) + alignment slack
constexpr int kWdMaxOut = 4;
constexpr int kWdFlushTiles = 2;              // weight-gradient accumulator leaves tensor memory every this many tiles

int64_t wide_block_floats() { return 2 * kWdTile / 4; }   // forward image (B[n=out][k=in]) then reverse image (B[n=in][k=out])

bool wide_supported(const DevProg& prog, int d, int order) {
  if (order > 3) return false;
  if (jet_count(d, order) > 4) return false;
  if (prog.n_ops < 3 || prog.has_skip || prog.n_out > kWdMaxOut) return false;
  const DevOp& first = prog.ops[0];
  if (first.kind != MPINN_OP_LINEAR || first.in_dim != d || first.out_dim != kWd || first.flags) return false;
  for (int l = 1; l < prog.n_ops - 1; ++l) {
    const DevOp& op = prog.ops[l];
    if (op.kind != MPINN_OP_LINEAR || op.in_dim != kWd || op.out_dim != kWd || op.tc_off < 0 || op.flags) return false;
  }
  const DevOp& last = prog.ops[prog.n_ops - 1];
  return last.kind == MPINN_OP_LINEAR && last.in_dim == kWd && last.act == MPINN_ACT_NONE && !last.flags;
}

// extra stash floats behind the pre-activation planes: two plane sets for the reverse pass's Zbar ping-pong
int64_t wide_scratch_floats(const DevProg& prog, int d, int order, int64_t n_pad) {
  return wide_supported(prog, d, order) ? 2 * (int64_t)jet_count(d, order) * kWd * n_pad : 0;
}

// ---------------------------------------------------------------------------------------------
// weight operand images
// ---------------------------------------------------------------------------------------------
__global__ void pack_wide_kernel(const __grid_constant__ DevProg prog, const float* __restrict__ params,
                                 float* __restrict__ packed) {
  const DevOp& op = prog.ops[blockIdx.y];
  if (op.kind != MPINN_OP_LINEAR || op.tc_off < 0 || op.in_dim != kWd || op.out_dim != kWd) return;
  const float* W = params + op.w_off;  // [out][in]
  char* dst = reinterpret_cast<char*>(packed + op.tc_off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kWd * kWd; i += gridDim.x * blockDim.x) {
    const int o = i / kWd, k = i % kWd;
    const float w = W[i];
    const __nv_bfloat16 b1 = __float2bfloat16_rn(w);
    const float r1 = w - __bfloat162float(b1);
    const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
    const __nv_bfloat16 b[3] = {b1, b2, b3};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      // forward operand: row n = out index, K = in index
      *reinterpret_cast<__nv_bfloat16*>(dst + (q * 2 + k / 64) * kWdRegion + sw128_offset_bf16(o, k % 64)) = b[q];
      // reverse operand: row n = in index, K = out index
      *reinterpret_cast<__nv_bfloat16*>(dst + kWdTile + (q * 2 + o / 64) * kWdRegion + sw128_offset_bf16(k, o % 64)) = b[q];
    }
  }
}

cudaError_t launch_pack_wide(const DevProg& prog, const float* params, float* packed, cudaStream_t stream) {
  bool any = false;
  for (int l = 0; l < prog.n_ops; ++l)
    any |= prog.ops[l].tc_off >= 0 && prog.ops[l].in_dim == kWd && prog.ops[l].out_dim == kWd;
  if (!any) return cudaSuccess;
  dim3 grid(8, prog.n_ops);
  pack_wide_kernel<<<grid, 256, 0, stream>>>(prog, params, packed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
// plane set of one op: [tile][channel][feature][32 points]
template <int J>
__device__ __forceinline__ int64_t plane_off(int64_t tile, int j, int f, int p) {
  return ((tile * J + j) * kWd + f) * kWdPts + p;
}

__device__ __forceinline__ void red_add_v4w(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// eight consecutive features of one row -> one 16-byte chunk in each of the three piece regions
__device__ __forceinline__ void stage_row8(uint8_t* tile, int row, int fg, const float (&v)[8]) {
  uint32_t p1[4], p2[4], p3[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16x3(v[2 * i], v[2 * i + 1], p1[i], p2[i], p3[i]);
  const uint32_t off = (uint32_t)(fg >> 3) * kWdRegion + (uint32_t)row * 128u + ((((uint32_t)fg & 7u) ^ ((uint32_t)row & 7u)) << 4);
  *reinterpret_cast<uint4*>(tile + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
  *reinterpret_cast<uint4*>(tile + 2 * kWdRegion + off) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  *reinterpret_cast<uint4*>(tile + 4 * kWdRegion + off) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
}

// the same chunk from pieces split earlier (the pipelined forward kernel splits while the previous tile's MMAs run)
__device__ __forceinline__ void stage_pieces8(uint8_t* tile, int row, int fg, const uint32_t (&p1)[4],
                                              const uint32_t (&p2)[4], const uint32_t (&p3)[4]) {
  const uint32_t off = (uint32_t)(fg >> 3) * kWdRegion + (uint32_t)row * 128u + ((((uint32_t)fg & 7u) ^ ((uint32_t)row & 7u)) << 4);
  *reinterpret_cast<uint4*>(tile + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
  *reinterpret_cast<uint4*>(tile + 2 * kWdRegion + off) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  *reinterpret_cast<uint4*>(tile + 4 * kWdRegion + off) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
}

// D[128 x 128] = A[128 rows x 128] * B[128 x 128]^T, K-major bf16x3 tiles, smallest piece products first
__device__ __forceinline__ void issue_wide_kmajor(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
  constexpr int pa[6] = {2, 1, 0, 1, 0, 0};
  constexpr int pb[6] = {0, 1, 2, 0, 1, 0};
  const uint64_t a0 = make_desc_sw128(a_tile, 16, 1024);
  const uint64_t b0 = make_desc_sw128(b_tile, 16, 1024);
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {   // K = 16 bf16 = 32 bytes per MMA; k-steps 4..7 live in the second feature half
      const uint32_t ao = (pa[q] * 2 + (ks >> 2)) * kWdRegion + (ks & 3) * 32;
      const uint32_t bo = (pb[q] * 2 + (ks >> 2)) * kWdRegion + (ks & 3) * 32;
      mma_bf16_ss(d_tmem, a0 + (ao >> 4), b0 + (bo >> 4), idesc, acc);
      acc = 1;
    }
  }
}

// D[128 x 128] (+)= Z^T H over the 128 rows of the two tiles (MN-major views: M / N halves one region apart)
__device__ __forceinline__ void issue_wide_mnmajor(uint32_t d_tmem, uint32_t z_tile, uint32_t h_tile, uint32_t accumulate_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);
  constexpr int pa[6] = {2, 1, 0, 1, 0, 0};
  constexpr int pb[6] = {0, 1, 2, 0, 1, 0};
  const uint64_t z0 = make_desc_sw128(z_tile, kWdRegion, 1024);
  const uint64_t h0 = make_desc_sw128(h_tile, kWdRegion, 1024);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {   // K = 16 rows = two 8-row swizzle atoms
      mma_bf16_ss(d_tmem, z0 + ((pa[q] * 2 * kWdRegion + ks * 2048) >> 4), h0 + ((pb[q] * 2 * kWdRegion + ks * 2048) >> 4),
                  idesc, acc);
      acc = 1;
    }
  }
}

#ifndef MPINN_WIDE_PREFETCH
#define MPINN_WIDE_PREFETCH 1   // software pipelining of the plane loads (next tile's loads in flight under the MMAs)
#endif

__device__ __forceinline__ void l2_prefetch_bulk_w(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void copy_image(uint8_t* dst, const float* __restrict__ src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < kWdTile / 16; i += kWdThreads) d4[i] = __ldg(s4 + i);
}

// ---------------------------------------------------------------------------------------------
// first layer (d -> 128) and last layer (128 -> o): CUDA-core kernels
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER>
__global__ void __launch_bounds__(kWdThreads)
wide_first_fwd(const __grid_constant__ DevProg prog, const float* __restrict__ params, const float* __restrict__ x,
               int64_t n, float* __restrict__ z0, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  const DevOp& op = prog.ops[0];
  const int p = threadIdx.x & 31, fg = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = tile * kWdPts + p;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = pt < n ? x[pt * D + i] : 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = 8 * fg + k;
      float z[J];
#pragma unroll
      for (int j = 0; j < J; ++j) z[j] = 0.0f;
      z[0] = __ldg(params + op.b_off + f);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float w = __ldg(params + op.w_off + f * D + i);
        z[0] = fmaf(w, xv[i], z[0]);
        if (ORDER >= 1) z[1 + i] = w;
      }
#pragma unroll
      for (int j = 0; j < J; ++j) z0[plane_off<J>(tile, j, f, p)] = z[j];
    }
  }
}

// dW0, db0 from Zbar_0 (the reverse of wide_first_fwd); per-thread sums over the CTA's tiles, one reduction at the end
template <int D, int ORDER>
__global__ void __launch_bounds__(kWdThreads)
wide_first_bwd(const __grid_constant__ DevProg prog, const float* __restrict__ x, int64_t n,
               const float* __restrict__ zb0, float* __restrict__ part, int64_t slab_floats, int n_slabs, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  const DevOp& op = prog.ops[0];
  float* slab = part + (int64_t)(blockIdx.x % n_slabs) * slab_floats;   // HBM-bound: several CTAs per SM share the slabs
  const int p = threadIdx.x & 31, fg = threadIdx.x >> 5;
  float accw[8][D], accb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    accb[k] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) accw[k][i] = 0.0f;
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = tile * kWdPts + p;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = pt < n ? x[pt * D + i] : 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = 8 * fg + k;
      const float g0 = zb0[plane_off<J>(tile, 0, f, p)];
      accb[k] += g0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float g = g0 * xv[i];
        if (ORDER >= 1) g += zb0[plane_off<J>(tile, 1 + i, f, p)];
        accw[k][i] += g;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], off);
#pragma unroll
      for (int i = 0; i < D; ++i) accw[k][i] += __shfl_xor_sync(0xffffffffu, accw[k][i], off);
    }
    if (p == 0) {
      const int f = 8 * fg + k;
      atomicAdd(slab + op.b_off + f, accb[k]);
#pragma unroll
      for (int i = 0; i < D; ++i) atomicAdd(slab + op.w_off + f * D + i, accw[k][i]);
    }
  }
}

// out[j][pt][o] = sum_f Wl[o][f] act_jets(z)[j][f] + b[o] (j = 0)
template <int D, int ORDER, int NO>   // NO: compile-time bound of the output count (1 or kWdMaxOut)
__global__ void __launch_bounds__(kWdThreads)
wide_last_fwd(const __grid_constant__ DevProg prog, const float* __restrict__ params, const float* __restrict__ zin,
              int64_t n, float* __restrict__ out, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  const DevOp& op = prog.ops[prog.n_ops - 1];
  const int act = prog.ops[prog.n_ops - 2].act;
  const int no = op.out_dim;
  __shared__ float red[16][NO * J][kWdPts + 1];
  const int p = threadIdx.x & 31, fg = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#if MPINN_WIDE_PREFETCH
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles)
      l2_prefetch_bulk_w(zin + plane_off<J>(tile + gridDim.x, 0, 0, 0), (uint32_t)(J * kWd * kWdPts * 4));
#endif
    float acc[NO][J];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int j = 0; j < J; ++j) acc[o][j] = 0.0f;
    float zv[J][8];   // all of the tile's loads in flight at once
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) zv[j][k] = zin[plane_off<J>(tile, j, 8 * fg + k, p)];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = 8 * fg + k;
      float z[J];
#pragma unroll
      for (int j = 0; j < J; ++j) z[j] = zv[j][k];
      act_jets<D, ORDER>(act, z, prog.lap_c);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if (o < no) {
          const float w = __ldg(params + op.w_off + o * kWd + f);
#pragma unroll
          for (int j = 0; j < J; ++j) acc[o][j] = fmaf(w, z[j], acc[o][j]);
        }
      }
    }
    __syncthreads();   // previous tile's readers are done
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int j = 0; j < J; ++j) red[fg][o * J + j][p] = acc[o][j];
    __syncthreads();
    for (int i = threadIdx.x; i < no * J * kWdPts; i += kWdThreads) {
      const int pp = i % kWdPts, oj = i / kWdPts, o = oj / J, j = oj % J;
      float s = 0.0f;
#pragma unroll
      for (int g = 0; g < 16; ++g) s += red[g][o * J + j][pp];
      if (j == 0) s += __ldg(params + op.b_off + o);
      const int64_t pt = tile * kWdPts + pp;
      if (pt < n) out[((int64_t)j * n + pt) * no + o] = s;
    }
  }
}

// reverse of the last layer: Zbar of the last hidden op (written to zb), dWl, dbl
template <int D, int ORDER, int NO>
__global__ void __launch_bounds__(kWdThreads)
wide_last_bwd(const __grid_constant__ DevProg prog, const float* __restrict__ params, const float* __restrict__ zin,
              const float* __restrict__ dout, int64_t n, float* __restrict__ zb, float* __restrict__ part,
              int64_t slab_floats, int n_slabs, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  const DevOp& op = prog.ops[prog.n_ops - 1];
  const int act = prog.ops[prog.n_ops - 2].act;
  const int no = op.out_dim;
  float* slab = part + (int64_t)(blockIdx.x % n_slabs) * slab_floats;   // CTAs beyond the slab count share (atomics)
  const int p = threadIdx.x & 31, fg = threadIdx.x >> 5;
  float accw[NO][8], accb[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    accb[o] = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[o][k] = 0.0f;
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = tile * kWdPts + p;
#if MPINN_WIDE_PREFETCH
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles)
      l2_prefetch_bulk_w(zin + plane_off<J>(tile + gridDim.x, 0, 0, 0), (uint32_t)(J * kWd * kWdPts * 4));
#endif
    float g[NO][J];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int j = 0; j < J; ++j) g[o][j] = (o < no && pt < n) ? dout[((int64_t)j * n + pt) * no + o] : 0.0f;
#pragma unroll
    for (int o = 0; o < NO; ++o) accb[o] += g[o][0];
    float zv[J][8];   // all of the tile's loads in flight at once
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) zv[j][k] = zin[plane_off<J>(tile, j, 8 * fg + k, p)];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = 8 * fg + k;
      float z[J], a[J], ab[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        z[j] = zv[j][k];
        a[j] = z[j];
        ab[j] = 0.0f;
      }
      act_jets<D, ORDER>(act, a, prog.lap_c);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if (o < no) {
          const float w = __ldg(params + op.w_off + o * kWd + f);
#pragma unroll
          for (int j = 0; j < J; ++j) {
            ab[j] = fmaf(g[o][j], w, ab[j]);
            accw[o][k] = fmaf(g[o][j], a[j], accw[o][k]);
          }
        }
      }
      act_jets_bwd<D, ORDER>(act, z, ab, prog.lap_c);
#pragma unroll
      for (int j = 0; j < J; ++j) zb[plane_off<J>(tile, j, f, p)] = ab[j];
    }
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    if (o < no) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        accb[o] += __shfl_xor_sync(0xffffffffu, accb[o], off);
#pragma unroll
        for (int k = 0; k < 8; ++k) accw[o][k] += __shfl_xor_sync(0xffffffffu, accw[o][k], off);
      }
      if (p == 0) {
        if (fg == 0) atomicAdd(slab + op.b_off + o, accb[o]);
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(slab + op.w_off + o * kWd + 8 * fg + k, accw[o][k]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// hidden layers on the tensor core
// ---------------------------------------------------------------------------------------------
// MODE 0 (forward, op l):          zout = act_jets(zin) W^T + b      zin = planes of op l-1, zout = planes of op l
// MODE 1 (input gradient, op l):   gout = act_jets_bwd(zprev; gin W) gin = Zbar_l, zprev = planes of op l-1, gout = Zbar_{l-1}
template <int D, int ORDER, int MODE>
__global__ void __launch_bounds__(kWdThreads, 1)
wide_layer_kernel(const __grid_constant__ DevProg prog, int l, const float* __restrict__ params,
                  const float* __restrict__ packed, const float* __restrict__ src, const float* __restrict__ zprev,
                  float* __restrict__ dst, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* a_tile = smem;
  uint8_t* w_tile = smem + kWdTile;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  const int p = t & 31, fg = t >> 5;          // staging role
  const int lq = fg & 3, cq = fg >> 2;        // accumulator role: lanes 32 lq .. (channel lq), columns 32 cq ..
  const DevOp& op = prog.ops[l];
  const int act_below = prog.ops[l - 1].act;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 128);
  copy_image(w_tile, packed + op.tc_off + (MODE == 0 ? 0 : kWdTile / 4));
  // rows of unused channel slots: zero once (their products land in unused accumulator lanes)
  for (int i = t; i < kWdTile / 16; i += kWdThreads) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;
  constexpr uint32_t kTileBytes = (uint32_t)(J * kWd * kWdPts * 4);   // one tile's planes are contiguous
  float v[J][8];
#if MPINN_WIDE_PREFETCH
  if (blockIdx.x < n_tiles) {
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(blockIdx.x, j, 8 * fg + k, p)];
  }
#endif
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- stage the A operand: every channel of 8 features of one point
    {
#if MPINN_WIDE_PREFETCH
      const int64_t nxt = tile + gridDim.x;
      if (t == 0) {   // pull the tile after this one (and this tile's z planes for the epilogue) towards L2
        if (MODE == 1) l2_prefetch_bulk_w(zprev + plane_off<J>(tile, 0, 0, 0), kTileBytes);
        if (nxt + gridDim.x < n_tiles) l2_prefetch_bulk_w(src + plane_off<J>(nxt + gridDim.x, 0, 0, 0), kTileBytes);
      }
#else
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(tile, j, 8 * fg + k, p)];
#endif
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float z[J];
#pragma unroll
          for (int j = 0; j < J; ++j) z[j] = v[j][k];
          act_jets<D, ORDER>(act_below, z, prog.lap_c);
#pragma unroll
          for (int j = 0; j < J; ++j) v[j][k] = z[j];
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) stage_row8(a_tile, 32 * j + p, fg, v[j]);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      issue_wide_kmajor(tmem, smem_u32(a_tile), smem_u32(w_tile));
      mma_commit(&bar);
    }
#if MPINN_WIDE_PREFETCH
    {   // the next tile's planes travel while the tensor core works
      const int64_t nxt = tile + gridDim.x;
      if (nxt < n_tiles) {
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(nxt, j, 8 * fg + k, p)];
      }
    }
#endif
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after_sync();
    // ---- accumulator -> result planes
    if (MODE == 0) {
      if (lq < J) {
        float acc[32];
        tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + cq * 32, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int f = 32 * cq + i;
          const float b = lq == 0 ? __ldg(params + op.b_off + f) : 0.0f;
          dst[plane_off<J>(tile, lq, f, p)] = acc[i] + b;
        }
      }
      tc_fence_before_sync();
      __syncthreads();
    } else {
      // the activation's reverse needs every channel of a (point, feature): exchange through shared memory
      // (the A tile is free once the MMAs have completed), X[feature][row]
      float* X = reinterpret_cast<float*>(a_tile);
      if (lq < J) {
        float acc[32];
        tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + cq * 32, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) X[(32 * cq + i) * 128 + 32 * lq + p] = acc[i];
      }
      tc_fence_before_sync();
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int f = 8 * fg + k;
        float z[J], ab[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          ab[j] = X[f * 128 + 32 * j + p];
          z[j] = zprev[plane_off<J>(tile, j, f, p)];
        }
        act_jets_bwd<D, ORDER>(act_below, z, ab, prog.lap_c);
#pragma unroll
        for (int j = 0; j < J; ++j) dst[plane_off<J>(tile, j, f, p)] = ab[j];
      }
      __syncthreads();   // X is the next tile's A operand
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 128);
}

#ifndef MPINN_WIDE_PIPELINE
#define MPINN_WIDE_PIPELINE 1   // forward layers: two accumulator blocks, the epilogue of tile i-1 and the activation
#endif                          // jets + bf16 splitting of tile i+1 run under the MMAs of tile i

// Forward of hidden op l, software-pipelined over the CTA's tiles.  Per iteration i: activation jets and bf16x3
// pieces of tile i in registers and the plane loads of tile i+1 issued (the MMAs of tile i-1 are still running),
// wait for those MMAs (they read the A tile), store the pieces, issue the MMAs of tile i into the other accumulator
// block, then drain tile i-1's accumulator to its planes while they run.
template <int D, int ORDER>
__global__ void __launch_bounds__(kWdThreads, 1)
wide_fwd_layer_kernel(const __grid_constant__ DevProg prog, int l, const float* __restrict__ params,
                      const float* __restrict__ packed, const float* __restrict__ src, float* __restrict__ dst,
                      int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr uint32_t kTileBytes = (uint32_t)(J * kWd * kWdPts * 4);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* a_tile = smem;
  uint8_t* w_tile = smem + kWdTile;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  const int p = t & 31, fg = t >> 5;
  const int lq = fg & 3, cq = fg >> 2;
  const DevOp& op = prog.ops[l];
  const int act_below = prog.ops[l - 1].act;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 256);
  copy_image(w_tile, packed + op.tc_off);
  for (int i = t; i < kWdTile / 16; i += kWdThreads) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  float bias[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) bias[i] = lq == 0 ? __ldg(params + op.b_off + 32 * cq + i) : 0.0f;

  auto drain = [&](int64_t tile, uint32_t block) {   // accumulator block -> planes of op l
    if (lq < J) {
      float acc[32];
      tmem_ld32(tmem + block * 128u + ((uint32_t)(lq * 32) << 16) + cq * 32, acc);
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[plane_off<J>(tile, lq, 32 * cq + i, p)] = acc[i] + bias[i];
    }
  };

  float v[J][8];
  if (blockIdx.x < n_tiles) {
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(blockIdx.x, j, 8 * fg + k, p)];
  }
  uint32_t phase = 0, it = 0;
  int64_t prev = -1;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    // ---- activation jets + bf16x3 pieces of this tile, in registers
    uint32_t p1[J][4], p2[J][4], p3[J][4];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      float z0[J], z1[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        z0[j] = v[j][k];
        z1[j] = v[j][k + 1];
      }
      act_jets<D, ORDER>(act_below, z0, prog.lap_c);
      act_jets<D, ORDER>(act_below, z1, prog.lap_c);
#pragma unroll
      for (int j = 0; j < J; ++j) split_bf16x3(z0[j], z1[j], p1[j][k >> 1], p2[j][k >> 1], p3[j][k >> 1]);
    }
    // ---- the next tile's planes start travelling
    const int64_t nxt = tile + gridDim.x;
    if (nxt < n_tiles) {
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(nxt, j, 8 * fg + k, p)];
      if (t == 0 && nxt + gridDim.x < n_tiles) l2_prefetch_bulk_w(src + plane_off<J>(nxt + gridDim.x, 0, 0, 0), kTileBytes);
    }
    // ---- the previous tile's MMAs read the A tile: wait for them, then overwrite it
    if (prev >= 0) {
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after_sync();
    }
#pragma unroll
    for (int j = 0; j < J; ++j) stage_pieces8(a_tile, 32 * j + p, fg, p1[j], p2[j], p3[j]);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      issue_wide_kmajor(tmem + (it & 1u) * 128u, smem_u32(a_tile), smem_u32(w_tile));
      mma_commit(&bar);
    }
    // ---- previous tile's accumulator -> planes, under this tile's MMAs
    if (prev >= 0) drain(prev, (it & 1u) ^ 1u);
    prev = tile;
  }
  if (prev >= 0) {
    mbar_wait(&bar, phase);
    tc_fence_after_sync();
    drain(prev, (it & 1u) ^ 1u);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 256);
}

// Input gradient of hidden op l (gin = Zbar_l, gout = Zbar_{l-1}), reordered so that the expensive part of a tile's
// epilogue -- the activation's reverse and the plane stores -- runs under the NEXT tile's MMAs: the accumulator is
// exchanged through the A tile (channel-major lanes -> all channels of a (point, feature) in one thread) into
// registers first, then the next tile is staged and issued, then the registers are worked off.
template <int D, int ORDER>
__global__ void __launch_bounds__(kWdThreads, 1)
wide_dgrad_layer_kernel(const __grid_constant__ DevProg prog, int l, const float* __restrict__ packed,
                        const float* __restrict__ src, const float* __restrict__ zprev, float* __restrict__ dst,
                        int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr uint32_t kTileBytes = (uint32_t)(J * kWd * kWdPts * 4);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* a_tile = smem;
  uint8_t* w_tile = smem + kWdTile;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  const int p = t & 31, fg = t >> 5;
  const int lq = fg & 3, cq = fg >> 2;
  const DevOp& op = prog.ops[l];
  const int act_below = prog.ops[l - 1].act;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 128);
  copy_image(w_tile, packed + op.tc_off + kWdTile / 4);
  for (int i = t; i < kWdTile / 16; i += kWdThreads) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  float v[J][8];
  // ---- prologue: stage and issue the CTA's first tile
  if (blockIdx.x < n_tiles) {
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(blockIdx.x, j, 8 * fg + k, p)];
#pragma unroll
    for (int j = 0; j < J; ++j) stage_row8(a_tile, 32 * j + p, fg, v[j]);
    if (t == 0) {
      l2_prefetch_bulk_w(zprev + plane_off<J>(blockIdx.x, 0, 0, 0), kTileBytes);
      if (blockIdx.x + gridDim.x < n_tiles) l2_prefetch_bulk_w(src + plane_off<J>(blockIdx.x + gridDim.x, 0, 0, 0), kTileBytes);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (t == 0 && blockIdx.x < n_tiles) {
    issue_wide_kmajor(tmem, smem_u32(a_tile), smem_u32(w_tile));
    mma_commit(&bar);
  }
  uint32_t phase = 0;
  float* X = reinterpret_cast<float*>(a_tile);   // exchange buffer [feature][row]; the A tile is free once the MMAs are done
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t nxt = tile + gridDim.x;
    const bool more = nxt < n_tiles;
    if (t == 0 && more) {
      l2_prefetch_bulk_w(zprev + plane_off<J>(nxt, 0, 0, 0), kTileBytes);
      if (nxt + gridDim.x < n_tiles) l2_prefetch_bulk_w(src + plane_off<J>(nxt + gridDim.x, 0, 0, 0), kTileBytes);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after_sync();
    if (lq < J) {
      float acc[32];
      tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + cq * 32, acc);
#pragma unroll
      for (int i = 0; i < 32; ++i) X[(32 * cq + i) * 128 + 32 * lq + p] = acc[i];
    }
    if (more) {   // the next tile's Zbar planes travel while the exchange completes
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = src[plane_off<J>(nxt, j, 8 * fg + k, p)];
    }
    tc_fence_before_sync();
    __syncthreads();
    float ab[J][8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < J; ++j) ab[j][k] = X[(8 * fg + k) * 128 + 32 * j + p];
    __syncthreads();   // X consumed: the A tile may be overwritten
    if (more) {
#pragma unroll
      for (int j = 0; j < J; ++j) stage_row8(a_tile, 32 * j + p, fg, v[j]);
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[j][k] = zprev[plane_off<J>(tile, j, 8 * fg + k, p)];
    if (more) {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (t == 0) {
        tc_fence_after_sync();
        issue_wide_kmajor(tmem, smem_u32(a_tile), smem_u32(w_tile));
        mma_commit(&bar);
      }
    }
    // ---- activation reverse + plane stores of this tile, under the next tile's MMAs
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z[J], g[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        z[j] = v[j][k];
        g[j] = ab[j][k];
      }
      act_jets_bwd<D, ORDER>(act_below, z, g, prog.lap_c);
#pragma unroll
      for (int j = 0; j < J; ++j) dst[plane_off<J>(tile, j, 8 * fg + k, p)] = g[j];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 128);
}

// weight / bias gradient of hidden op l:  dW_l += Zbar_l^T act_jets(z_{l-1}),  db_l += sum_p Zbar_l[0]
template <int D, int ORDER>
__global__ void __launch_bounds__(kWdThreads, 1)
wide_wgrad_kernel(const __grid_constant__ DevProg prog, int l, const float* __restrict__ zbar,
                  const float* __restrict__ zprev, float* __restrict__ part, int64_t slab_floats, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* z_tile = smem;
  uint8_t* h_tile = smem + kWdTile;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  const int p = t & 31, fg = t >> 5;
  const int lq = fg & 3, cq = fg >> 2;
  const DevOp& op = prog.ops[l];
  const int act_below = prog.ops[l - 1].act;
  float* slab = part + (int64_t)blockIdx.x * slab_floats;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 128);
  // K runs over the rows: rows of unused channel slots must be zero in both tiles
  for (int i = t; i < 2 * kWdTile / 16; i += kWdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;
  int pending = 0;     // MMAs committed and not yet waited for
  int in_group = 0;    // tiles accumulated since the last flush
  float accb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) accb[k] = 0.0f;

  auto flush = [&]() {
    // accumulator row m = out feature (lane), column n = in feature -> slab[w_off + m * 128 + n]
    float acc[32];
    tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + cq * 32, acc);
    float* d = slab + op.w_off + (32 * lq + p) * kWd + 32 * cq;
#pragma unroll
    for (int i = 0; i < 8; ++i) red_add_v4w(d + 4 * i, acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    tc_fence_before_sync();
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    float zb[J][8], zp[J][8];
#if MPINN_WIDE_PREFETCH
    if (t == 0 && tile + gridDim.x < n_tiles) {
      l2_prefetch_bulk_w(zbar + plane_off<J>(tile + gridDim.x, 0, 0, 0), (uint32_t)(J * kWd * kWdPts * 4));
      l2_prefetch_bulk_w(zprev + plane_off<J>(tile + gridDim.x, 0, 0, 0), (uint32_t)(J * kWd * kWdPts * 4));
    }
#endif
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        zb[j][k] = zbar[plane_off<J>(tile, j, 8 * fg + k, p)];
        zp[j][k] = zprev[plane_off<J>(tile, j, 8 * fg + k, p)];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) accb[k] += zb[0][k];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z[J];
#pragma unroll
      for (int j = 0; j < J; ++j) z[j] = zp[j][k];
      act_jets<D, ORDER>(act_below, z, prog.lap_c);
#pragma unroll
      for (int j = 0; j < J; ++j) zp[j][k] = z[j];
    }
    if (pending) {   // the previous tile's MMAs still read the operand tiles
      mbar_wait(&bar, phase);
      phase ^= 1;
      pending = 0;
      tc_fence_after_sync();
      if (in_group == kWdFlushTiles) {
        flush();
        in_group = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      stage_row8(z_tile, 32 * j + p, fg, zb[j]);
      stage_row8(h_tile, 32 * j + p, fg, zp[j]);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      issue_wide_mnmajor(tmem, smem_u32(z_tile), smem_u32(h_tile), in_group > 0 ? 1u : 0u);
      mma_commit(&bar);
    }
    pending = 1;
    ++in_group;
  }
  if (pending) {
    mbar_wait(&bar, phase);
    tc_fence_after_sync();
    flush();
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], off);
    if (p == 0) atomicAdd(slab + op.b_off + 8 * fg + k, accb[k]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 128);
}

// dparams[i] += sum_c part[c][i]
__global__ void fold_partials_wide(const float* __restrict__ part, int n_slabs, int64_t n_params, int64_t slab_floats,
                                   float* __restrict__ dparams, const double* __restrict__ bsum, int bsum_off,
                                   int n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  float s = 0.0f;
  if (!take_bsum(bsum, bsum_off, n_out, i, &s))
    for (int c = 0; c < n_slabs; ++c) s += part[(int64_t)c * slab_floats + i];
  dparams[i] += s;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER>
static cudaError_t fwd_wide_t(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                              float* out, float* stash, int64_t n_pad, int sm_count, cudaStream_t stream) {
  const int n_tiles = (int)((n + kWdPts - 1) / kWdPts);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  auto kern = wide_layer_kernel<D, ORDER, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWdSmem);
  if (e != cudaSuccess) return e;
  const int g_cc = n_tiles < 4 * sm_count ? n_tiles : 4 * sm_count;
  wide_first_fwd<D, ORDER><<<g_cc, kWdThreads, 0, stream>>>(prog, params, x, n, stash + (int64_t)prog.ops[0].stash_plane * n_pad,
                                                           n_tiles);
#if MPINN_WIDE_PIPELINE
  auto kpipe = wide_fwd_layer_kernel<D, ORDER>;
  e = cudaFuncSetAttribute(kpipe, cudaFuncAttributeMaxDynamicSharedMemorySize, kWdSmem);
  if (e != cudaSuccess) return e;
  for (int l = 1; l < prog.n_ops - 1; ++l)
    kpipe<<<grid, kWdThreads, kWdSmem, stream>>>(prog, l, params, packed, stash + (int64_t)prog.ops[l - 1].stash_plane * n_pad,
                                                 stash + (int64_t)prog.ops[l].stash_plane * n_pad, n_tiles);
#else
  for (int l = 1; l < prog.n_ops - 1; ++l)
    kern<<<grid, kWdThreads, kWdSmem, stream>>>(prog, l, params, packed, stash + (int64_t)prog.ops[l - 1].stash_plane * n_pad,
                                                nullptr, stash + (int64_t)prog.ops[l].stash_plane * n_pad, n_tiles);
#endif
  if (prog.n_out == 1)
    wide_last_fwd<D, ORDER, 1><<<g_cc, kWdThreads, 0, stream>>>(
        prog, params, stash + (int64_t)prog.ops[prog.n_ops - 2].stash_plane * n_pad, n, out, n_tiles);
  else
    wide_last_fwd<D, ORDER, kWdMaxOut><<<g_cc, kWdThreads, 0, stream>>>(
        prog, params, stash + (int64_t)prog.ops[prog.n_ops - 2].stash_plane * n_pad, n, out, n_tiles);
  return cudaGetLastError();
}

template <int D, int ORDER>
static cudaError_t bwd_wide_t(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                              const float* dout, float* stash, int64_t n_pad, float* part, int max_slabs,
                              int64_t n_params, float* dparams, int sm_count, cudaStream_t stream) {
  constexpr int J = jet_count(D, ORDER);
  const int n_tiles = (int)((n + kWdPts - 1) / kWdPts);
  int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid > max_slabs - 1) grid = max_slabs - 1;
  if (grid < 1) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(part) & 15) != 0) return cudaErrorMisalignedAddress;
  const int64_t slab = (n_params + 3) & ~int64_t(3);
  cudaError_t e = cudaMemsetAsync(part, 0, (size_t)grid * slab * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  auto kin = wide_layer_kernel<D, ORDER, 1>;
  auto kw = wide_wgrad_kernel<D, ORDER>;
  e = cudaFuncSetAttribute(kin, cudaFuncAttributeMaxDynamicSharedMemorySize, kWdSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, kWdSmem);
  if (e != cudaSuccess) return e;
  auto kdg = wide_dgrad_layer_kernel<D, ORDER>;
  e = cudaFuncSetAttribute(kdg, cudaFuncAttributeMaxDynamicSharedMemorySize, kWdSmem);
  if (e != cudaSuccess) return e;
  float* zb[2] = {stash + (int64_t)prog.total_planes * n_pad, stash + ((int64_t)prog.total_planes + J * kWd) * n_pad};
  const int L = prog.n_ops;
  int cur = 0;
  if (prog.n_out == 1) {   // 64 registers: two CTAs per SM
    const int g_last = n_tiles < 2 * sm_count ? n_tiles : 2 * sm_count;
    wide_last_bwd<D, ORDER, 1><<<g_last, kWdThreads, 0, stream>>>(prog, params, stash + (int64_t)prog.ops[L - 2].stash_plane * n_pad,
                                                                 dout, n, zb[cur], part, slab, grid, n_tiles);
  } else {
    wide_last_bwd<D, ORDER, kWdMaxOut><<<grid, kWdThreads, 0, stream>>>(
        prog, params, stash + (int64_t)prog.ops[L - 2].stash_plane * n_pad, dout, n, zb[cur], part, slab, grid, n_tiles);
  }
  for (int l = L - 2; l >= 1; --l) {
    const float* zprev = stash + (int64_t)prog.ops[l - 1].stash_plane * n_pad;
    kw<<<grid, kWdThreads, kWdSmem, stream>>>(prog, l, zb[cur], zprev, part, slab, n_tiles);
#if MPINN_WIDE_PIPELINE
    kdg<<<grid, kWdThreads, kWdSmem, stream>>>(prog, l, packed, zb[cur], zprev, zb[cur ^ 1], n_tiles);
#else
    kin<<<grid, kWdThreads, kWdSmem, stream>>>(prog, l, params, packed, zb[cur], zprev, zb[cur ^ 1], n_tiles);
#endif
    cur ^= 1;
  }
  const int g_first = n_tiles < 3 * sm_count ? n_tiles : 3 * sm_count;
  wide_first_bwd<D, ORDER><<<g_first, kWdThreads, 0, stream>>>(prog, x, n, zb[cur], part, slab, grid, n_tiles);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fold_partials_wide<<<(unsigned)((n_params + 255) / 256), 256, 0, stream>>>(part, grid, n_params, slab, dparams, prog.bsum,
                                                                             prog.bsum_off, prog.n_out);
  return cudaGetLastError();
}

#define WIDE_CASES(CALL)        \
  if (d == 1 && order == 0) CALL(1, 0); \
  if (d == 1 && order == 1) CALL(1, 1); \
  if (d == 1 && order == 2) CALL(1, 2); \
  if (d == 1 && order == 3) CALL(1, 3); \
  if (d == 2 && order == 0) CALL(2, 0); \
  if (d == 2 && order == 1) CALL(2, 1); \
  if (d == 2 && order == 3) CALL(2, 3); \
  if (d == 3 && order == 0) CALL(3, 0); \
  if (d == 3 && order == 1) CALL(3, 1);

cudaError_t launch_fwd_wide(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                            int d, int order, float* out, float* stash, int64_t n_pad, int sm_count,
                            cudaStream_t stream) {
#define CALL(D_, O_) return fwd_wide_t<D_, O_>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream)
  WIDE_CASES(CALL)
#undef CALL
  return cudaErrorInvalidValue;
}

cudaError_t launch_bwd_wide(const DevProg& prog, const float* params, const float* packed, const float* x, int64_t n,
                            int d, int order, const float* dout, float* stash, int64_t n_pad, float* part,
                            int max_slabs, int64_t n_params, float* dparams, int sm_count, cudaStream_t stream) {
#define CALL(D_, O_) \
  return bwd_wide_t<D_, O_>(prog, params, packed, x, n, dout, stash, n_pad, part, max_slabs, n_params, dparams, sm_count, stream)
  WIDE_CASES(CALL)
#undef CALL
  return cudaErrorInvalidValue;
}

}  // namespace mpinn
