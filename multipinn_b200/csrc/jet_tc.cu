// jet_tc.cu -- tcgen05 (5th-generation tensor core) engine for the square 64-wide hidden layers.
//
// Tile = 128 collocation points (one TMEM lane per point).  For every jet channel j the layer GEMM
//     Z_j[128 x 64] = A_j[128 x 64] * W^T          (A_j = activation jets of the previous layer)
// runs as 3xTF32 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, 24 tcgen05.mma of shape M128 N64 K8) with both
// operands in shared memory (128B-swizzle canonical tiles) and the fp32 accumulator of channel j in
// TMEM columns [64j, 64j+64).  The activation and its derivatives are evaluated by 256 threads
// (2 per point, 32 features each) while staging A_j ("prologue"), sigma'/sigma'' living in registers
// across the channel loop; the MMA for channel j overlaps the staging of channel j+1 (two A buffers,
// mbarrier hand-off via tcgen05.commit).  Pre-activation jets go to the HBM stash (same planes as
// the FFMA engine, so the two engines' forward/reverse kernels are interchangeable) and are read
// back L2-hot by the next layer's prologue.
#include <cuda_bf16.h>

#include "common.cuh"
#include "launch.h"
#include "umma.cuh"

namespace mpinn {

using namespace umma;

constexpr int kTcW = 64;                 // layer width handled by this engine
constexpr int kTcThreads = 256;          // 8 warps: 2 threads per point
constexpr int kTileRows = 128;
constexpr int kRegionBytes = kTileRows * 128;       // [128 rows x 32 fp32]  = 16 KiB
constexpr int kATileBytes = 4 * kRegionBytes;       // hi kb0, hi kb1, lo kb0, lo kb1 = 64 KiB
constexpr int kWRegionBytes = kTcW * 128;           // [64 rows x 32 fp32]   = 8 KiB
constexpr int kWTileBytes = 4 * kWRegionBytes;      // hi kb0, hi kb1, lo kb0, lo kb1 = 32 KiB
// packed block per layer: forward operand (B[n=out][k=in] = W) then reverse operand (B[n=in][k=out] = W^T)
constexpr int kTcBlockFloats = 2 * kWTileBytes / 4;

int64_t tc_block_floats(int width) { return width == kTcW ? kTcBlockFloats : 0; }

// ---------------------------------------------------------------------------------------------
// weight packing: global image of the swizzled shared-memory operand tiles
// ---------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const __grid_constant__ DevProg prog, const float* __restrict__ params,
                               float* __restrict__ packed) {
  const DevOp& op = prog.ops[blockIdx.y];
  if (op.kind != MPINN_OP_LINEAR || op.tc_off < 0) return;
  const float* W = params + op.w_off;  // [out][in]
  char* dst = reinterpret_cast<char*>(packed + op.tc_off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTcW * kTcW; i += gridDim.x * blockDim.x) {
    const int o = i / kTcW, k = i % kTcW;
    const float w = W[o * kTcW + k];
    float hi, lo;
    split_tf32(w, hi, lo);
    // forward operand: row n = o, column k = in index
    {
      const uint32_t off = (uint32_t)(k / 32) * kWRegionBytes + sw128_offset(o, k % 32);
      *reinterpret_cast<float*>(dst + off) = hi;
      *reinterpret_cast<float*>(dst + 2 * kWRegionBytes + off) = lo;
    }
    // reverse operand: row n = in index, column k = o
    {
      const uint32_t off = (uint32_t)(o / 32) * kWRegionBytes + sw128_offset(k, o % 32);
      *reinterpret_cast<float*>(dst + kWTileBytes + off) = hi;
      *reinterpret_cast<float*>(dst + kWTileBytes + 2 * kWRegionBytes + off) = lo;
    }
  }
}

cudaError_t launch_pack_tc(const DevProg& prog, const float* params, float* packed, cudaStream_t stream) {
  bool any = false;
  for (int l = 0; l < prog.n_ops; ++l) any |= prog.ops[l].tc_off >= 0;
  if (!any) return cudaSuccess;
  dim3 grid(4, prog.n_ops);
  pack_tc_kernel<<<grid, 256, 0, stream>>>(prog, params, packed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// MMA issue helpers (called by one thread)
// ---------------------------------------------------------------------------------------------
// Z[128 x 64] (+)= A[128 x 64] * B[64 x 64]^T, 3xTF32, A/B K-major tiles (hi kb0, hi kb1, lo kb0, lo kb1)
__device__ __forceinline__ void issue_gemm_kmajor(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
  uint32_t acc = 0;
  // the two small correction products go first: the tensor core truncates when it adds into the
  // accumulator, so the large hi*hi partial sums should be the last (and fewest) things added
#pragma unroll
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_base = a_tile + (prod == 0 ? 2 * kRegionBytes : 0);    // A_lo * W_hi first
    const uint32_t b_base = b_tile + (prod == 1 ? 2 * kWRegionBytes : 0);   // then A_hi * W_lo, then hi*hi
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t a_addr = a_base + (ks >> 2) * kRegionBytes + (ks & 3) * 32;
      const uint32_t b_addr = b_base + (ks >> 2) * kWRegionBytes + (ks & 3) * 32;
      mma_tf32_ss(d_tmem, make_desc_sw128(a_addr, 16, 1024), make_desc_sw128(b_addr, 16, 1024), idesc, acc);
      acc = 1;
    }
  }
}

// dW-shaped product: D[128 x 64] (+)= [Zhi | Zlo]^T (features x points) * (Hhi + Hlo) (points x features)
// Z and H tiles are [128 points x 64 features] hi/lo tiles viewed MN-major (K = points).
// rows 0..63 of D: Zhi^T H, rows 64..127: Zlo^T H  (the caller adds the halves)
__device__ __forceinline__ void issue_gemm_mnmajor(uint32_t d_tmem, uint32_t z_tile, uint32_t h_tile,
                                                   uint32_t accumulate_first) {
  constexpr uint32_t idesc = make_idesc_tf32(128, 64, 1, 1);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int part = 0; part < 2; ++part) {                     // H_hi then H_lo
    const uint32_t h_base = h_tile + part * 2 * kRegionBytes;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {                        // 8 points per MMA
      const uint32_t a_addr = z_tile + ks * 1024;            // 4 MN groups (hi kb0, hi kb1, lo kb0, lo kb1)
      const uint32_t b_addr = h_base + ks * 1024;            // 2 MN groups (kb0, kb1)
      mma_tf32_ss(d_tmem, make_desc_sw128(a_addr, kRegionBytes, 1024),
                  make_desc_sw128(b_addr, kRegionBytes, 1024), idesc, acc);
      acc = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BF16x3 operands (reverse pass).  a = a1 + a2 + a3 with three round-to-nearest bf16 pieces (24
// mantissa bits); a product needs the six piece products whose weight is >= 2^-16.
// A [rows x 64 features] piece is ONE 128B-swizzle region (row = 64 bf16 = 128 bytes), so the same
// bytes are a K-major operand (K = features) and an MN-major operand (K = rows).
// ---------------------------------------------------------------------------------------------
constexpr int kBfRegion = kTileRows * 128;    // one piece of a [128 x 64] tile: 16 KiB
constexpr int kBfTile = 4 * kBfRegion;        // piece1, piece2, piece3, ZERO region
constexpr int kBfWRegion = kTcW * 128;        // one piece of a [64 x 64] weight operand: 8 KiB
constexpr int kBfWTile = 3 * kBfWRegion;

__device__ __forceinline__ void split_bf16x3(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  // two values at once: each piece packs (a, b) as a bf16x2 word, a in the low half
  const __nv_bfloat162 h1 = __floats2bfloat162_rn(a, b);
  const float ra = a - __low2float(h1), rb = b - __high2float(h1);
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(ra, rb);
  const float sa = ra - __low2float(h2), sb = rb - __high2float(h2);
  const __nv_bfloat162 h3 = __floats2bfloat162_rn(sa, sb);
  p1 = *reinterpret_cast<const uint32_t*>(&h1);
  p2 = *reinterpret_cast<const uint32_t*>(&h2);
  p3 = *reinterpret_cast<const uint32_t*>(&h3);
}

// byte offset of bf16 element (row, col<64) inside a swizzled region
__host__ __device__ __forceinline__ uint32_t sw128_offset_bf16(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[128 x 64] = A[128 x 64] * B[64 x 64]^T, both K-major bf16x3 tiles; smallest products first
__device__ __forceinline__ void issue_bf16x3_kmajor(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  // (a piece, b piece): weights 2^-16 first, then 2^-8, then the leading product
  constexpr int pa[6] = {2, 1, 0, 1, 0, 0};
  constexpr int pb[6] = {0, 1, 2, 0, 1, 0};
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {   // K = 16 bf16 = 32 bytes per MMA
      const uint32_t a_addr = a_tile + pa[q] * kBfRegion + ks * 32;
      const uint32_t b_addr = b_tile + pb[q] * kBfWRegion + ks * 32;
      mma_bf16_ss(d_tmem, make_desc_sw128(a_addr, 16, 1024), make_desc_sw128(b_addr, 16, 1024), idesc, acc);
      acc = 1;
    }
  }
}

// D[128 x 64] (+)= stacked(Z)^T * H over the 128 rows (points) of the tiles, MN-major views.
//   rows 0..63 : Z1^T (H1 + H2 + H3) + Z3^T H1        rows 64..127 : Z2^T (H1 + H2) (+ Z3^T H3, negligible)
// Z tile regions: [Z1][Z2][Z3][ZERO]; H tile regions: [H1][H2][H3].  The caller adds the two halves.
__device__ __forceinline__ void issue_bf16x3_mnmajor(uint32_t d_tmem, uint32_t z_tile, uint32_t h_tile,
                                                     uint32_t accumulate_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
  // (first Z region, region stride between the two M halves, H piece)
  constexpr int zr[4] = {2, 0, 0, 0};
  constexpr int zs[4] = {1, 2, 1, 1};
  constexpr int hp[4] = {0, 2, 1, 0};
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {   // K = 16 points = two 8-row swizzle atoms (SBO = 1024)
      const uint32_t a_addr = z_tile + zr[q] * kBfRegion + ks * 2048;
      const uint32_t b_addr = h_tile + hp[q] * kBfRegion + ks * 2048;
      mma_bf16_ss(d_tmem, make_desc_sw128(a_addr, zs[q] * kBfRegion, 1024),
                  make_desc_sw128(b_addr, kBfRegion, 1024), idesc, acc);
      acc = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// self-test kernels (diagnostics for the descriptor/layout conventions)
// ---------------------------------------------------------------------------------------------
// mode 0: C[128][64] = A[128][64] * B[64][64]^T      (K-major path)
// mode 1: C[128][64] = [Zhi|Zlo]^T H, Z=A[128][64], H=B2[128][64]  (MN-major path; host adds the halves)
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, const float* __restrict__ A,
                                                            const float* __restrict__ B, float* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = smem;                    // 64 KiB
  uint8_t* b_tile = smem + kATileBytes;      // up to 64 KiB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 64);
  if (mode >= 2) {
    // bf16x3 staging: A [128][64] -> regions [A1][A2][A3][ZERO]; B [64 or 128][64] -> [B1][B2][B3]
    for (int i = t; i < 128 * 32; i += 128) {
      const int r = i / 32, c = (i % 32) * 2;
      uint32_t p1, p2, p3;
      split_bf16x3(A[r * 64 + c], A[r * 64 + c + 1], p1, p2, p3);
      const uint32_t off = sw128_offset_bf16(r, c);
      *reinterpret_cast<uint32_t*>(a_tile + off) = p1;
      *reinterpret_cast<uint32_t*>(a_tile + kBfRegion + off) = p2;
      *reinterpret_cast<uint32_t*>(a_tile + 2 * kBfRegion + off) = p3;
      *reinterpret_cast<uint32_t*>(a_tile + 3 * kBfRegion + off) = 0u;
    }
    const int brows2 = mode == 2 ? 64 : 128;
    const int breg2 = brows2 * 128;
    for (int i = t; i < brows2 * 32; i += 128) {
      const int r = i / 32, c = (i % 32) * 2;
      uint32_t p1, p2, p3;
      split_bf16x3(B[r * 64 + c], B[r * 64 + c + 1], p1, p2, p3);
      const uint32_t off = sw128_offset_bf16(r, c);
      *reinterpret_cast<uint32_t*>(b_tile + off) = p1;
      *reinterpret_cast<uint32_t*>(b_tile + breg2 + off) = p2;
      *reinterpret_cast<uint32_t*>(b_tile + 2 * breg2 + off) = p3;
    }
  }
  // stage A (always [128][64] -> hi/lo regions)
  for (int i = (mode >= 2 ? 128 * 64 : t); i < 128 * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    float hi, lo;
    split_tf32(A[i], hi, lo);
    const uint32_t off = (c / 32) * kRegionBytes + sw128_offset(r, c % 32);
    *reinterpret_cast<float*>(a_tile + off) = hi;
    *reinterpret_cast<float*>(a_tile + 2 * kRegionBytes + off) = lo;
  }
  const int brows = mode == 0 ? 64 : 128;
  const int bregion = brows * 128;
  for (int i = (mode >= 2 ? brows * 64 : t); i < brows * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    float hi, lo;
    split_tf32(B[i], hi, lo);
    const uint32_t off = (c / 32) * bregion + sw128_offset(r, c % 32);
    *reinterpret_cast<float*>(b_tile + off) = hi;
    *reinterpret_cast<float*>(b_tile + 2 * bregion + off) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    if (mode == 0) issue_gemm_kmajor(tmem, smem_u32(a_tile), smem_u32(b_tile));
    else if (mode == 1) issue_gemm_mnmajor(tmem, smem_u32(a_tile), smem_u32(b_tile), 0);
    else if (mode == 2) issue_bf16x3_kmajor(tmem, smem_u32(a_tile), smem_u32(b_tile));
    else issue_bf16x3_mnmajor(tmem, smem_u32(a_tile), smem_u32(b_tile), 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int warp = t / 32;
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) C[t * 64 + half * 32 + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 64);
}

cudaError_t launch_umma_selftest(int mode, const float* A, const float* B, float* C, cudaStream_t stream) {
  const int smem = 2 * kATileBytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_selftest_kernel<<<1, 128, smem, stream>>>(mode, A, B, C);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct TcSmem {
  static constexpr int kA0 = 0;
  static constexpr int kA1 = kATileBytes;
  static constexpr int kW = 2 * kATileBytes;
  static constexpr int kBias = kW + kWTileBytes;            // 64 floats
  static constexpr int kRed = kBias + 256;                  // 128 points x 8 floats scratch
  static constexpr int kTotal = kRed + 128 * 8 * 4 + 1024;  // + alignment slack
};

// which TMEM quarter / point / feature half a thread owns
struct TcThread {
  int p, fh, warp;
  uint32_t lane_base;  // TMEM lane field for tcgen05.ld/st issued by this warp
};

template <int D, int ORDER>
__global__ void __launch_bounds__(kTcThreads, 1)
jet_fwd_tc(const __grid_constant__ DevProg prog, const float* __restrict__ params,
           const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
           float* __restrict__ out, float* __restrict__ stash, int64_t n_pad, int n_tiles) {
  constexpr int J = 1 + ORDER * D;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_buf[2] = {smem + TcSmem::kA0, smem + TcSmem::kA1};
  uint8_t* w_buf = smem + TcSmem::kW;
  float* bias_s = reinterpret_cast<float*>(smem + TcSmem::kBias);
  float* red_s = reinterpret_cast<float*>(smem + TcSmem::kRed);
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int t = threadIdx.x;
  TcThread th;
  th.p = t % 128;
  th.fh = t / 128;
  th.warp = t / 32;
  th.lane_base = (uint32_t)((th.warp % 4) * 32) << 16;

  if (t == 0) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    fence_mbar_init();
  }
  if (th.warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t bar_phase[2] = {0, 0};   // parity of the next completion to wait for, per A buffer
  int pending[2] = {0, 0};          // an MMA group is in flight on that buffer

  const int o = prog.n_out;
  const int f0 = th.fh * 32;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * 128 + th.p;
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;

    // ---- op 0: LINEAR(d -> 64) on the FMA pipe: z = W x + b, z_i = W[:, i], z_ii = 0  -> stash
    {
      const DevOp& op = prog.ops[0];
      const float* W = params + op.w_off;
      const float* b = params + op.b_off;
#pragma unroll 4
      for (int ff = 0; ff < 32; ++ff) {
        const int f = f0 + ff;
        float z0 = __ldg(b + f);
        float w[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          w[i] = __ldg(W + f * D + i);
          z0 = fmaf(w[i], xv[i], z0);
        }
        float* sp = stash + (int64_t)(op.stash_plane + f) * n_pad + pt;
        sp[0] = z0;
        if (ORDER >= 1) {
#pragma unroll
          for (int i = 0; i < D; ++i) sp[(int64_t)(1 + i) * kTcW * n_pad] = w[i];
        }
        if (ORDER >= 2) {
#pragma unroll
          for (int i = 0; i < D; ++i) sp[(int64_t)(1 + D + i) * kTcW * n_pad] = 0.0f;
        }
      }
    }
    __syncthreads();  // stash planes of this tile are written by both feature halves

    // ---- hidden GEMM layers
    for (int l = 1; l < prog.n_ops - 1; ++l) {
      const DevOp& op = prog.ops[l];
      const DevOp& pv = prog.ops[l - 1];
      // make sure no MMA still reads the weight tile, then load this layer's operand + bias
      for (int b = 0; b < 2; ++b) {
        if (pending[b]) {
          mbar_wait(&mma_bar[b], bar_phase[b]);
          bar_phase[b] ^= 1;
          pending[b] = 0;
        }
      }
      {
        const float4* src = reinterpret_cast<const float4*>(packed + op.tc_off);
        float4* dst = reinterpret_cast<float4*>(w_buf);
        for (int i = t; i < kWTileBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
        if (t < kTcW) bias_s[t] = __ldg(params + op.b_off + t);
      }
      // sigma'(z0), sigma''(z0) of the previous layer for this thread's 32 features
      float s1[32], s2[32];
      const float* zprev = stash + (int64_t)pv.stash_plane * n_pad + pt;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int buf = j & 1;
        if (pending[buf]) {
          mbar_wait(&mma_bar[buf], bar_phase[buf]);
          bar_phase[buf] ^= 1;
          pending[buf] = 0;
        }
        uint8_t* at = a_buf[buf];
        // prologue: activation jet channel j of the previous layer for (point, 32 features)
#pragma unroll
        for (int q = 0; q < 8; ++q) {     // 8 chunks of 4 features = one 16-byte store each (hi and lo)
          float hi4[4], lo4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ff = q * 4 + e;
            const int f = f0 + ff;
            float a;
            if (j == 0) {
              const float z0 = __ldcg(zprev + (int64_t)f * n_pad);
              float c0, c1, c2, c3;
              act_coeffs<false>(pv.act, z0, c0, c1, c2, c3);
              s1[ff] = c1;
              s2[ff] = c2;
              a = c0;
            } else if (j <= D) {
              const float zi = __ldcg(zprev + (int64_t)(j * kTcW + f) * n_pad);
              a = s1[ff] * zi;
            } else {
              const float zi = __ldcg(zprev + (int64_t)((j - D) * kTcW + f) * n_pad);
              const float zii = __ldcg(zprev + (int64_t)(j * kTcW + f) * n_pad);
              a = fmaf(s2[ff] * zi, zi, s1[ff] * zii);
            }
            split_tf32(a, hi4[e], lo4[e]);
          }
          const int col = q * 4;  // column inside the 32-feature K block owned by this thread half
          const uint32_t off = (uint32_t)th.fh * kRegionBytes + sw128_offset(th.p, col);
          *reinterpret_cast<float4*>(at + off) = make_float4(hi4[0], hi4[1], hi4[2], hi4[3]);
          *reinterpret_cast<float4*>(at + 2 * kRegionBytes + off) = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (t == 0) {
          tc_fence_after_sync();
          issue_gemm_kmajor(tmem + j * kTcW, smem_u32(at), smem_u32(w_buf));
          mma_commit(&mma_bar[buf]);
        }
        pending[buf] = 1;
      }
      // wait for all channels
      for (int b = 0; b < 2; ++b) {
        if (pending[b]) {
          mbar_wait(&mma_bar[b], bar_phase[b]);
          bar_phase[b] ^= 1;
          pending[b] = 0;
        }
      }
      tc_fence_after_sync();
      // epilogue: z = D + b  -> stash (the next layer's prologue, or the output layer, reads it back)
      float* zout = stash + (int64_t)op.stash_plane * n_pad + pt;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        float v[32];
        tmem_ld32(tmem + th.lane_base + j * kTcW + f0, v);
#pragma unroll
        for (int ff = 0; ff < 32; ++ff) {
          float z = v[ff];
          if (j == 0) z += bias_s[f0 + ff];
          zout[(int64_t)(j * kTcW + f0 + ff) * n_pad] = z;
        }
      }
      tc_fence_before_sync();
      __syncthreads();
    }

    // ---- output op: LINEAR(64 -> o) on the FMA pipe, from the last hidden layer's stash
    {
      const DevOp& op = prog.ops[prog.n_ops - 1];
      const DevOp& pv = prog.ops[prog.n_ops - 2];
      const float* W = params + op.w_off;
      const float* zprev = stash + (int64_t)pv.stash_plane * n_pad + pt;
      float acc[J][4];
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.0f;
#pragma unroll 2
      for (int ff = 0; ff < 32; ++ff) {
        const int f = f0 + ff;
        float z[J];
#pragma unroll
        for (int j = 0; j < J; ++j) z[j] = __ldcg(zprev + (int64_t)(j * kTcW + f) * n_pad);
        act_jets<D, ORDER>(pv.act, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < o) {
            const float w = __ldg(W + k * kTcW + f);
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j][k] = fmaf(w, z[j], acc[j][k]);
          }
        }
      }
      // combine the two feature halves of each point through shared memory
      for (int base = 0; base < J * o; base += 8) {
        if (th.fh == 1) {
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = j * o + k - base;
              if (k < o && idx >= 0 && idx < 8) red_s[th.p * 8 + idx] = acc[j][k];
            }
        }
        __syncthreads();
        if (th.fh == 0 && valid) {
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = j * o + k - base;
              if (k < o && idx >= 0 && idx < 8) {
                float v = acc[j][k] + red_s[th.p * 8 + idx];
                if (j == 0) v += __ldg(params + op.b_off + k);
                out[((int64_t)j * n + pt) * o + k] = v;
              }
            }
        }
        __syncthreads();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (th.warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_forward_supported(const DevProg& prog, int d, int order) {
  const int J = 1 + order * d;
  if (J * kTcW > 448) return false;
  if (prog.n_ops < 3) return false;
  if (prog.n_out > 4) return false;
  const DevOp& first = prog.ops[0];
  if (first.kind != MPINN_OP_LINEAR || first.in_dim != d || first.out_dim != kTcW || first.flags) return false;
  for (int l = 1; l < prog.n_ops - 1; ++l) {
    const DevOp& op = prog.ops[l];
    if (op.kind != MPINN_OP_LINEAR || op.in_dim != kTcW || op.out_dim != kTcW || op.flags || op.tc_off < 0)
      return false;
  }
  const DevOp& last = prog.ops[prog.n_ops - 1];
  return last.kind == MPINN_OP_LINEAR && last.in_dim == kTcW && last.act == MPINN_ACT_NONE && !last.flags;
}

bool tc_backward_supported(const DevProg&, int, int) { return false; }

template <int D, int ORDER>
static cudaError_t launch_fwd_tc_t(const DevProg& prog, const float* params, const float* packed,
                                   const float* x, int64_t n, float* out, float* stash, int64_t n_pad,
                                   int sm_count, cudaStream_t stream) {
  auto kern = jet_fwd_tc<D, ORDER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::kTotal);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n + 127) / 128);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  kern<<<grid, kTcThreads, TcSmem::kTotal, stream>>>(prog, params, packed, x, n, out, stash, n_pad, n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_fwd_tc(const DevProg& prog, const float* params, const float* packed, const float* x,
                          int64_t n, int d, int order, float* out, float* stash, int64_t n_pad,
                          int sm_count, cudaStream_t stream) {
  if (stash == nullptr) return cudaErrorInvalidValue;  // this engine always runs through the stash
#define TC_FWD(D_, O_) \
  if (d == D_ && order == O_) \
    return launch_fwd_tc_t<D_, O_>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream)
  TC_FWD(1, 0); TC_FWD(1, 1); TC_FWD(1, 2);
  TC_FWD(2, 0); TC_FWD(2, 1); TC_FWD(2, 2);
  TC_FWD(3, 0); TC_FWD(3, 1); TC_FWD(3, 2);
#undef TC_FWD
  return cudaErrorInvalidValue;
}

cudaError_t launch_bwd_tc(const DevProg& prog, const float* params, const float* packed, const float* x,
                          int64_t n, int d, int order, const float* dout, const float* stash, int64_t n_pad,
                          float* part, int max_slabs, int64_t n_params, float* dparams, float* dx,
                          int sm_count, cudaStream_t stream) {
  (void)sm_count;
  return cudaErrorNotSupported;
}

}  // namespace mpinn
