// jet_tc.cu -- tcgen05 (5th-generation tensor core) engine for the square 64-wide hidden layers.
//
// Tile = 128 collocation points (one TMEM lane per point).  For every jet channel j the layer GEMM
//     Z_j[128 x 64] = A_j[128 x 64] * W^T          (A_j = activation jets of the previous layer)
// runs as 3xTF32 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, 24 tcgen05.mma of shape M128 N64 K8) with both
// operands in shared memory (128B-swizzle canonical tiles) and the fp32 accumulator of channel j in
// TMEM columns [64j, 64j+64).  The activation and its derivatives are evaluated by 256 threads
// (2 per point, 32 features each) while staging A_j ("prologue"), sigma'/sigma'' living in registers
// across the channel loop; the MMA for channel j overlaps the staging of channel j+1 (two A buffers,
// mbarrier hand-off via tcgen05.commit).  TMEM is managed as eight rotating 64-column blocks: while layer
// l's accumulators are consumed channel by channel (bias, activation jets, hi/lo split, staging), layer
// l+1's accumulate into blocks already freed, so the whole forward chain stays on chip; pre-activation
// jets are written once to the tile-blocked HBM stash for the reverse pass.
#include <cuda_bf16.h>

#include "common.cuh"
#include "launch.h"
#include "umma.cuh"

namespace mpinn {

using namespace umma;

constexpr int kTcW = 64;                 // layer width handled by this engine
constexpr int kTcThreads = 512;          // 16 warps: 4 threads per point, 16 features each
constexpr int kFpt = kTcW / (kTcThreads / 128);  // features per thread
constexpr int kTcBlock = kTcThreads;
constexpr int kTileRows = 128;
constexpr int kRegionBytes = kTileRows * 128;       // [128 rows x 32 fp32]  = 16 KiB
constexpr int kATileBytes = 4 * kRegionBytes;       // hi kb0, hi kb1, lo kb0, lo kb1 = 64 KiB
constexpr int kWRegionBytes = kTcW * 128;           // [64 rows x 32 fp32]   = 8 KiB
constexpr int kWTileBytes = 4 * kWRegionBytes;      // hi kb0, hi kb1, lo kb0, lo kb1 = 32 KiB
// packed block per layer: forward operand (B[n=out][k=in] = W) then reverse operand (B[n=in][k=out] = W^T)
constexpr int kTcBlockFloats = 2 * kWTileBytes / 4;

int64_t tc_block_floats(int width) { return width == kTcW ? kTcBlockFloats : 0; }

constexpr int kBfRegion = kTileRows * 128;    // one bf16 piece of a [128 x 64] tile: 16 KiB
constexpr int kBfTile = 4 * kBfRegion;        // piece1, piece2, piece3, ZERO region
constexpr int kBfWRegion = kTcW * 128;        // one bf16 piece of a [64 x 64] weight operand: 8 KiB
constexpr int kBfWTile = 3 * kBfWRegion;

// byte offset of bf16 element (row, col<64) inside a swizzled region
__host__ __device__ __forceinline__ uint32_t sw128_offset_bf16(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

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
    // reverse operand (BF16x3, K-major): row n = in index, column k = o
    {
      const __nv_bfloat16 b1 = __float2bfloat16_rn(w);
      const float r1 = w - __bfloat162float(b1);
      const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
      const uint32_t off = sw128_offset_bf16(k, o);
      *reinterpret_cast<__nv_bfloat16*>(dst + kWTileBytes + off) = b1;
      *reinterpret_cast<__nv_bfloat16*>(dst + kWTileBytes + kBfWRegion + off) = b2;
      *reinterpret_cast<__nv_bfloat16*>(dst + kWTileBytes + 2 * kBfWRegion + off) = b3;
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
template <bool ELECT = false>
__device__ __forceinline__ void issue_gemm_kmajor(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
  // descriptors differ only in their 14-bit start-address field (units of 16 bytes): build them once
  // and add compile-time constants, so the single issuing thread spends ~3 instructions per MMA
  const uint64_t a0 = make_desc_sw128(a_tile, 16, 1024);
  const uint64_t b0 = make_desc_sw128(b_tile, 16, 1024);
  uint32_t acc = 0;
  // the two small correction products go first: the tensor core truncates when it adds into the
  // accumulator, so the large hi*hi partial sums should be the last (and fewest) things added
#pragma unroll
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_base = (prod == 0 ? 2 * kRegionBytes : 0);    // A_lo * W_hi first
    const uint32_t b_base = (prod == 1 ? 2 * kWRegionBytes : 0);   // then A_hi * W_lo, then hi*hi
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t a_off = a_base + (ks >> 2) * kRegionBytes + (ks & 3) * 32;
      const uint32_t b_off = b_base + (ks >> 2) * kWRegionBytes + (ks & 3) * 32;
      if (ELECT) mma_tf32_ss_elect(d_tmem, a0 + (a_off >> 4), b0 + (b_off >> 4), idesc, acc);
      else mma_tf32_ss(d_tmem, a0 + (a_off >> 4), b0 + (b_off >> 4), idesc, acc);
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
template <bool ELECT = false>
__device__ __forceinline__ void issue_bf16x3_kmajor(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  // (a piece, b piece): weights 2^-16 first, then 2^-8, then the leading product
  constexpr int pa[6] = {2, 1, 0, 1, 0, 0};
  constexpr int pb[6] = {0, 1, 2, 0, 1, 0};
  const uint64_t a0 = make_desc_sw128(a_tile, 16, 1024);
  const uint64_t b0 = make_desc_sw128(b_tile, 16, 1024);
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {   // K = 16 bf16 = 32 bytes per MMA
      const uint64_t ad = a0 + ((pa[q] * kBfRegion + ks * 32) >> 4), bd = b0 + ((pb[q] * kBfWRegion + ks * 32) >> 4);
      if (ELECT) mma_f16_ss_elect(d_tmem, ad, bd, idesc, acc);
      else mma_bf16_ss(d_tmem, ad, bd, idesc, acc);
      acc = 1;
    }
  }
}

// D[128 x 64] (+)= stacked(Z)^T * H over the 128 rows (points) of the tiles, MN-major views.
//   rows 0..63 : Z1^T (H1 + H2 + H3) + Z3^T H1        rows 64..127 : Z2^T (H1 + H2) (+ Z3^T H3, negligible)
// Z tile regions: [Z1][Z2][Z3][ZERO]; H tile regions: [H1][H2][H3].  The caller adds the two halves.
template <bool ELECT = false>
__device__ __forceinline__ void issue_bf16x3_mnmajor(uint32_t d_tmem, uint32_t z_tile, uint32_t h_tile,
                                                     uint32_t accumulate_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
  // (first Z region, region stride between the two M halves, H piece)
  constexpr int zr[4] = {2, 0, 0, 0};
  constexpr int zs[4] = {1, 2, 1, 1};
  constexpr int hp[4] = {0, 2, 1, 0};
  const uint64_t z1 = make_desc_sw128(z_tile, 1 * kBfRegion, 1024);   // M halves one region apart
  const uint64_t z2 = make_desc_sw128(z_tile, 2 * kBfRegion, 1024);   // M halves two regions apart
  const uint64_t h0 = make_desc_sw128(h_tile, kBfRegion, 1024);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {   // K = 16 points = two 8-row swizzle atoms (SBO = 1024)
      const uint64_t zd = (zs[q] == 1 ? z1 : z2) + ((zr[q] * kBfRegion + ks * 2048) >> 4);
      if (ELECT) mma_f16_ss_elect(d_tmem, zd, h0 + ((hp[q] * kBfRegion + ks * 2048) >> 4), idesc, acc);
      else mma_bf16_ss(d_tmem, zd, h0 + ((hp[q] * kBfRegion + ks * 2048) >> 4), idesc, acc);
      acc = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// self-test kernels (diagnostics for the descriptor/layout conventions)
// ---------------------------------------------------------------------------------------------
// mode 0: C[128][64] = A[128][64] * B[64][64]^T               3xTF32, K-major operands
// mode 2: same product, BF16x3 K-major operands
// mode 3: C[128][64] = stacked(A)^T B, A,B [128][64]          BF16x3 MN-major operands (host adds halves)
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
// optional cycle accounting of the reverse kernel (build with -DMPINN_TC_TIMING=1): CTA 0, threads 0
// and 224 accumulate clock64() deltas per code segment; read back with mpinn_debug_counters().
// ---------------------------------------------------------------------------------------------
#ifndef MPINN_TC_TIMING
#define MPINN_TC_TIMING 0
#endif
__device__ long long g_tc_dbg[2][16];
#if MPINN_TC_TIMING
#define TC_TIME(slot)                                                   \
  do {                                                                  \
    if (blockIdx.x == 0 && (t == 0 || t == 224)) {                      \
      const long long now_ = clock64();                                 \
      g_tc_dbg[t == 0 ? 0 : 1][slot] += now_ - dbg_last;                \
      dbg_last = now_;                                                  \
    }                                                                   \
  } while (0)
#else
#define TC_TIME(slot) do {} while (0)
#endif

cudaError_t read_tc_debug(long long* out, int n) {
  long long h[32];
  cudaError_t e = cudaMemcpyFromSymbol(h, g_tc_dbg, sizeof(h));
  if (e != cudaSuccess) return e;
  for (int i = 0; i < n && i < 32; ++i) out[i] = h[i];
  long long z[32] = {0};
  return cudaMemcpyToSymbol(g_tc_dbg, z, sizeof(z));
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct TcSmem {
  static constexpr int kA0 = 0;
  static constexpr int kA1 = kATileBytes;
  static constexpr int kW = 2 * kATileBytes;
  static constexpr int kBias = kW + kWTileBytes;            // 64 floats
  static constexpr int kRed = kBias + 256;                  // [3 quarters][128 points][8] floats scratch
  static constexpr int kTotal = kRed + 3 * 128 * 8 * 4 + 1024;  // + alignment slack
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
  float* red_s = reinterpret_cast<float*>(smem + TcSmem::kRed);   // [3][128][8]
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int t = threadIdx.x;
  const int p = t % 128, fq = t / 128, warp = t / 32;
  const uint32_t lane_base = (uint32_t)((warp % 4) * 32) << 16;
  const int f0 = fq * kFpt;                                   // first feature of this thread
  const uint32_t tile_kb = (uint32_t)(f0 / 32) * kRegionBytes;  // K block of this thread's features
  const int col0 = f0 % 32;

  if (t == 0) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t bar_phase0 = 0, bar_phase1 = 0;   // parity of the next completion to wait for, per A buffer
  int pending0 = 0, pending1 = 0;            // an MMA group is in flight on that buffer
  const int o = prog.n_out;
  const int n_ops = prog.n_ops;

  // TMEM is managed as eight 64-column blocks.  blk[j] = block holding channel j of the layer being
  // consumed; channel j of the layer being produced goes to a free block.  Every thread tracks the
  // same (deterministic) bookkeeping.
  int blk[J], nblk[J];
  uint32_t free_mask = 0xFFu;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * 128 + p;
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;
    float* tstash = stash + (int64_t)tile * prog.total_planes * 128 + p;   // [plane][128 points]
    // jets travelling along a skip connection: this CTA's [channel][feature][point] scratch (L2 resident)
    float* skip_s = stash + (int64_t)prog.total_planes * n_pad + (int64_t)blockIdx.x * (J * kTcW * 128) + p;

    // ---- GEMM layers L = 1 .. n_ops-2; the source layer S = L-1 is consumed channel by channel:
    //      its pre-activation jets come from registers (S = 0: z = W x + b on the FMA pipe) or from
    //      TMEM (S >= 1), go to the stash for the reverse pass, and become the A operand of layer L.
    for (int L = 1; L < n_ops - 1; ++L) {
      const DevOp& op = prog.ops[L];
      const DevOp& sv = prog.ops[L - 1];
      const int S = L - 1;
      const bool add_skip = (S >= 1) && (sv.flags & MPINN_FLAG_ADD_SKIP);   // z^S includes the block input
      const bool save_skip = (op.flags & MPINN_FLAG_SAVE_SKIP);            // a^S is a block input
      {
        const float4* src = reinterpret_cast<const float4*>(packed + op.tc_off);
        float4* dst = reinterpret_cast<float4*>(w_buf);
        for (int i = t; i < kWTileBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
        if (t < kTcW) bias_s[t] = __ldg(params + sv.b_off + t);
      }
      __syncthreads();
      float s1[kFpt], s2[kFpt], zi_keep[kFpt];
      float* zst = tstash + sv.stash_plane * 128;
#pragma unroll 1
      for (int idx = 0; idx < J; ++idx) {
        // production order: value, then (first, second) derivative per direction
        int j;
        if (idx == 0) j = 0;
        else if (ORDER == 2) j = ((idx - 1) & 1) ? 1 + D + ((idx - 1) >> 1) : 1 + ((idx - 1) >> 1);
        else j = idx;
        const bool is_first = (j >= 1 && j <= D);
        const bool is_second = (j > D);
        const int dir = is_first ? j - 1 : (is_second ? j - 1 - D : 0);
        const int buf = idx & 1;
        uint8_t* at = a_buf[buf];
#pragma unroll
        for (int c = 0; c < kFpt / 8; ++c) {
          float z[8];
          if (S == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int f = f0 + c * 8 + e;
              if (j == 0) {
                float v = bias_s[f];
#pragma unroll
                for (int i = 0; i < D; ++i) v = fmaf(__ldg(params + sv.w_off + f * D + i), xv[i], v);
                z[e] = v;
              } else if (is_first) {
                z[e] = __ldg(params + sv.w_off + f * D + dir);
              } else {
                z[e] = 0.0f;
              }
            }
          } else {
            tmem_ld8(tmem + lane_base + blk[j] * kTcW + f0 + c * 8, z);
            if (j == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) z[e] += bias_s[f0 + c * 8 + e];
            }
            if (add_skip) {
#pragma unroll
              for (int e = 0; e < 8; ++e) z[e] += skip_s[(j * kTcW + f0 + c * 8 + e) * 128];
            }
          }
          float a[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int ff = c * 8 + e;
            zst[(j * kTcW + f0 + ff) * 128] = z[e];          // stash for the reverse pass
            if (j == 0) {
              float c0, c1, c2, c3;
              act_coeffs<false>(sv.act, z[e], c0, c1, c2, c3);
              s1[ff] = c1;
              s2[ff] = c2;
              a[e] = c0;
            } else if (is_first) {
              zi_keep[ff] = z[e];
              a[e] = s1[ff] * z[e];
            } else {
              a[e] = fmaf(s2[ff] * zi_keep[ff], zi_keep[ff], s1[ff] * z[e]);
            }
          }
          if (save_skip) {
#pragma unroll
            for (int e = 0; e < 8; ++e) skip_s[(j * kTcW + f0 + c * 8 + e) * 128] = a[e];
          }
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_tf32(a[e], hi[e], lo[e]);
          if (c == 0) {   // the GEMM issued two channels ago still reads this buffer
            if (buf == 0 && pending0) { mbar_wait(&mma_bar[0], bar_phase0); bar_phase0 ^= 1; pending0 = 0; }
            if (buf == 1 && pending1) { mbar_wait(&mma_bar[1], bar_phase1); bar_phase1 ^= 1; pending1 = 0; }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t off = tile_kb + sw128_offset(p, col0 + c * 8 + h * 4);
            *reinterpret_cast<float4*>(at + off) = make_float4(hi[h * 4], hi[h * 4 + 1], hi[h * 4 + 2], hi[h * 4 + 3]);
            *reinterpret_cast<float4*>(at + 2 * kRegionBytes + off) =
                make_float4(lo[h * 4], lo[h * 4 + 1], lo[h * 4 + 2], lo[h * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        // blocks of the source layer that nobody needs any more
        if (S >= 1) {
          if (j == 0) free_mask |= 1u << blk[0];
          else if (ORDER == 1) free_mask |= 1u << blk[j];
          else if (is_second) free_mask |= (1u << blk[j]) | (1u << blk[1 + dir]);
        }
        const int nb = __ffs(free_mask) - 1;
        free_mask &= ~(1u << nb);
        nblk[j] = nb;
        if (t == 0) {
          tc_fence_after_sync();
          issue_gemm_kmajor(tmem + nb * kTcW, smem_u32(at), smem_u32(w_buf));
          mma_commit(&mma_bar[buf]);
        }
        if (buf == 0) pending0 = 1; else pending1 = 1;
      }
      if (pending0) { mbar_wait(&mma_bar[0], bar_phase0); bar_phase0 ^= 1; pending0 = 0; }
      if (pending1) { mbar_wait(&mma_bar[1], bar_phase1); bar_phase1 ^= 1; pending1 = 0; }
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < J; ++j) blk[j] = nblk[j];
    }

    // ---- output op: LINEAR(64 -> o) on the FMA pipe, fed from the last hidden layer's TMEM blocks
    {
      const DevOp& op = prog.ops[n_ops - 1];
      const DevOp& sv = prog.ops[n_ops - 2];
      const float* W = params + op.w_off;
      float* zst = tstash + sv.stash_plane * 128;
      if (t < kTcW) bias_s[t] = __ldg(params + sv.b_off + t);
      __syncthreads();
      float acc[J][4];
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.0f;
#pragma unroll
      for (int c = 0; c < kFpt / 8; ++c) {
        float zj[J][8];
#pragma unroll
        for (int j = 0; j < J; ++j) tmem_ld8(tmem + lane_base + blk[j] * kTcW + f0 + c * 8, zj[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          float z[J];
#pragma unroll
          for (int j = 0; j < J; ++j) z[j] = zj[j][e];
          z[0] += bias_s[f];
          if (sv.flags & MPINN_FLAG_ADD_SKIP) {
#pragma unroll
            for (int j = 0; j < J; ++j) z[j] += skip_s[(j * kTcW + f) * 128];
          }
#pragma unroll
          for (int j = 0; j < J; ++j) zst[(j * kTcW + f) * 128] = z[j];
          act_jets<D, ORDER>(sv.act, z);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < o) {
              const float w = __ldg(W + k * kTcW + f);
#pragma unroll
              for (int j = 0; j < J; ++j) acc[j][k] = fmaf(w, z[j], acc[j][k]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) free_mask |= 1u << blk[j];
      // combine the feature quarters of each point through shared memory, 8 outputs at a time
      for (int base = 0; base < J * o; base += 8) {
        if (fq > 0) {
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = j * o + k - base;
              if (k < o && idx >= 0 && idx < 8) red_s[((fq - 1) * 128 + p) * 8 + idx] = acc[j][k];
            }
        }
        __syncthreads();
        if (fq == 0 && valid) {
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = j * o + k - base;
              if (k < o && idx >= 0 && idx < 8) {
                float v = acc[j][k];
                for (int q = 0; q < kTcThreads / 128 - 1; ++q) v += red_s[(q * 128 + p) * 8 + idx];
                if (j == 0) v += __ldg(params + op.b_off + k);
                out[((int64_t)j * n + pt) * o + k] = v;
              }
            }
        }
        __syncthreads();
      }
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// reverse
// ---------------------------------------------------------------------------------------------
// Per 128-point tile, walking the layers backwards.  For hidden layer l (64 -> 64) and jet channel j:
//     zbar_j   = act'(z^l) applied to abar_j (abar lives in TMEM, left there by the previous GEMM)
//     a_j      = activation jets of layer l-1, recomputed from the stash
//     abar'_j  = zbar_j * W^l            tcgen05 BF16x3, K-major views     -> TMEM cols [64j, 64j+64)
//     dW^l    += zbar_j^T * a_j          tcgen05 BF16x3, MN-major views of the SAME staged tiles
// Channels are processed first-derivative, second-derivative (per direction), value last, so that every
// abar_j is consumed before the GEMM that overwrites its TMEM columns is issued, and the value
// channel's zbar can collect the cross terms of all directions.  The transcendental part of every
// activation is evaluated once per (point, feature, layer) and kept as two basis values in shared memory.
struct TcBwdSmem {
  static constexpr int kZ = 0;                               // zbar tile: 3 pieces + ZERO   64 KiB
  static constexpr int kH = kZ + kBfTile;                    // a tile: 3 pieces             48 KiB
  static constexpr int kW = kH + 3 * kBfRegion;              // W^T operand, 3 pieces        24 KiB
  static constexpr int kCache = kW + kBfWTile;               // basis values [2][64][128]    64 KiB
  static constexpr int kRed = kCache + 2 * 64 * 128 * 4;     // [64][8] reduction scratch      2 KiB
  static constexpr int kTotal = kRed + 64 * 8 * 4 + 16 * 16 * 4 + 1024;
  static constexpr int kXchStride = 65;                      // dW half exchange aliases the a tile
};

// one thread asks the L2 to fetch `bytes` (multiple of 16) starting at p
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Sum each of 16 per-thread values over the 32 lanes of the warp with 16 shuffles (instead of 80):
// fold the value array in half at every step while exchanging with the lane 16, 8, 4, 2 away.
// On return every lane holds the warp total of value index (lane >> 1).
__device__ __forceinline__ float warp_sum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float keep = up ? v[k + 8] : v[k], send = up ? v[k] : v[k + 8];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float keep = up ? a[k + 4] : a[k], send = up ? a[k] : a[k + 4];
      b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float keep = up ? b[k + 2] : b[k], send = up ? b[k] : b[k + 2];
      c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = lane & 2;
  float r = (up ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, up ? c[0] : c[1], 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// fire-and-forget 16-byte vector reduction (sm_90+): *p[0..3] += v
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int D, int ORDER>
__global__ void __launch_bounds__(kTcThreads, 1)
jet_bwd_tc(const __grid_constant__ DevProg prog, const float* __restrict__ params,
           const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
           const float* __restrict__ dout, const float* __restrict__ stash, int64_t n_pad,
           float* __restrict__ part_all, int64_t n_params, float* __restrict__ skipbar_all, int n_tiles) {
  constexpr int J = 1 + ORDER * D;
  constexpr uint32_t kDwCol = 448;  // TMEM columns of the dW accumulator
  constexpr uint32_t kT0Col = 320;  // TMEM columns caching the activation value of the layer below (J <= 5)
  constexpr uint32_t kT3Col = 384;  // TMEM columns caching its third derivative
  static_assert(J * kTcW <= (int)kT0Col, "tcgen05 reverse kernel supports at most 5 jet channels");
  constexpr int NC = kFpt / 8;      // 8-feature chunks per thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* z_tile = smem + TcBwdSmem::kZ;
  uint8_t* h_tile = smem + TcBwdSmem::kH;
  uint8_t* w_tile = smem + TcBwdSmem::kW;
  float* cache = reinterpret_cast<float*>(smem + TcBwdSmem::kCache);  // [2][64][128]
  float* red_s = reinterpret_cast<float*>(smem + TcBwdSmem::kRed);    // [64][8]
  float* bsum_s = red_s + 64 * 8;                                     // [16 warps][16] bias-gradient partials
  float* xch = reinterpret_cast<float*>(h_tile);                      // [64][65] (only between GEMM groups)
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int t = threadIdx.x;
  const int p = t % 128, fq = t / 128, warp = t / 32, lane = t % 32;
  const uint32_t lane_base = (uint32_t)((warp % 4) * 32) << 16;
  const int f0 = fq * kFpt;
  // gradient travelling back along a skip connection: this CTA's [channel][feature][point] scratch
  float* sb = skipbar_all + (int64_t)blockIdx.x * (J * kTcW * 128) + p;

  const int64_t slab = (n_params + 3) & ~int64_t(3);   // 16-byte aligned slabs
  float* part = part_all + (int64_t)blockIdx.x * slab;
  for (int64_t i = t; i < n_params; i += kTcThreads) part[i] = 0.0f;
  for (int i = t; i < kBfRegion / 16; i += kTcThreads)
    reinterpret_cast<float4*>(z_tile + 3 * kBfRegion)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t == 0) {
    mbar_init(&mma_bar, 2);   // two issuing threads commit per phase
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t bar_phase = 0;
  int pending = 0;
  const int o = prog.n_out;
  const int n_ops = prog.n_ops;

  // activation coefficients s0..s3 (value and first three derivatives) of layer `op` at this thread's
  // (point, features), evaluated once: s1/s2 go to shared memory, s0/s3 to spare TMEM columns
  auto fill_cache = [&](const DevOp& op, const float* tstash) {
    const float* z0p = tstash + op.stash_plane * 128;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float z[8], c0[8], c3[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = __ldg(z0p + (f0 + c * 8 + e) * 128);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = f0 + c * 8 + e;
        float c1, c2;
        act_coeffs<(ORDER >= 2)>(op.act, z[e], c0[e], c1, c2, c3[e]);
        cache[f * 128 + p] = c1;
        cache[(64 + f) * 128 + p] = c2;
      }
      tmem_st8(tmem + lane_base + kT0Col + f0 + c * 8, c0);
      tmem_st8(tmem + lane_base + kT3Col + f0 + c * 8, c3);
    }
  };
  // sum a per-thread value over the warp's 32 points and add it to the CTA scratch row f, slot < 8
  auto reduce_into = [&](float v, int f, int slot) {
    v = warp_sum(v);
    if (lane == 0) atomicAdd(&red_s[f * 8 + slot], v);
  };

#if MPINN_TC_TIMING
  long long dbg_last = clock64();
#endif
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * 128 + p;
    TC_TIME(15);
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;

    const float* tstash = stash + (int64_t)tile * prog.total_planes * 128 + p;
    if (t == 0 && tile == (int)blockIdx.x) {
      // first tile of this CTA: the two stash blocks the first hidden iteration reads
      l2_prefetch_bulk(tstash - p + prog.ops[n_ops - 2].stash_plane * 128, (uint32_t)(J * kTcW * 128 * 4));
      if (n_ops >= 4) l2_prefetch_bulk(tstash - p + prog.ops[n_ops - 3].stash_plane * 128, (uint32_t)(J * kTcW * 128 * 4));
    }

    // =========================== output op (64 -> o), FMA pipe ===========================
    {
      const DevOp& op = prog.ops[n_ops - 1];
      const DevOp& pv = prog.ops[n_ops - 2];
      float dz[J][4];
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dz[j][k] = (valid && k < o) ? __ldg(dout + ((int64_t)j * n + pt) * o + k) : 0.0f;
      for (int i = t; i < 64 * 8; i += kTcThreads) red_s[i] = 0.0f;
      fill_cache(pv, tstash);
      __syncthreads();
      const float* W = params + op.w_off;
      const float* zp = tstash + pv.stash_plane * 128;
      // abar^{L-1}_j[f] = sum_k W[k][f] dz[j][k]  -> TMEM, 8 features at a time
#pragma unroll
      for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int f = f0 + c * 8 + e;
            float a = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < o) a = fmaf(__ldg(W + k * kTcW + f), dz[j][k], a);
            v[e] = a;
          }
          tmem_st8(tmem + lane_base + j * kTcW + f0 + c * 8, v);
        }
      }
      // dW_last[k][f] += sum_pts sum_j dz[j][k] a_j[f] ; db_last[k] += sum_pts dz[0][k]
      float g[4][kFpt];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float t0[8];
        tmem_ld8(tmem + lane_base + kT0Col + f0 + c * 8, t0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          const float s1 = cache[f * 128 + p], s2 = cache[(64 + f) * 128 + p];
          float a[J];
          a[0] = t0[e];
          if (ORDER >= 1) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
              const float zi = __ldg(zp + ((1 + i) * kTcW + f) * 128);
              a[1 + i] = s1 * zi;
              if (ORDER >= 2) a[1 + D + i] = fmaf(s2 * zi, zi, s1 * __ldg(zp + ((1 + D + i) * kTcW + f) * 128));
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) acc = fmaf(dz[j][k], a[j], acc);
            g[k][c * 8 + e] = acc;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < o) {
          const float tot = warp_sum16(g[k], lane);
          if ((lane & 1) == 0) atomicAdd(&red_s[(f0 + (lane >> 1)) * 8 + k], tot);
        }
      }
      if (fq == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < o) reduce_into(dz[0][k], k, 4);   // bias gradient rides in slot 4 of row k
      }
      __syncthreads();
      for (int i = t; i < 64 * 4; i += kTcThreads) {
        const int f = i / 4, k = i % 4;
        if (k < o) part[op.w_off + k * kTcW + f] += red_s[f * 8 + k];
      }
      if (t < o) part[op.b_off + t] += red_s[t * 8 + 4];
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
    }

    TC_TIME(0);   // output op
    // =========================== hidden layers, tensor cores ===========================
    for (int l = n_ops - 2; l >= 1; --l) {
      const DevOp& op = prog.ops[l];
      const DevOp& pv = prog.ops[l - 1];
      const float* zl = tstash + op.stash_plane * 128;    // z^l jets
      const float* zm = tstash + pv.stash_plane * 128;    // z^{l-1} jets
      const bool store_sb = (op.flags & MPINN_FLAG_ADD_SKIP);                  // zbar^l also flows to the block input
      const bool add_sb = (prog.ops[l + 1].flags & MPINN_FLAG_SAVE_SKIP);      // abar^l gets the skip branch of the block above
      // first three derivatives of layer l's activation for this thread's features, from the cached basis
      float s1[kFpt], s2[kFpt], s3[kFpt], acc0[kFpt];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float v3[8];
        tmem_ld8(tmem + lane_base + kT3Col + f0 + c * 8, v3);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ff = c * 8 + e, f = f0 + ff;
          s1[ff] = cache[f * 128 + p];
          s2[ff] = cache[(64 + f) * 128 + p];
          s3[ff] = v3[e];
          acc0[ff] = 0.0f;
        }
      }
      for (int i = t; i < 64 * 8; i += kTcThreads) red_s[i] = 0.0f;
      __syncthreads();                 // (the caches are thread-private; this orders red_s)
      fill_cache(pv, tstash);          // now the cache holds layer l-1
      {
        const float4* src = reinterpret_cast<const float4*>(
            reinterpret_cast<const uint8_t*>(packed + op.tc_off) + kWTileBytes);
        float4* dst = reinterpret_cast<float4*>(w_tile);
        for (int i = t; i < kBfWTile / 16; i += kTcThreads) dst[i] = __ldg(src + i);
      }
      __syncthreads();

      // the layer below will be needed next (its z as "z^l", and the one under it as "z^{l-1}"):
      // pull that tile-layer block (J*64 planes x 512 B, contiguous) into L2 while this layer computes
      if (t == 0 && l >= 2) {
        const DevOp& nx = prog.ops[l - 2];
        l2_prefetch_bulk(tstash - p + nx.stash_plane * 128, (uint32_t)(J * kTcW * 128 * 4));
      }
      if (t == 0 && l == 1 && tile + (int)gridDim.x < n_tiles) {   // next tile's two top blocks
        const float* nt = stash + (int64_t)(tile + gridDim.x) * prog.total_planes * 128;
        l2_prefetch_bulk(nt + prog.ops[n_ops - 2].stash_plane * 128, (uint32_t)(J * kTcW * 128 * 4));
        if (n_ops >= 4) l2_prefetch_bulk(nt + prog.ops[n_ops - 3].stash_plane * 128, (uint32_t)(J * kTcW * 128 * 4));
      }
      TC_TIME(1);   // per-layer setup: coefficients, cache fill, W tile
      uint32_t dw_acc = 0;
      // channel schedule: (d/dx_i, d2/dx_i2) per direction, value channel last
#pragma unroll 1
      for (int step = 0; step < J; ++step) {
        int jet;
        if (step == J - 1) jet = 0;
        else if (ORDER == 2) jet = (step & 1) ? 1 + D + (step >> 1) : 1 + (step >> 1);
        else jet = 1 + step;
        const bool is_first = (jet >= 1 && jet <= D);
        const bool is_second = (jet > D);
        const int dir = is_first ? jet - 1 : (is_second ? jet - 1 - D : 0);
        const bool two = is_first && ORDER == 2;

        float zb0all[kFpt];                      // value-channel zbar of this phase (bias gradient)
#pragma unroll
        for (int c = 0; c < NC; ++c) {          // 8 features per chunk = one 16-byte bf16 store per piece
          const uint32_t off = sw128_offset_bf16(p, f0 + c * 8);
          // ---------------- zbar of layer l ----------------
          {
            float ab[8], ab2[8], zi[8], zii[8], zb[8];
            tmem_ld8(tmem + lane_base + jet * kTcW + f0 + c * 8, ab);
            if (two) tmem_ld8(tmem + lane_base + (1 + D + dir) * kTcW + f0 + c * 8, ab2);
            if (add_sb) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int f = f0 + c * 8 + e;
                ab[e] += sb[(jet * kTcW + f) * 128];
                if (two) ab2[e] += sb[((1 + D + dir) * kTcW + f) * 128];
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int f = f0 + c * 8 + e;
              zi[e] = is_first ? __ldg(zl + (jet * kTcW + f) * 128) : 0.0f;
              zii[e] = two ? __ldg(zl + ((1 + D + dir) * kTcW + f) * 128) : 0.0f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ff = c * 8 + e;
              if (jet == 0) {
                zb[e] = fmaf(ab[e], s1[ff], acc0[ff]);
              } else if (is_first) {
                float v = ab[e] * s1[ff];
                float a0 = fmaf(ab[e] * s2[ff], zi[e], acc0[ff]);
                if (ORDER == 2) {
                  v = fmaf(2.0f * ab2[e] * s2[ff], zi[e], v);
                  a0 = fmaf(ab2[e], fmaf(s3[ff] * zi[e], zi[e], s2[ff] * zii[e]), a0);
                }
                acc0[ff] = a0;
                zb[e] = v;
              } else {
                zb[e] = ab[e] * s1[ff];
              }
            }
            if (jet == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) zb0all[c * 8 + e] = zb[e];
            }
            if (store_sb) {
#pragma unroll
              for (int e = 0; e < 8; ++e) sb[(jet * kTcW + f0 + c * 8 + e) * 128] = zb[e];
            }
            uint32_t q1[4], q2[4], q3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x3(zb[2 * e], zb[2 * e + 1], q1[e], q2[e], q3[e]);
            // The GEMMs of the previous phase read both operand tiles but none of the TMEM columns this
            // phase has touched so far, so only now -- right before overwriting the tiles -- wait for them.
            if (c == 0) TC_TIME(2);   // loads + zbar math of chunk 0
            if (c == 0 && pending) {
              mbar_wait(&mma_bar, bar_phase);
              bar_phase ^= 1;
              pending = 0;
            }
            if (c == 0) TC_TIME(3);   // waiting for the previous phase's GEMMs
            *reinterpret_cast<uint4*>(z_tile + off) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
            *reinterpret_cast<uint4*>(z_tile + kBfRegion + off) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
            *reinterpret_cast<uint4*>(z_tile + 2 * kBfRegion + off) = make_uint4(q3[0], q3[1], q3[2], q3[3]);
          }
          // ---------------- activation jet channel of layer l-1 ----------------
          {
            float yi[8], yii[8], av[8];
            if (jet == 0) {
              tmem_ld8(tmem + lane_base + kT0Col + f0 + c * 8, av);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int f = f0 + c * 8 + e;
                yi[e] = __ldg(zm + ((1 + dir) * kTcW + f) * 128);
                yii[e] = is_second ? __ldg(zm + (jet * kTcW + f) * 128) : 0.0f;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int f = f0 + c * 8 + e;
                const float t1 = cache[f * 128 + p];
                av[e] = is_first ? t1 * yi[e] : fmaf(cache[(64 + f) * 128 + p] * yi[e], yi[e], t1 * yii[e]);
              }
            }
            uint32_t q1[4], q2[4], q3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x3(av[2 * e], av[2 * e + 1], q1[e], q2[e], q3[e]);
            *reinterpret_cast<uint4*>(h_tile + off) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
            *reinterpret_cast<uint4*>(h_tile + kBfRegion + off) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
            *reinterpret_cast<uint4*>(h_tile + 2 * kBfRegion + off) = make_uint4(q3[0], q3[1], q3[2], q3[3]);
          }
        }
        if (jet == 0) {
          static_assert(kFpt == 16, "warp_sum16 reduces exactly 16 features per thread");
          const float tot = warp_sum16(zb0all, lane);
          if ((lane & 1) == 0) bsum_s[warp * 16 + (lane >> 1)] = tot;
        }
        TC_TIME(4);     // rest of the phase's staging
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        TC_TIME(5);     // phase barrier
        if (t == 0) {            // one lane of warp 0 issues the input-gradient GEMM ...
          tc_fence_after_sync();
          issue_bf16x3_kmajor(tmem + jet * kTcW, smem_u32(z_tile), smem_u32(w_tile));
          mma_commit(&mma_bar);
        } else if (t == 32) {    // ... one lane of warp 1 the weight-gradient GEMM (independent accumulators)
          tc_fence_after_sync();
          issue_bf16x3_mnmajor(tmem + kDwCol, smem_u32(z_tile), smem_u32(h_tile), dw_acc);
          mma_commit(&mma_bar);
        }
        TC_TIME(6);     // MMA issue (thread 0 only)
        dw_acc = 1;
        pending = 1;
      }
      mbar_wait(&mma_bar, bar_phase);
      TC_TIME(7);       // final GEMM wait of the layer
      bar_phase ^= 1;
      pending = 0;
      tc_fence_after_sync();
      // ---- flush this (tile, layer)'s dW into the CTA's private slab with fire-and-forget reductions
      // (no round trip is exposed); accumulator rows r and r+64 both belong to output feature r % 64
      {
        const int row = ((warp % 4) * 32 + lane) & 63;
        float* dst = part + op.w_off + row * kTcW + f0;
        const bool vec_ok = ((op.w_off & 3) == 0);   // slabs are 16-byte aligned (see launch_bwd_tc)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float v[8];
          tmem_ld8(tmem + lane_base + kDwCol + f0 + c * 8, v);
          if (vec_ok) {
            red_add_v4(dst + c * 8, v[0], v[1], v[2], v[3]);
            red_add_v4(dst + c * 8 + 4, v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(dst + c * 8 + e, v[e]);
          }
        }
      }
      if (t < 64) {   // feature t was summed by the 4 warps owning feature quarter t/16
        const int q = t / kFpt, k = t % kFpt;
        part[op.b_off + t] += bsum_s[(q * 4 + 0) * 16 + k] + bsum_s[(q * 4 + 1) * 16 + k] +
                              bsum_s[(q * 4 + 2) * 16 + k] + bsum_s[(q * 4 + 3) * 16 + k];
      }
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
      TC_TIME(8);       // dW flush
    }

    // =========================== op 0 (d -> 64), FMA pipe ===========================
    {
      const DevOp& op = prog.ops[0];
      const float* z0p = tstash + op.stash_plane * 128;
      for (int i = t; i < 64 * 8; i += kTcThreads) red_s[i] = 0.0f;
      __syncthreads();
      const bool add_sb0 = (prog.ops[1].flags & MPINN_FLAG_SAVE_SKIP);
      float zb0[kFpt], zbd[D > 0 ? D : 1][kFpt];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float ab0[8];
        tmem_ld8(tmem + lane_base + f0 + c * 8, ab0);
        if (add_sb0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) ab0[e] += sb[(f0 + c * 8 + e) * 128];
        }
        float s1[8], s2[8], s3[8];
        tmem_ld8(tmem + lane_base + kT3Col + f0 + c * 8, s3);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          s1[e] = cache[f * 128 + p];
          s2[e] = cache[(64 + f) * 128 + p];
          zb0[c * 8 + e] = ab0[e] * s1[e];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float abi[8], abii[8];
          if (ORDER >= 1) tmem_ld8(tmem + lane_base + (1 + i) * kTcW + f0 + c * 8, abi);
          if (ORDER == 2) tmem_ld8(tmem + lane_base + (1 + D + i) * kTcW + f0 + c * 8, abii);
          if (add_sb0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (ORDER >= 1) abi[e] += sb[((1 + i) * kTcW + f0 + c * 8 + e) * 128];
              if (ORDER == 2) abii[e] += sb[((1 + D + i) * kTcW + f0 + c * 8 + e) * 128];
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int f = f0 + c * 8 + e;
            float zbi = 0.0f;
            if (ORDER >= 1) {
              const float zi = __ldg(z0p + ((1 + i) * kTcW + f) * 128);
              zbi = abi[e] * s1[e];
              zb0[c * 8 + e] = fmaf(abi[e] * s2[e], zi, zb0[c * 8 + e]);
              if (ORDER == 2) {
                const float zii = __ldg(z0p + ((1 + D + i) * kTcW + f) * 128);
                zbi = fmaf(2.0f * abii[e] * s2[e], zi, zbi);
                zb0[c * 8 + e] = fmaf(abii[e], fmaf(s3[e] * zi, zi, s2[e] * zii), zb0[c * 8 + e]);
              }
            }
            zbd[i][c * 8 + e] = zbi;          // d z_i / d W[f][i] = 1
          }
        }
      }
      {
        const int fr = f0 + (lane >> 1);
        const bool w = (lane & 1) == 0;
        float tot = warp_sum16(zb0, lane);    // bias
        if (w) atomicAdd(&red_s[fr * 8 + 0], tot);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float tmp[kFpt];
#pragma unroll
          for (int e = 0; e < kFpt; ++e) tmp[e] = fmaf(zb0[e], xv[i], zbd[i][e]);   // dL/dW[f][i]
          tot = warp_sum16(tmp, lane);
          if (w) atomicAdd(&red_s[fr * 8 + 4 + i], tot);
        }
      }
      __syncthreads();
      for (int i = t; i < 64 * D; i += kTcThreads) {
        const int f = i / D, k = i % D;
        part[op.w_off + f * D + k] += red_s[f * 8 + 4 + k];
      }
      if (t < 64) part[op.b_off + t] += red_s[t * 8 + 0];
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
    }
  }
  TC_TIME(9);           // op 0 of the last tile (+ tail)
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
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
    if (op.kind != MPINN_OP_LINEAR || op.in_dim != kTcW || op.out_dim != kTcW || op.tc_off < 0) return false;
    // residual blocks: a SAVE_SKIP layer immediately followed by the ADD_SKIP layer that closes the block
    if (op.flags == MPINN_FLAG_SAVE_SKIP) {
      if (l + 1 > prog.n_ops - 2 || prog.ops[l + 1].flags != MPINN_FLAG_ADD_SKIP) return false;
    } else if (op.flags == MPINN_FLAG_ADD_SKIP) {
      if (prog.ops[l - 1].flags != MPINN_FLAG_SAVE_SKIP) return false;
    } else if (op.flags != 0) {
      return false;
    }
  }
  const DevOp& last = prog.ops[prog.n_ops - 1];
  return last.kind == MPINN_OP_LINEAR && last.in_dim == kTcW && last.act == MPINN_ACT_NONE && !last.flags;
}

bool tc_backward_supported(const DevProg& prog, int d, int order) {
  // s0 / s3 of the layer below are cached in the 128 TMEM columns that 5 channels leave free
  return (1 + order * d) <= 5 && tc_forward_supported(prog, d, order);
}

template <int D, int ORDER>
static cudaError_t launch_fwd_tc_t(const DevProg& prog, const float* params, const float* packed,
                                   const float* x, int64_t n, float* out, float* stash, int64_t n_pad,
                                   int sm_count, cudaStream_t stream) {
  auto kern = jet_fwd_tc<D, ORDER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::kTotal);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n + 127) / 128);
  int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid > 148) grid = 148;   // the skip scratch holds 148 CTA regions
  kern<<<grid, kTcBlock, TcSmem::kTotal, stream>>>(prog, params, packed, x, n, out, stash, n_pad, n_tiles);
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

// dparams[i] += sum_c part[c][i]  (same fold as the FFMA engine)
__global__ void fold_partials_tc(const float* __restrict__ part, int n_slabs, int64_t n_params,
                                 float* __restrict__ dparams) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  float s = 0.0f;
  const int64_t slab = (n_params + 3) & ~int64_t(3);
  for (int c = 0; c < n_slabs; ++c) s += part[(int64_t)c * slab + i];
  dparams[i] += s;
}

template <int D, int ORDER>
static cudaError_t launch_bwd_tc_t(const DevProg& prog, const float* params, const float* packed,
                                   const float* x, int64_t n, const float* dout, const float* stash,
                                   int64_t n_pad, float* part, int max_slabs, int64_t n_params,
                                   float* dparams, float* skip_scratch, int sm_count, cudaStream_t stream) {
  auto kern = jet_bwd_tc<D, ORDER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwdSmem::kTotal);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n + 127) / 128);
  int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid > max_slabs - 1) grid = max_slabs - 1;   // slabs are padded to 16 bytes: keep one in reserve
  if (grid < 1) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(part) & 15) != 0) return cudaErrorMisalignedAddress;
  if (grid > 148) grid = 148;   // the skip scratch holds 148 CTA regions
  kern<<<grid, kTcThreads, TcBwdSmem::kTotal, stream>>>(prog, params, packed, x, n, dout, stash, n_pad, part,
                                                         n_params, skip_scratch, n_tiles);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fold_partials_tc<<<(unsigned)((n_params + 255) / 256), 256, 0, stream>>>(part, grid, n_params, dparams);
  return cudaGetLastError();
}

cudaError_t launch_bwd_tc(const DevProg& prog, const float* params, const float* packed, const float* x,
                          int64_t n, int d, int order, const float* dout, const float* stash, int64_t n_pad,
                          float* part, int max_slabs, int64_t n_params, float* dparams, float* dx,
                          float* skip_scratch, int sm_count, cudaStream_t stream) {
  if (dx != nullptr) return cudaErrorNotSupported;
#define TC_BWD(D_, O_) \
  if (d == D_ && order == O_) \
    return launch_bwd_tc_t<D_, O_>(prog, params, packed, x, n, dout, stash, n_pad, part, max_slabs, \
                                   n_params, dparams, skip_scratch, sm_count, stream)
  TC_BWD(1, 0); TC_BWD(1, 1); TC_BWD(1, 2);
  TC_BWD(2, 0); TC_BWD(2, 1); TC_BWD(2, 2);
  TC_BWD(3, 0); TC_BWD(3, 1);
#undef TC_BWD
  return cudaErrorInvalidValue;
}

}  // namespace mpinn
