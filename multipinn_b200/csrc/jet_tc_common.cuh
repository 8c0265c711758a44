// jet_tc_common.cuh -- constants, operand-tile layouts and tcgen05 issue helpers shared by the forward
// (jet_tc.cu) and reverse (jet_tc_bwd.cu) translation units of the tensor-core engine.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "launch.h"
#include "umma.cuh"

namespace mpinn {

using namespace umma;

constexpr int kTcW = 64;                 // layer width handled by this engine
constexpr int kTcThreads = 512;          // 16 warps: 4 threads per point, 16 features each
constexpr int kFpt = kTcW / (kTcThreads / 128);  // features per thread
constexpr int kTileRows = 128;
constexpr int kRegionBytes = kTileRows * 128;       // [128 rows x 32 fp32]  = 16 KiB
constexpr int kATileBytes = 4 * kRegionBytes;       // hi kb0, hi kb1, lo kb0, lo kb1 = 64 KiB
constexpr int kWRegionBytes = kTcW * 128;           // [64 rows x 32 fp32]   = 8 KiB
constexpr int kWTileBytes = 4 * kWRegionBytes;      // hi kb0, hi kb1, lo kb0, lo kb1 = 32 KiB
// packed block per layer: forward operand (B[n=out][k=in] = W) then reverse operand (B[n=in][k=out] = W^T)
constexpr int kTcBlockFloats = 2 * kWTileBytes / 4;


// activation applied to feature f0.. of op (a thread's 16 features never straddle the sin/cos halves)
__device__ __forceinline__ int act_of(const DevOp& op, int f0) {
  if (op.kind == MPINN_OP_FOURIER) return f0 < op.out_dim / 2 ? MPINN_ACT_SINE : kActCos;
  return op.act;
}
// weight of input i for feature f of the first op: W[f][i] (LINEAR) or B[i][f mod m] (FOURIER)
__device__ __forceinline__ float first_w(const DevOp& op, const float* __restrict__ params, int f, int i, int d) {
  if (op.kind == MPINN_OP_FOURIER) {
    const int m = op.out_dim / 2;
    return __ldg(params + op.w_off + i * m + (f % m));
  }
  return __ldg(params + op.w_off + f * d + i);
}

constexpr int kBfRegion = kTileRows * 128;    // one bf16 piece of a [128 x 64] tile: 16 KiB
constexpr int kBfTile = 4 * kBfRegion;        // piece1, piece2, piece3, ZERO region
constexpr int kBfWRegion = kTcW * 128;        // one bf16 piece of a [64 x 64] weight operand: 8 KiB
constexpr int kBfWTile = 3 * kBfWRegion;

// byte offset of bf16 element (row, col<64) inside a swizzled region
__host__ __device__ __forceinline__ uint32_t sw128_offset_bf16(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

// ---------------------------------------------------------------------------------------------
// MMA issue helpers (called by one thread)
// ---------------------------------------------------------------------------------------------
// Z[128 x 64] (+)= A[128 x 64] * B[64 x 64]^T, 3xTF32, A/B K-major tiles (hi kb0, hi kb1, lo kb0, lo kb1)
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
      mma_tf32_ss(d_tmem, a0 + (a_off >> 4), b0 + (b_off >> 4), idesc, acc);
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
      mma_bf16_ss(d_tmem, ad, bd, idesc, acc);
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
  const uint64_t z1 = make_desc_sw128(z_tile, 1 * kBfRegion, 1024);   // M halves one region apart
  const uint64_t z2 = make_desc_sw128(z_tile, 2 * kBfRegion, 1024);   // M halves two regions apart
  const uint64_t h0 = make_desc_sw128(h_tile, kBfRegion, 1024);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {   // K = 16 points = two 8-row swizzle atoms (SBO = 1024)
      const uint64_t zd = (zs[q] == 1 ? z1 : z2) + ((zr[q] * kBfRegion + ks * 2048) >> 4);
      mma_bf16_ss(d_tmem, zd, h0 + ((hp[q] * kBfRegion + ks * 2048) >> 4), idesc, acc);
      acc = 1;
    }
  }
}

// The same product with the two leading H pieces N-stacked: ONE M128 N128 group computes Z1^T H1 | Z1^T H2 (rows 0..63)
// and Z2^T H1 | Z2^T H2 (rows 64..127) -- four piece products for 8 KB of shared-memory operands per MMA instead of
// two groups at 6 KB each -- into 128 accumulator columns; the two small products go to the first 64 columns:
//   cols 0..63  : rows 0..63  Z1^T H1 + Z1^T H3 + Z3^T H1      rows 64..127  Z2^T H1 (+ Z3^T H3, negligible)
//   cols 64..127: rows 0..63  Z1^T H2                           rows 64..127  Z2^T H2
// The caller adds the four quadrants.  In a layer's first phase (accumulate_first == 0) the wide group goes first: it
// is the one MMA sequence that touches all 128 columns, so it carries the "overwrite" flag; in the other phases the
// two small products go first (the tensor core truncates when it adds into the accumulator).
__device__ __forceinline__ void issue_bf16x3_mnmajor_wide(uint32_t d_tmem, uint32_t z_tile, uint32_t h_tile,
                                                          uint32_t accumulate_first) {
  constexpr uint32_t idesc_w = make_idesc_bf16(128, 128, 1, 1);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
  const uint64_t z1 = make_desc_sw128(z_tile, 1 * kBfRegion, 1024);   // (Z1 ; Z2)
  const uint64_t z2 = make_desc_sw128(z_tile, 2 * kBfRegion, 1024);   // (Z1 ; Z3)
  const uint64_t h0 = make_desc_sw128(h_tile, kBfRegion, 1024);       // N halves one region apart: [H1 | H2]
  auto wide = [&](uint32_t acc) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      mma_bf16_ss(d_tmem, z1 + ((ks * 2048) >> 4), h0 + ((ks * 2048) >> 4), idesc_w, acc);
      acc = 1;
    }
  };
  auto small = [&]() {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)      // (Z3 ; 0)^T H1
      mma_bf16_ss(d_tmem, z1 + ((2 * kBfRegion + ks * 2048) >> 4), h0 + ((ks * 2048) >> 4), idesc, 1);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)      // (Z1 ; Z3)^T H3
      mma_bf16_ss(d_tmem, z2 + ((ks * 2048) >> 4), h0 + ((2 * kBfRegion + ks * 2048) >> 4), idesc, 1);
  };
  if (accumulate_first == 0) {
    wide(0);
    small();
  } else {
    small();
    wide(1);
  }
}

// compile-time loop: f(IC<I>{}) for I in [I0, N)
template <int I>
struct IC {
  static constexpr int value = I;
};
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for<I + 1, N>(f);
  }
}

static inline int tc_feat(const DevProg& prog) {
  if (prog.ops[0].kind == MPINN_OP_FOURIER) return 2;
  return prog.has_skip ? 1 : 0;
}


}  // namespace mpinn
