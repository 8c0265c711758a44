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

#include "jet_tc_common.cuh"

namespace mpinn {

int64_t tc_block_floats(int width) { return width == kTcW ? kTcBlockFloats : (width == 128 ? wide_block_floats() : 0); }

// optional cycle accounting of the forward kernel (build with -DMPINN_TC_TIMING=1, see jet_tc_bwd.cu): CTA 0,
// threads 0 and 224 accumulate clock64() deltas per code segment; mpinn_debug_counters() returns them as
// counters 32..63.
#ifndef MPINN_TC_TIMING
#define MPINN_TC_TIMING 0
#endif
__device__ long long g_tc_dbg_fwd[2][16];
#if MPINN_TC_TIMING
#define TCF_TIME(slot)                                                  \
  do {                                                                  \
    if (blockIdx.x == 0 && (t == 0 || t == 224)) {                      \
      const long long now_ = clock64();                                 \
      g_tc_dbg_fwd[t == 0 ? 0 : 1][slot] += now_ - dbg_last;            \
      dbg_last = now_;                                                  \
    }                                                                   \
  } while (0)
#else
#define TCF_TIME(slot) do {} while (0)
#endif

cudaError_t read_tc_debug_fwd(long long* out, int n) {
  long long h[32];
  cudaError_t e = cudaMemcpyFromSymbol(h, g_tc_dbg_fwd, sizeof(h));
  if (e != cudaSuccess) return e;
  for (int i = 0; i < n && i < 32; ++i) out[i] = h[i];
  long long z[32] = {0};
  return cudaMemcpyToSymbol(g_tc_dbg_fwd, z, sizeof(z));
}

// ---------------------------------------------------------------------------------------------
// weight packing: global image of the swizzled shared-memory operand tiles
// ---------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const __grid_constant__ DevProg prog, const float* __restrict__ params,
                               float* __restrict__ packed) {
  const DevOp& op = prog.ops[blockIdx.y];
  if (op.kind != MPINN_OP_LINEAR || op.tc_off < 0 || op.in_dim != kTcW) return;   // 128-wide blocks: jet_tc_wide.cu
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
  for (int l = 0; l < prog.n_ops; ++l) any |= prog.ops[l].tc_off >= 0 && prog.ops[l].in_dim == kTcW;
  if (!any) return cudaSuccess;
  dim3 grid(4, prog.n_ops);
  pack_tc_kernel<<<grid, 256, 0, stream>>>(prog, params, packed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// self-test kernels (diagnostics for the descriptor/layout conventions)
// ---------------------------------------------------------------------------------------------
// mode 0: C[128][64] = A[128][64] * B[64][64]^T               3xTF32, K-major operands
// mode 2: same product, BF16x3 K-major operands
// mode 3: C[128][64] = stacked(A)^T B, A,B [128][64]          BF16x3 MN-major operands (host adds halves)
// mode 4: C[128][128] = the same product with N-stacked B pieces (host adds the four quadrants)
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, const float* __restrict__ A,
                                                            const float* __restrict__ B, float* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* a_tile = smem;                    // 64 KiB
  uint8_t* b_tile = smem + kATileBytes;      // up to 64 KiB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (t < 32) tmem_alloc(&tmem_base_s, 128);
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
    else if (mode == 3) issue_bf16x3_mnmajor(tmem, smem_u32(a_tile), smem_u32(b_tile), 0);
    else issue_bf16x3_mnmajor_wide(tmem, smem_u32(a_tile), smem_u32(b_tile), 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int warp = t / 32;
  const int ncol = mode == 4 ? 128 : 64;
  for (int part = 0; part < ncol / 32; ++part) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + part * 32, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) C[t * ncol + part * 32 + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 128);
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
  static constexpr int kW = 2 * kATileBytes;                // two weight-operand buffers (next layer prefetched)
  static constexpr int kBias = kW + 2 * kWTileBytes;        // 2 x 64 floats
  static constexpr int kRed = kBias + 512;                  // [3 quarters][128 points][8] floats scratch
  static constexpr int kTotal = kRed + 3 * 128 * 8 * 4 + 1024;  // + alignment slack
};

// FEAT selects the program family at compile time so the plain-FNN kernel carries no extra registers:
//   0 = plain FNN, 1 = residual blocks (SAVE_SKIP/ADD_SKIP), 2 = Fourier-feature first op
//   KEEP = false: value-only / inference pass, nothing is written for a reverse pass (instantiated for order 0)
template <int D, int ORDER, int FEAT, bool KEEP = true>
__global__ void __launch_bounds__(kTcThreads, 1)
jet_fwd_tc(const __grid_constant__ DevProg prog, const float* __restrict__ params,
           const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
           float* __restrict__ out, float* __restrict__ stash, int64_t n_pad, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr bool LAP = (ORDER == 3);   // one second-order channel  L = sum_i c_i d2/dx_i2  (channel D+1)
  constexpr bool HAS_SKIP = (FEAT == 1), HAS_FOURIER = (FEAT == 2);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  float* bias_s = reinterpret_cast<float*>(smem + TcSmem::kBias);
  float* red_s = reinterpret_cast<float*>(smem + TcSmem::kRed);   // [3][128][8]
  __shared__ uint64_t mma_bar[2];
  __shared__ uint64_t w_bar[2];
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
    mbar_init(&w_bar[0], 1);
    mbar_init(&w_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  // Weight operands and biases run one stage ahead of their use: the TMA engine copies the next layer's
  // packed operand image into the other buffer while this layer computes.
  if (t < kTcW)
    bias_s[t] = (HAS_FOURIER && prog.ops[0].kind == MPINN_OP_FOURIER) ? 0.0f : __ldg(params + prog.ops[0].b_off + t);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t bar_phase0 = 0, bar_phase1 = 0;   // parity of the next completion to wait for, per A buffer
  int pending0 = 0, pending1 = 0;            // an MMA group is in flight on that buffer
  const int o = prog.n_out;
  const int n_ops = prog.n_ops;
  // stage state: bit 0 = bias buffer of the current stage (toggles every stage, output stage included),
  // bit 1 = weight buffer of the current GEMM layer (toggles every GEMM layer)
  uint32_t st = 0;
  uint32_t w_par = 0;   // (thread 0) parity of the next completion of w_bar[0], w_bar[1]; kept out of `st` so that
                        // `st`, and with it the weight buffer's shared-memory address, stays warp-uniform: the
                        // B descriptors of the MMAs are then built on the uniform datapath instead of going
                        // through a per-MMA register-to-uniform loop (13 -> 5 instructions per MMA)
#define ST_B (st & 1u)
#define ST_W ((st >> 1) & 1u)
  if (t == 0 && (int)blockIdx.x < n_tiles) {
    mbar_arrive_expect_tx(&w_bar[0], kWTileBytes);
    bulk_copy_g2s(smem + TcSmem::kW, packed + prog.ops[1].tc_off, kWTileBytes, &w_bar[0]);
  }

  // TMEM is managed as eight 64-column blocks.  blk[j] = block holding channel j of the layer being
  // consumed; channel j of the layer being produced goes to a free block.  Every thread tracks the
  // same (deterministic) bookkeeping.
  uint32_t blk_p = 0, nblk_p = 0;   // 4 bits per channel (indexed with a run-time channel number)
#define BLK(j) ((blk_p >> (4 * (j))) & 7u)
  uint32_t free_mask = 0xFFu;

#if MPINN_TC_TIMING
  long long dbg_last = clock64();
#endif
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = (int64_t)tile * 128 + p;
    const bool valid = pt < n;
    float xv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xv[i] = valid ? __ldg(x + pt * D + i) : 0.0f;
    TCF_TIME(15);
    constexpr bool keep = KEEP;
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
      const bool add_skip = HAS_SKIP && (S >= 1) && (sv.flags & MPINN_FLAG_ADD_SKIP);   // z^S includes the block input
      const bool save_skip = HAS_SKIP && (op.flags & MPINN_FLAG_SAVE_SKIP);            // a^S is a block input
      if (t == 0) {   // every GEMM that read the other buffer has completed (waited at the end of its layer)
        const bool more = (L + 1 < n_ops - 1) || (tile + (int)gridDim.x < n_tiles);
        if (more) {
          const DevOp& nx = prog.ops[(L + 1 < n_ops - 1) ? L + 1 : 1];
          const int nbuf = ST_W ^ 1;
          mbar_arrive_expect_tx(&w_bar[nbuf], kWTileBytes);
          bulk_copy_g2s(smem + TcSmem::kW + nbuf * kWTileBytes, packed + nx.tc_off, kWTileBytes, &w_bar[nbuf]);
        }
      }
      if (t < kTcW) bias_s[(ST_B ^ 1) * kTcW + t] = __ldg(params + op.b_off + t);   // read one stage later
      TCF_TIME(10);   // layer start: weight prefetch issue, bias staging
      const int act_src = HAS_FOURIER ? act_of(sv, f0) : sv.act;
      float s1[kFpt], s2[kFpt], zi_keep[kFpt];
      float* zst = tstash + sv.stash_plane * 128;
      // One channel of the source layer.  The channel index is a compile-time constant (the J calls below
      // are J specialised copies): which of the three activation-jet formulas applies, the TMEM block, the
      // stash plane and the A buffer are all resolved by the compiler instead of per-element branches.
      // (With seven channels that many copies thrash the instruction cache: then only the channel KIND is the
      // compile-time constant -- three copies -- and the direction is a run-time value.  `spec_c` = IC<idx> or
      // IC<-1 - kind>.)
      auto channel = [&](auto spec_c, const int dir_rt) {
        constexpr int spec = decltype(spec_c)::value;
        constexpr int idx_ct = spec >= 0 ? spec : 0;
        // production order: value, then (first, second) derivative per direction; in Laplace mode value, the D
        // first derivatives, then L (which needs all of them)
        constexpr int j_ct = (idx_ct == 0) ? 0
                             : (ORDER == 2 ? (((idx_ct - 1) & 1) ? 1 + D + ((idx_ct - 1) >> 1) : 1 + ((idx_ct - 1) >> 1))
                                           : idx_ct);
        constexpr int kind = spec >= 0 ? (j_ct == 0 ? 0 : (j_ct <= D ? 1 : 2)) : -1 - spec;
        constexpr bool is_first = (kind == 1);
        constexpr bool is_second = (kind == 2);
        const int dir = spec >= 0 ? (is_first ? j_ct - 1 : (is_second ? j_ct - 1 - D : 0)) : dir_rt;
        const int j = spec >= 0 ? j_ct : (kind == 0 ? 0 : (is_first ? 1 + dir : 1 + D + dir));
        const int idx = spec >= 0 ? idx_ct : (kind == 0 ? 0 : (ORDER == 2 ? 1 + 2 * dir + (kind - 1) : 1 + dir));
        const int buf = idx & 1;
        uint8_t* at = smem + (buf ? TcSmem::kA1 : TcSmem::kA0);
#pragma unroll
        for (int c = 0; c < kFpt / 8; ++c) {
          float z[8];
          if (S == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int f = f0 + c * 8 + e;
              if (kind == 0) {
                float v = bias_s[ST_B * kTcW + f];
#pragma unroll
                for (int i = 0; i < D; ++i)
                  v = fmaf(HAS_FOURIER ? first_w(sv, params, f, i, D) : __ldg(params + sv.w_off + f * D + i), xv[i], v);
                z[e] = v;
              } else if (is_first) {
                z[e] = HAS_FOURIER ? first_w(sv, params, f, dir, D) : __ldg(params + sv.w_off + f * D + dir);
              } else {
                z[e] = 0.0f;
              }
            }
          } else {
            tmem_ld8(tmem + lane_base + BLK(j) * kTcW + f0 + c * 8, z);
            if (kind == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) z[e] += bias_s[ST_B * kTcW + f0 + c * 8 + e];
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
            if (keep) zst[(j * kTcW + f0 + ff) * 128] = z[e];          // stash for the reverse pass
            if (kind == 0) {
              float c0, c1, c2, c3;
              act_coeffs<false>(act_src, z[e], c0, c1, c2, c3);
              s1[ff] = c1;
              s2[ff] = c2;
              a[e] = c0;
            } else if (is_first) {
              if (LAP) {   // zi_keep accumulates q = sum_i c_i z_i^2
                const float cz = prog.lap_c[dir] * z[e];
                zi_keep[ff] = (dir == 0) ? cz * z[e] : fmaf(cz, z[e], zi_keep[ff]);
              } else {
                zi_keep[ff] = z[e];
              }
              a[e] = s1[ff] * z[e];
            } else if (LAP) {
              a[e] = fmaf(s2[ff], zi_keep[ff], s1[ff] * z[e]);
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
            TCF_TIME(11);   // source jets + activation math + stash stores + split, first chunk
            if (buf == 0 && pending0) { mbar_wait(&mma_bar[0], bar_phase0); bar_phase0 ^= 1; pending0 = 0; }
            if (buf == 1 && pending1) { mbar_wait(&mma_bar[1], bar_phase1); bar_phase1 ^= 1; pending1 = 0; }
            TCF_TIME(12);   // wait for the A buffer
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t off = tile_kb + sw128_offset(p, col0 + c * 8 + h * 4);
            *reinterpret_cast<float4*>(at + off) = make_float4(hi[h * 4], hi[h * 4 + 1], hi[h * 4 + 2], hi[h * 4 + 3]);
            *reinterpret_cast<float4*>(at + 2 * kRegionBytes + off) =
                make_float4(lo[h * 4], lo[h * 4 + 1], lo[h * 4 + 2], lo[h * 4 + 3]);
          }
        }
        TCF_TIME(13);     // rest of the channel's staging (second chunk, tile stores)
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        TCF_TIME(14);     // channel barrier
        // blocks of the source layer that nobody needs any more
        if (S >= 1) {
          if (kind == 0) free_mask |= 1u << BLK(0);
          else if (ORDER == 1 || LAP) free_mask |= 1u << BLK(j);
          else if (is_second) free_mask |= (1u << BLK(j)) | (1u << BLK(1 + dir));
        }
        const int nb = __ffs(free_mask) - 1;
        free_mask &= ~(1u << nb);
        nblk_p |= (uint32_t)nb << (4 * j);
        if (t == 0) {
          if (kind == 0) {   // this layer's weight operand has landed
            mbar_wait(&w_bar[ST_W], (w_par >> ST_W) & 1u);
            w_par ^= 1u << ST_W;
          }
          tc_fence_after_sync();
          issue_gemm_kmajor(tmem + nb * kTcW, smem_u32(at), smem_u32(smem + TcSmem::kW + ST_W * kWTileBytes));
          mma_commit(&mma_bar[buf]);
        }
        TCF_TIME(6);      // MMA issue (thread 0 only)
        if (buf == 0) pending0 = 1; else pending1 = 1;
      };
      if constexpr (J <= 5) {
        channel(IC<0>{}, 0);
        if constexpr (J > 1) channel(IC<1>{}, 0);
        if constexpr (J > 2) channel(IC<2>{}, 0);
        if constexpr (J > 3) channel(IC<3>{}, 0);
        if constexpr (J > 4) channel(IC<4>{}, 0);
      } else {
        channel(IC<-1>{}, 0);
#pragma unroll 1
        for (int dir = 0; dir < D; ++dir) {
          channel(IC<-2>{}, dir);
          if (ORDER == 2) channel(IC<-3>{}, dir);
        }
      }
      if (pending0) { mbar_wait(&mma_bar[0], bar_phase0); bar_phase0 ^= 1; pending0 = 0; }
      if (pending1) { mbar_wait(&mma_bar[1], bar_phase1); bar_phase1 ^= 1; pending1 = 0; }
      TCF_TIME(7);        // layer-end GEMM wait
      tc_fence_after_sync();
      blk_p = nblk_p;
      nblk_p = 0;
      st ^= 3u;
    }

    // ---- output op: LINEAR(64 -> o) on the FMA pipe, fed from the last hidden layer's TMEM blocks
    {
      const DevOp& op = prog.ops[n_ops - 1];
      const DevOp& sv = prog.ops[n_ops - 2];
      const float* W = params + op.w_off;
      float* zst = tstash + sv.stash_plane * 128;
      if (t < kTcW)   // first-op bias for this CTA's next tile
        bias_s[(ST_B ^ 1) * kTcW + t] =
            (HAS_FOURIER && prog.ops[0].kind == MPINN_OP_FOURIER) ? 0.0f : __ldg(params + prog.ops[0].b_off + t);
      float acc[J][4];
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.0f;
#pragma unroll
      for (int c = 0; c < kFpt / 8; ++c) {
        float zj[J][8];
#pragma unroll
        for (int j = 0; j < J; ++j) tmem_ld8(tmem + lane_base + BLK(j) * kTcW + f0 + c * 8, zj[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          float z[J];
#pragma unroll
          for (int j = 0; j < J; ++j) z[j] = zj[j][e];
          z[0] += bias_s[ST_B * kTcW + f];
          if (HAS_SKIP && (sv.flags & MPINN_FLAG_ADD_SKIP)) {
#pragma unroll
            for (int j = 0; j < J; ++j) z[j] += skip_s[(j * kTcW + f) * 128];
          }
#pragma unroll
          if (keep) {
#pragma unroll
            for (int j = 0; j < J; ++j) zst[(j * kTcW + f) * 128] = z[j];
          }
          act_jets<D, ORDER>(sv.act, z, prog.lap_c);   // (the last hidden layer is never the Fourier op)
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
      for (int j = 0; j < J; ++j) free_mask |= 1u << BLK(j);
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
      st ^= 1u;
      TCF_TIME(9);        // output op
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
#undef BLK
#undef ST_B
#undef ST_W
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_forward_supported(const DevProg& prog, int d, int order) {
  if (order > 3) return false;   // mixed second derivatives run on the FP32-FMA engine
  const int J = jet_count(d, order);
  if (prog.has_skip && prog.ops[0].kind == MPINN_OP_FOURIER) return false;   // no combined kernel variant
  if (J * kTcW > 448) return false;
  if (prog.n_ops < 3) return false;
  if (prog.n_out > 4) return false;
  const DevOp& first = prog.ops[0];
  if ((first.kind != MPINN_OP_LINEAR && first.kind != MPINN_OP_FOURIER) || first.in_dim != d ||
      first.out_dim != kTcW || first.flags)
    return false;
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

template <int D, int ORDER, int FEAT, bool KEEP = true>
static cudaError_t launch_fwd_tc_t(const DevProg& prog, const float* params, const float* packed,
                                   const float* x, int64_t n, float* out, float* stash, int64_t n_pad,
                                   int sm_count, cudaStream_t stream) {
  auto kern = jet_fwd_tc<D, ORDER, FEAT, KEEP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::kTotal);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n + 127) / 128);
  int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid > 148) grid = 148;   // the skip scratch holds 148 CTA regions
  kern<<<grid, kTcThreads, TcSmem::kTotal, stream>>>(prog, params, packed, x, n, out, stash, n_pad, n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_fwd_tc(const DevProg& prog, const float* params, const float* packed, const float* x,
                          int64_t n, int d, int order, float* out, float* stash, int64_t n_pad,
                          int sm_count, cudaStream_t stream) {
  const int feat = tc_feat(prog);
  if (stash == nullptr) {   // value-only pass: order 0, no skip connections (their scratch lives behind the stash)
    if (order != 0 || prog.has_skip) return cudaErrorInvalidValue;
#define TC_FWD0(D_)                                                                                             \
  if (d == D_) {                                                                                                \
    if (feat == 0) return launch_fwd_tc_t<D_, 0, 0, false>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream); \
    return launch_fwd_tc_t<D_, 0, 2, false>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream);    \
  }
    TC_FWD0(1); TC_FWD0(2); TC_FWD0(3);
#undef TC_FWD0
    return cudaErrorInvalidValue;
  }
#define TC_FWD(D_, O_)                                                                                      \
  if (d == D_ && order == O_) {                                                                             \
    if (feat == 0) return launch_fwd_tc_t<D_, O_, 0>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream); \
    if (feat == 1) return launch_fwd_tc_t<D_, O_, 1>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream); \
    return launch_fwd_tc_t<D_, O_, 2>(prog, params, packed, x, n, out, stash, n_pad, sm_count, stream);     \
  }
  TC_FWD(1, 0); TC_FWD(1, 1); TC_FWD(1, 2);
  TC_FWD(2, 0); TC_FWD(2, 1); TC_FWD(2, 2);
  TC_FWD(3, 0); TC_FWD(3, 1); TC_FWD(3, 2);
  TC_FWD(1, 3); TC_FWD(2, 3); TC_FWD(3, 3);
#undef TC_FWD
  return cudaErrorInvalidValue;
}

}  // namespace mpinn
