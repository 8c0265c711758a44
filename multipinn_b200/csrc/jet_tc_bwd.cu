// jet_tc_bwd.cu -- reverse pass of the tcgen05 engine (see jet_tc.cu for the forward and the design notes).
#include "jet_tc_common.cuh"

namespace mpinn {

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
};

// The two GEMMs of a phase, out of line (one copy instead of one per specialised phase) and with the operand
// tiles addressed from the warp-uniform shared-memory base inside the function rather than through arguments:
// descriptor arithmetic then stays on the uniform datapath instead of a register -> uniform-register move loop
// around every MMA (21 -> ~8 instructions per MMA on the issuing threads, which the other warps wait for).
extern __shared__ __align__(1024) uint8_t mpinn_bwd_smem_raw[];
static __device__ __noinline__ void issue_input_gradient_gemm(uint32_t d_tmem) {
  uint8_t* smem = align_smem_1024(mpinn_bwd_smem_raw);
  issue_bf16x3_kmajor(d_tmem, smem_u32(smem + TcBwdSmem::kZ), smem_u32(smem + TcBwdSmem::kW));
}
template <bool WIDE>   // WIDE: 128 accumulator columns, N-stacked H pieces (see issue_bf16x3_mnmajor_wide)
static __device__ __noinline__ void issue_weight_gradient_gemm(uint32_t d_tmem, uint32_t accumulate) {
  uint8_t* smem = align_smem_1024(mpinn_bwd_smem_raw);
  if (WIDE) issue_bf16x3_mnmajor_wide(d_tmem, smem_u32(smem + TcBwdSmem::kZ), smem_u32(smem + TcBwdSmem::kH), accumulate);
  else issue_bf16x3_mnmajor(d_tmem, smem_u32(smem + TcBwdSmem::kZ), smem_u32(smem + TcBwdSmem::kH), accumulate);
}

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

template <int D, int ORDER, int FEAT>
__global__ void __launch_bounds__(kTcThreads, 1)
jet_bwd_tc(const __grid_constant__ DevProg prog, const float* __restrict__ params,
           const float* __restrict__ packed, const float* __restrict__ x, int64_t n,
           const float* __restrict__ dout, const float* __restrict__ stash, int64_t n_pad,
           float* __restrict__ part_all, int64_t n_params, float* __restrict__ skipbar_all,
           float* __restrict__ dx, int n_tiles) {
  constexpr int J = jet_count(D, ORDER);
  constexpr bool LAP = (ORDER == 3);   // one second-order channel L = sum_i c_i d2/dx_i2 at jet index 1 + D
  constexpr bool HAS_SKIP = (FEAT == 1), HAS_FOURIER = (FEAT == 2);
  // tensor-memory map: [0, 64 J) abar channels | s0 cache | s3 cache | dW accumulator.  With at most four channels
  // the accumulator gets 128 columns and the weight-gradient GEMM runs with N-stacked operand pieces (DW_WIDE).
  constexpr bool DW_WIDE = (J <= 4);
  constexpr uint32_t kDwCol = DW_WIDE ? 384 : 448;  // TMEM columns of the dW accumulator
  constexpr uint32_t kT0Col = DW_WIDE ? 256 : 320;  // TMEM columns caching the activation value of the layer below (J <= 5)
  constexpr uint32_t kT3Col = kT0Col + 64;          // TMEM columns caching its third derivative
  // TCACHE: five channels leave 128 TMEM columns free, used to cache s0/s3 of the layer below (s1/s2 in
  // shared memory).  With seven channels TMEM is full: shared memory then holds the two basis values of
  // the activation instead and s0..s3 are re-derived from them (a few FMAs + one z0 reload) where needed.
  constexpr bool TCACHE = (J * kTcW <= (int)kT0Col);
  static_assert(kT3Col + 64 <= kDwCol && kDwCol + (DW_WIDE ? 128 : 64) <= 512, "tensor-memory map");
  constexpr int NC = kFpt / 8;      // 8-feature chunks per thread
  uint8_t* smem = align_smem_1024(mpinn_bwd_smem_raw);
  uint8_t* z_tile = smem + TcBwdSmem::kZ;
  uint8_t* h_tile = smem + TcBwdSmem::kH;
  uint8_t* w_tile = smem + TcBwdSmem::kW;
  float* cache = reinterpret_cast<float*>(smem + TcBwdSmem::kCache);  // [2][64][128]
  float* red_s = reinterpret_cast<float*>(smem + TcBwdSmem::kRed);    // [64][8]
  float* bsum_s = red_s + 64 * 8;                                     // [16 warps][16] bias-gradient partials
  __shared__ uint64_t mma_bar;
  __shared__ uint64_t w_bar;      // the TMA engine drops each layer's W^T operand image into w_tile
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
    mbar_init(&w_bar, 1);
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
  // bytes [kWTileBytes, kWTileBytes + kBfWTile) of a layer's packed block are its bf16x3 W^T operand image
  auto fetch_w = [&](const DevOp& wop) {
    mbar_arrive_expect_tx(&w_bar, kBfWTile);
    bulk_copy_g2s(w_tile, reinterpret_cast<const uint8_t*>(packed + wop.tc_off) + kWTileBytes, kBfWTile, &w_bar);
  };
  if (t == 0 && (int)blockIdx.x < n_tiles && n_ops >= 3) fetch_w(prog.ops[n_ops - 2]);

  // activation coefficients s0..s3 (value and first three derivatives) of layer `op` at this thread's
  // (point, features), evaluated once: s1/s2 go to shared memory, s0/s3 to spare TMEM columns
  auto fill_cache = [&](const DevOp& op, const float* tstash) {
    const float* z0p = tstash + op.stash_plane * 128;
    const int act = HAS_FOURIER ? act_of(op, f0) : op.act;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float z[8], c0[8], c3[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = __ldg(z0p + (f0 + c * 8 + e) * 128);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = f0 + c * 8 + e;
        float c1, c2;
        act_coeffs<(ORDER >= 2)>(act, z[e], c0[e], c1, c2, c3[e]);
        cache[f * 128 + p] = c1;
        cache[(64 + f) * 128 + p] = c2;
      }
      if (TCACHE) {
        tmem_st8(tmem + lane_base + kT0Col + f0 + c * 8, c0);
        tmem_st8(tmem + lane_base + kT3Col + f0 + c * 8, c3);
      }
    }
  };
  // With seven channels tensor memory has no room for s0 / s3: they are evaluated again from the stashed
  // pre-activation where needed (s3 once per layer, s0 in the value phase); s1 / s2 always come from the cache.
  auto s0_s3_at = [&](const DevOp& op, int act, const float* tstash, int f, float& c0, float& c3) {
    const float z0 = __ldg(tstash + (op.stash_plane + f) * 128);
    float c1, c2;
    act_coeffs<(ORDER >= 2)>(act, z0, c0, c1, c2, c3);
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
        if (TCACHE) tmem_ld8(tmem + lane_base + kT0Col + f0 + c * 8, t0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          const float s1 = cache[f * 128 + p], s2 = cache[(64 + f) * 128 + p];
          if (!TCACHE) {
            float s3;
            s0_s3_at(pv, HAS_FOURIER ? act_of(pv, f0) : pv.act, tstash, f, t0[e], s3);
          }
          float a[J];
          a[0] = t0[e];
          if (ORDER >= 1) {
            float q = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              const float zi = __ldg(zp + ((1 + i) * kTcW + f) * 128);
              a[1 + i] = s1 * zi;
              if (ORDER == 2) a[1 + D + i] = fmaf(s2 * zi, zi, s1 * __ldg(zp + ((1 + D + i) * kTcW + f) * 128));
              if (LAP) q = fmaf(prog.lap_c[i] * zi, zi, q);
            }
            if (LAP) a[1 + D] = fmaf(s2, q, s1 * __ldg(zp + ((1 + D) * kTcW + f) * 128));
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
      const bool store_sb = HAS_SKIP && (op.flags & MPINN_FLAG_ADD_SKIP);              // zbar^l also flows to the block input
      const bool add_sb = HAS_SKIP && (prog.ops[l + 1].flags & MPINN_FLAG_SAVE_SKIP);  // abar^l gets the skip branch from above
      // first three derivatives of layer l's activation for this thread's features, from the cached basis
      float s1[kFpt], s2[kFpt], s3[kFpt], acc0[kFpt];
      const int act_l = HAS_FOURIER ? act_of(op, f0) : op.act;
      const int act_m = HAS_FOURIER ? act_of(pv, f0) : pv.act;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float v3[8];
        if (TCACHE) tmem_ld8(tmem + lane_base + kT3Col + f0 + c * 8, v3);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ff = c * 8 + e, f = f0 + ff;
          s1[ff] = cache[f * 128 + p];
          s2[ff] = cache[(64 + f) * 128 + p];
          if (TCACHE) {
            s3[ff] = v3[e];
          } else {
            float s0;
            s0_s3_at(op, act_l, tstash, f, s0, s3[ff]);
          }
          acc0[ff] = 0.0f;
        }
      }
      // (no barrier needed here: the coefficient caches are thread-private and red_s is not used by the GEMM layers)
      fill_cache(pv, tstash);          // now the cache holds layer l-1

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
      // One phase = one channel.  The phase index is a compile-time constant (J specialised copies), so the
      // channel kind, TMEM columns and stash planes are resolved by the compiler, not by per-element branches.
      // Up to five channels every phase gets its own specialised copy (the phase index is the compile-time
      // constant).  With seven channels that much code thrashes the instruction cache, so only the channel KIND
      // (value / first / second derivative) is a compile-time constant and the direction is a run-time value:
      // three copies.  `spec_c` = IC<step> (>= 0) or IC<-1 - kind>.
      auto phase = [&](auto spec_c, const int dir_rt, const bool first_rt) {
        constexpr int spec = decltype(spec_c)::value;
        constexpr int step = spec >= 0 ? spec : 0;
        constexpr int jet_ct = (step == J - 1) ? 0
                               : (ORDER == 2 ? ((step & 1) ? 1 + D + (step >> 1) : 1 + (step >> 1)) : 1 + step);
        constexpr int kind = spec >= 0 ? (jet_ct == 0 ? 0 : (jet_ct <= D ? 1 : 2)) : -1 - spec;
        constexpr bool is_first = (kind == 1);
        constexpr bool is_second = (kind == 2);
        constexpr bool two = is_first && ORDER >= 2;   // this phase also consumes the second-order abar of its direction
        constexpr bool lap2 = LAP && is_second;         // the L channel's own phase
        const int dir = spec >= 0 ? (is_first ? jet_ct - 1 : (is_second ? jet_ct - 1 - D : 0)) : dir_rt;
        const int jet = spec >= 0 ? jet_ct : (kind == 0 ? 0 : (is_first ? 1 + dir : 1 + D + dir));
        const int sec = LAP ? 1 + D : 1 + D + dir;      // jet index of the second-order channel paired with `dir`
        const float lcd = LAP ? prog.lap_c[dir] : 1.0f;
        const bool first_phase = spec >= 0 ? (step == 0) : first_rt;

        float zb0all[kFpt];                      // value-channel zbar of this phase (bias gradient)
        float zi_c[NC][8], zii_c[NC][8];
        uint32_t zq[NC][3][4], hq[NC][3][4];   // this phase's bf16x3 operand pieces, stored after the GEMM wait
        // The GEMMs of the previous phase read both operand tiles but none of the TMEM columns this phase
        // touches, so they are waited for as late as possible: with ALL of this phase's math done and its bf16x3
        // pieces in registers (kLateWait), or -- for the residual-block variant, which has the skip gradient's
        // registers live as well and would spill -- after the first chunk's math.
        constexpr bool kLateWait = (FEAT != 1);
        auto wait_prev_gemms = [&]() {
          TC_TIME(2);   // loads + math
          if (pending) {
            mbar_wait(&mma_bar, bar_phase & 1);
            bar_phase ^= 1;
            pending = 0;
          }
          TC_TIME(3);   // waiting for the previous phase's GEMMs
        };
        auto store_chunk = [&](int c) {
          const uint32_t off = sw128_offset_bf16(p, f0 + c * 8);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            *reinterpret_cast<uint4*>(z_tile + q * kBfRegion + off) = make_uint4(zq[c][q][0], zq[c][q][1], zq[c][q][2], zq[c][q][3]);
            *reinterpret_cast<uint4*>(h_tile + q * kBfRegion + off) = make_uint4(hq[c][q][0], hq[c][q][1], hq[c][q][2], hq[c][q][3]);
          }
        };
        auto request_zl = [&](int c) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int f = f0 + c * 8 + e;
            zi_c[c][e] = (is_first || lap2) ? __ldg(zl + (jet * kTcW + f) * 128) : 0.0f;   // z_i  (L phase: z_L)
            zii_c[c][e] = (two && !LAP) ? __ldg(zl + ((1 + D + dir) * kTcW + f) * 128) : 0.0f;
          }
        };
        request_zl(0);
#pragma unroll
        for (int c = 0; c < NC; ++c) {          // 8 features per chunk = one 16-byte bf16 store per piece
          const uint32_t off = sw128_offset_bf16(p, f0 + c * 8);
          // the layer-below jets needed by the second half of this chunk are requested first, so that their
          // latency overlaps the zbar math (nothing between here and their use can be reordered by the
          // compiler across the barrier / tensor-memory instructions)
          float yi[8], yii[8];
          if (kind != 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int f = f0 + c * 8 + e;
              if (lap2) {   // yi <- q = sum_i c_i y_i^2 over all directions
                float q = 0.0f;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                  const float y = __ldg(zm + ((1 + i) * kTcW + f) * 128);
                  q = fmaf(prog.lap_c[i] * y, y, q);
                }
                yi[e] = q;
              } else {
                yi[e] = __ldg(zm + ((1 + dir) * kTcW + f) * 128);
              }
              yii[e] = is_second ? __ldg(zm + (jet * kTcW + f) * 128) : 0.0f;
            }
          }
          float (&zi)[8] = zi_c[c];           // (this layer's jets: requested one chunk ahead, see below)
          float (&zii)[8] = zii_c[c];
          uint32_t (&zq1)[4] = zq[c][0], (&zq2)[4] = zq[c][1], (&zq3)[4] = zq[c][2];   // bf16x3 pieces of zbar
          // ---------------- zbar of layer l ----------------
          {
            float ab[8], ab2[8], zb[8];
            tmem_ld8(tmem + lane_base + jet * kTcW + f0 + c * 8, ab);
            if (two) tmem_ld8(tmem + lane_base + sec * kTcW + f0 + c * 8, ab2);
            if (add_sb) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int f = f0 + c * 8 + e;
                ab[e] += sb[(jet * kTcW + f) * 128];
                if (two) ab2[e] += sb[(sec * kTcW + f) * 128];
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ff = c * 8 + e;
              if (kind == 0) {
                zb[e] = fmaf(ab[e], s1[ff], acc0[ff]);
              } else if (is_first) {
                float v = ab[e] * s1[ff];
                float a0 = fmaf(ab[e] * s2[ff], zi[e], acc0[ff]);
                if (ORDER == 2) {
                  v = fmaf(2.0f * ab2[e] * s2[ff], zi[e], v);
                  a0 = fmaf(ab2[e], fmaf(s3[ff] * zi[e], zi[e], s2[ff] * zii[e]), a0);
                }
                if (LAP) {   // abar_L reaches z_i through q = sum_i c_i z_i^2
                  const float w = lcd * ab2[e];
                  v = fmaf(2.0f * w * s2[ff], zi[e], v);
                  a0 = fmaf(w * s3[ff] * zi[e], zi[e], a0);
                }
                acc0[ff] = a0;
                zb[e] = v;
              } else {
                zb[e] = ab[e] * s1[ff];
                if (lap2) acc0[ff] = fmaf(ab[e] * s2[ff], zi[e], acc0[ff]);   // s'' z_L term of the value channel
              }
            }
            if (kind == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) zb0all[c * 8 + e] = zb[e];
            }
            if (store_sb) {
#pragma unroll
              for (int e = 0; e < 8; ++e) sb[(jet * kTcW + f0 + c * 8 + e) * 128] = zb[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x3(zb[2 * e], zb[2 * e + 1], zq1[e], zq2[e], zq3[e]);
          }
          // ---------------- activation jet channel of layer l-1 ----------------
          {
            float av[8];
            if (kind == 0 && TCACHE) {
              tmem_ld8(tmem + lane_base + kT0Col + f0 + c * 8, av);
            } else if (kind == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float t3;
                s0_s3_at(pv, act_m, tstash, f0 + c * 8 + e, av[e], t3);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int f = f0 + c * 8 + e;
                const float t1 = cache[f * 128 + p], t2 = cache[(64 + f) * 128 + p];
                av[e] = is_first ? t1 * yi[e] : (lap2 ? fmaf(t2, yi[e], t1 * yii[e]) : fmaf(t2 * yi[e], yi[e], t1 * yii[e]));
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x3(av[2 * e], av[2 * e + 1], hq[c][0][e], hq[c][1][e], hq[c][2][e]);
            if (c + 1 < NC) request_zl(c + 1);
          }
          if (!kLateWait) {   // residual-block programs: store chunk by chunk (fewer live registers)
            if (c == 0) wait_prev_gemms();
            store_chunk(c);
          }
        }
        if (kLateWait) {
          wait_prev_gemms();
#pragma unroll
          for (int c = 0; c < NC; ++c) store_chunk(c);
        }
        if (kind == 0) {
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
          if (first_phase) {     // this layer's W^T operand has landed (bit 1 of bar_phase: parity of w_bar)
            mbar_wait(&w_bar, (bar_phase >> 1) & 1);
            bar_phase ^= 2;
          }
          tc_fence_after_sync();
          issue_input_gradient_gemm(tmem + jet * kTcW);
          mma_commit(&mma_bar);
        } else if (t == 32) {    // ... one lane of warp 1 the weight-gradient GEMM (independent accumulators)
          tc_fence_after_sync();
          issue_weight_gradient_gemm<DW_WIDE>(tmem + kDwCol, dw_acc);
          mma_commit(&mma_bar);
        }
        TC_TIME(6);     // MMA issue (thread 0 only)
        dw_acc = 1;
        pending = 1;
      };
      if constexpr (J <= 5) {
        phase(IC<0>{}, 0, false);
        if constexpr (J > 1) phase(IC<1>{}, 0, false);
        if constexpr (J > 2) phase(IC<2>{}, 0, false);
        if constexpr (J > 3) phase(IC<3>{}, 0, false);
        if constexpr (J > 4) phase(IC<4>{}, 0, false);
      } else {   // (first, second) derivative per direction, value channel last
#pragma unroll 1
        for (int dir = 0; dir < D; ++dir) {
          phase(IC<-2>{}, dir, dir == 0);
          if (ORDER == 2) phase(IC<-3>{}, dir, false);
        }
        phase(IC<-1>{}, 0, false);
      }
      mbar_wait(&mma_bar, bar_phase & 1);
      TC_TIME(7);       // final GEMM wait of the layer
      bar_phase ^= 1;
      pending = 0;
      tc_fence_after_sync();
      if (t == 0) {     // w_tile is free again: fetch the next layer's (or the next tile's top layer's) operand
        if (l > 1) fetch_w(prog.ops[l - 1]);
        else if (tile + (int)gridDim.x < n_tiles) fetch_w(prog.ops[n_ops - 2]);
      }
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
          if (DW_WIDE) {   // second column block: the H2 products of the same (out, in) features
            float v2[8];
            tmem_ld8(tmem + lane_base + kDwCol + 64 + f0 + c * 8, v2);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += v2[e];
          }
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
      // A block barrier is needed here only with a single channel: then the next layer's one phase writes the
      // bias partials (bsum_s) that warps 0-1 may still be reading above.  With more channels at least one phase
      // barrier lies in between; the dW accumulator columns read above are not overwritten before the next
      // layer's first phase barrier either, and everything else a layer touches is thread-private.
      if (J == 1) {
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
      }
      TC_TIME(8);       // dW flush
    }

    // =========================== op 0 (d -> 64), FMA pipe ===========================
    if (!HAS_FOURIER) {   // (a frozen Fourier encoding has no gradient)
      const DevOp& op = prog.ops[0];
      const float* z0p = tstash + op.stash_plane * 128;
      for (int i = t; i < 64 * 8; i += kTcThreads) red_s[i] = 0.0f;
      __syncthreads();
      const bool add_sb0 = HAS_SKIP && (prog.ops[1].flags & MPINN_FLAG_SAVE_SKIP);
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
        if (TCACHE) tmem_ld8(tmem + lane_base + kT3Col + f0 + c * 8, s3);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int f = f0 + c * 8 + e;
          s1[e] = cache[f * 128 + p];
          s2[e] = cache[(64 + f) * 128 + p];
          if (!TCACHE) {
            float s0;
            s0_s3_at(op, op.act, tstash, f, s0, s3[e]);
          }
          zb0[c * 8 + e] = ab0[e] * s1[e];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float abi[8], abii[8];
          const int sec0 = LAP ? 1 + D : 1 + D + i;
          if (ORDER >= 1) tmem_ld8(tmem + lane_base + (1 + i) * kTcW + f0 + c * 8, abi);
          if (ORDER >= 2) tmem_ld8(tmem + lane_base + sec0 * kTcW + f0 + c * 8, abii);
          if (add_sb0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (ORDER >= 1) abi[e] += sb[((1 + i) * kTcW + f0 + c * 8 + e) * 128];
              if (ORDER >= 2) abii[e] += sb[(sec0 * kTcW + f0 + c * 8 + e) * 128];
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
              if (LAP) {   // (z_L of the first layer is identically zero: z = W x + b)
                const float w = prog.lap_c[i] * abii[e];
                zbi = fmaf(2.0f * w * s2[e], zi, zbi);
                zb0[c * 8 + e] = fmaf(w * s3[e] * zi, zi, zb0[c * 8 + e]);
              }
            }
            zbd[i][c * 8 + e] = zbi;          // d z_i / d W[f][i] = 1
          }
        }
      }
      // gradient w.r.t. the points: x enters the network only through the value channel of op 0
      // (z = W x + b; dz_i/dx = 0), so dL/dx_i = sum_f W[f][i] zbar_0[f]; the four feature quarters of a point
      // meet in global memory (dx is zeroed by the caller of the kernel)
      if (dx != nullptr && valid) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float s = 0.0f;
#pragma unroll
          for (int e = 0; e < kFpt; ++e) s = fmaf(__ldg(params + op.w_off + (f0 + e) * D + i), zb0[e], s);
          atomicAdd(dx + pt * D + i, s);
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
bool tc_backward_supported(const DevProg& prog, int d, int order) {
  return tc_forward_supported(prog, d, order);
}

// dparams[i] += sum_c part[c][i]  (same fold as the FFMA engine)
__global__ void fold_partials_tc(const float* __restrict__ part, int n_slabs, int64_t n_params,
                                 float* __restrict__ dparams, const double* __restrict__ bsum, int bsum_off, int n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  float s = 0.0f;
  const int64_t slab = (n_params + 3) & ~int64_t(3);
  if (!take_bsum(bsum, bsum_off, n_out, i, &s))
    for (int c = 0; c < n_slabs; ++c) s += part[(int64_t)c * slab + i];
  dparams[i] += s;
}

template <int D, int ORDER, int FEAT>
static cudaError_t launch_bwd_tc_t(const DevProg& prog, const float* params, const float* packed,
                                   const float* x, int64_t n, const float* dout, const float* stash,
                                   int64_t n_pad, float* part, int max_slabs, int64_t n_params,
                                   float* dparams, float* dx, float* skip_scratch, int sm_count,
                                   cudaStream_t stream) {
  auto kern = jet_bwd_tc<D, ORDER, FEAT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwdSmem::kTotal);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n + 127) / 128);
  int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid > max_slabs - 1) grid = max_slabs - 1;   // slabs are padded to 16 bytes: keep one in reserve
  if (grid < 1) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(part) & 15) != 0) return cudaErrorMisalignedAddress;
  if (grid > 148) grid = 148;   // the skip scratch holds 148 CTA regions
  kern<<<grid, kTcThreads, TcBwdSmem::kTotal, stream>>>(prog, params, packed, x, n, dout, stash, n_pad, part,
                                                         n_params, skip_scratch, dx, n_tiles);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fold_partials_tc<<<(unsigned)((n_params + 255) / 256), 256, 0, stream>>>(part, grid, n_params, dparams, prog.bsum,
                                                                           prog.bsum_off, prog.n_out);
  return cudaGetLastError();
}

cudaError_t launch_bwd_tc(const DevProg& prog, const float* params, const float* packed, const float* x,
                          int64_t n, int d, int order, const float* dout, const float* stash, int64_t n_pad,
                          float* part, int max_slabs, int64_t n_params, float* dparams, float* dx,
                          float* skip_scratch, int sm_count, cudaStream_t stream) {
  const int feat = tc_feat(prog);
  if (dx != nullptr) {
    if (feat == 2) return cudaErrorNotSupported;   // (checked with a message at the C ABI)
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)n * d * sizeof(float), stream);
    if (e != cudaSuccess) return e;
  }
#define TC_BWD_ARGS (prog, params, packed, x, n, dout, stash, n_pad, part, max_slabs, n_params, dparams, dx, skip_scratch, sm_count, stream)
#define TC_BWD(D_, O_)                                            \
  if (d == D_ && order == O_) {                                   \
    if (feat == 0) return launch_bwd_tc_t<D_, O_, 0> TC_BWD_ARGS; \
    if (feat == 1) return launch_bwd_tc_t<D_, O_, 1> TC_BWD_ARGS; \
    return launch_bwd_tc_t<D_, O_, 2> TC_BWD_ARGS;                \
  }
  TC_BWD(1, 0); TC_BWD(1, 1); TC_BWD(1, 2);
  TC_BWD(2, 0); TC_BWD(2, 1); TC_BWD(2, 2);
  TC_BWD(3, 0); TC_BWD(3, 1); TC_BWD(3, 2);
  TC_BWD(1, 3); TC_BWD(2, 3); TC_BWD(3, 3);
#undef TC_BWD
#undef TC_BWD_ARGS
  return cudaErrorInvalidValue;
}

}  // namespace mpinn
