// common.cuh -- device-side program description and activation-jet algebra shared by all kernels.
//
// Jet algebra (restated from the reference's autograd behaviour, see oracle/jet_ref.py and
// SURVEY.md section 8 a-J): per point and feature carry z (value), z_i = dz/dx_i, z_ii = d2z/dx_i2.
//   a    = s0(z)
//   a_i  = s1(z) z_i
//   a_ii = s2(z) z_i^2 + s1(z) z_ii
// with s_k the k-th derivative of the activation.  GELU is the reference's literal
// x*0.5*(1+erf(x/1.41421)) (multipinn/neural_network/activation_function.py:14), Sine is sin x (:26).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpinn_b200.h"

namespace mpinn {

constexpr int kMaxOps = MPINN_MAX_OPS;

struct DevOp {
  int kind, in_dim, out_dim, act, flags;
  int w_off, b_off;      // float offsets into the flat parameter vector
  int pk_fwd, pk_bwd;    // float offsets into the packed buffer: Wt[in][outp] and Wn[out][inp]
  int outp, inp;         // dims padded to a multiple of 8
  int stash_plane;       // first stash plane (J*out_dim planes follow), -1 = not stashed
  int tc_off;            // float offset of the tensor-core operand block in the packed buffer, -1 = none
  int pad_;
};

struct DevProg {
  int n_ops, max_w, has_skip, d, n_out, total_planes;
  int64_t packed_floats;
  float lap_c[4];        // order 3 (Laplace mode): weights c_i of the single second-order channel sum_i c_i d2/dx_i2
  // reverse pass only: the last layer's bias gradient is the plain column sum of the value-channel seeds, usually the
  // largest entry of the whole gradient and a heavily cancelling sum.  api.cu sums it in double precision before
  // the engine runs (value_seed_colsum); the engines' fold kernels take entries [bsum_off, bsum_off + n_out) from it
  // instead of their own fp32 partials (nullptr: keep the engine's sum)
  const double* bsum;
  int bsum_off;
  DevOp ops[kMaxOps];
};

// fold kernels: entry i of the flat gradient comes from the double-precision seed sum when it is a last-layer bias
__device__ __forceinline__ bool take_bsum(const double* bsum, int bsum_off, int n_out, int64_t i, float* s) {
  if (bsum == nullptr || i < bsum_off || i >= bsum_off + n_out) return false;
  *s = (float)bsum[i - bsum_off];
  return true;
}

// number of jet channels: order 0/1/2 -> 1 + order*d;  order 3 (MPINN_ORDER_LAPLACE) -> value, d first derivatives, L
// order 4 (MPINN_ORDER_HESSIAN) -> value, d first, d diagonal second and d(d-1)/2 mixed second derivatives, the
// mixed pairs (i < j) in the order (0,1), (0,2), (1,2) after the diagonal ones
__host__ __device__ constexpr int jet_count(int d, int order) {
  return order == 3 ? d + 2 : (order == 4 ? 1 + 2 * d + d * (d - 1) / 2 : 1 + order * d);
}
// first / second direction of mixed pair m for dimension D
__host__ __device__ constexpr int mixed_i(int D, int m) { return (D == 3 && m == 2) ? 1 : 0; }
__host__ __device__ constexpr int mixed_j(int D, int m) { return (D == 3 && m >= 1) ? 2 : 1; }

// internal activation code: cos z, i.e. the second half of a Fourier encoding [sin(xB), cos(xB)] seen as a
// LINEAR(x -> xB) layer followed by an activation (reference neural_network/fourier_features.py:37-40)
constexpr int kActCos = 3;

constexpr float kGeluC = 1.0f / 1.41421f;             // the reference's literal constant
constexpr float kGeluC2 = kGeluC * kGeluC;
constexpr float kInvSqrtPi = 0.5641895835477563f;

// -DMPINN_ACCURATE_GELU=1 (diagnostic builds, scripts/build_variant.sh) swaps the two hardware approximations of the
// GELU basis for the accurate library functions, to measure what they cost in accuracy
#ifndef MPINN_ACCURATE_GELU
#define MPINN_ACCURATE_GELU 0
#endif
__device__ __forceinline__ float ex2_approx(float x) {
#if MPINN_ACCURATE_GELU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ float rcp_approx(float x) {
#if MPINN_ACCURATE_GELU
  return 1.0f / x;
#else
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// Phi(z) = 0.5 (1 + erf(c z)) and c/sqrt(pi) exp(-(c z)^2) from ONE exponential, branch-free:
//   0.5 erfc(|x|) = t (e0 + e1 t + ... + e5 t^5) exp(-x^2),  t = 1 / (1 + 0.55 |x|)
// (the classic rational-times-Gaussian form; coefficients are a minimax fit on [0, 6], approximation
// error 1.7e-8, i.e. below fp32 rounding; measured against fp64: rms 3e-8, max 2.6e-7 on Phi, the same
// level as 0.5 (1 + erff) while costing ~14 instructions instead of ~45 for erff + expf, and the negative
// tail keeps RELATIVE accuracy because it is never formed as 1 - (1 - tiny)).
__device__ __forceinline__ void gelu_basis(float z, float& P, float& Q) {
  const float cz = kGeluC * z;
  const float e = ex2_approx((cz * cz) * -1.4426950408889634f);
  const float t = rcp_approx(fmaf(0.55f, fabsf(cz), 1.0f));
  float h = 0.05366223313194056f;
  h = fmaf(h, t, -0.25901909217771745f);
  h = fmaf(h, t, 0.37985711822621195f);
  h = fmaf(h, t, -0.02071297485183388f);
  h = fmaf(h, t, 0.195420994420464f);
  h = fmaf(h, t, 0.15079172970262375f);
  h = (t * h) * e;
  P = cz < 0.0f ? h : 1.0f - h;
  Q = (kGeluC * kInvSqrtPi) * e;
}

// sincosf carries a ~160-instruction large-argument path; inlined at every (unrolled) call site it was a fifth of
// the reverse kernel's code although only the first layer uses it, so it is kept out of line.
static __device__ __noinline__ void sincos_call(float z, float& s, float& c) { sincosf(z, &s, &c); }

// The transcendental part of an activation at z, as two "basis" values from which s0..s3 follow with
// a few FMAs:  GELU -> (Phi(z), phi(z)),  Sine -> (sin z, cos z).
__device__ __forceinline__ void act_basis(int act, float z, float& P, float& Q) {
  if (act == MPINN_ACT_GELU) {
    gelu_basis(z, P, Q);
  } else if (act == MPINN_ACT_SINE || act == kActCos) {
    sincos_call(z, P, Q);
  } else {
    P = 0.0f;
    Q = 0.0f;
  }
}

// s0..s3 of the activation at z from its basis values.  NEED3 selects whether s3 is produced.
template <bool NEED3>
__device__ __forceinline__ void act_from_basis(int act, float z, float P, float Q, float& s0, float& s1,
                                               float& s2, float& s3) {
  if (act == MPINN_ACT_GELU) {
    const float cz = kGeluC * z;
    const float cz2 = cz * cz;
    s0 = z * P;
    s1 = fmaf(z, Q, P);
    s2 = Q * (2.0f - 2.0f * cz2);
    s3 = NEED3 ? (-2.0f * kGeluC2) * z * Q * (4.0f - 2.0f * cz2) : 0.0f;
  } else if (act == MPINN_ACT_SINE) {
    s0 = P;
    s1 = Q;
    s2 = -P;
    s3 = -Q;
  } else if (act == kActCos) {
    s0 = Q;
    s1 = -P;
    s2 = -Q;
    s3 = P;
  } else {
    s0 = z;
    s1 = 1.0f;
    s2 = 0.0f;
    s3 = 0.0f;
  }
}

// s0..s3 of the activation at z.
template <bool NEED3>
__device__ __forceinline__ void act_coeffs(int act, float z, float& s0, float& s1, float& s2,
                                           float& s3) {
  float P, Q;
  act_basis(act, z, P, Q);
  act_from_basis<NEED3>(act, z, P, Q, s0, s1, s2, s3);
}

// forward activation jets, in place: z[J] -> a[J]   (lc = DevProg::lap_c, read in Laplace mode only)
template <int D, int ORDER>
__device__ __forceinline__ void act_jets(int act, float* z, const float* lc) {
  if (act == MPINN_ACT_NONE) return;
  float s0, s1, s2, s3;
  act_coeffs<false>(act, z[0], s0, s1, s2, s3);
  z[0] = s0;
  if (ORDER == 3) {
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) q = fmaf(lc[i] * z[1 + i], z[1 + i], q);
    z[1 + D] = fmaf(s2, q, s1 * z[1 + D]);
  }
  if (ORDER == 4) {
#pragma unroll
    for (int m = 0; m < D * (D - 1) / 2; ++m)
      z[1 + 2 * D + m] = fmaf(s2 * z[1 + mixed_i(D, m)], z[1 + mixed_j(D, m)], s1 * z[1 + 2 * D + m]);
  }
  if (ORDER == 2 || ORDER == 4) {
#pragma unroll
    for (int i = 0; i < D; ++i) z[1 + D + i] = fmaf(s2 * z[1 + i], z[1 + i], s1 * z[1 + D + i]);
  }
  if (ORDER >= 1) {
#pragma unroll
    for (int i = 0; i < D; ++i) z[1 + i] *= s1;
  }
}

// cosine jets (Fourier encoding second half): z[J] -> cos jets
template <int D, int ORDER>
__device__ __forceinline__ void sincos_jets(const float* z, float* sj, float* cj, const float* lc) {
  float s, c;
  sincosf(z[0], &s, &c);
  sj[0] = s;
  cj[0] = c;
  if (ORDER >= 1) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      sj[1 + i] = c * z[1 + i];
      cj[1 + i] = -s * z[1 + i];
    }
  }
  if (ORDER == 2 || ORDER == 4) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float zi2 = z[1 + i] * z[1 + i];
      sj[1 + D + i] = fmaf(-s, zi2, c * z[1 + D + i]);
      cj[1 + D + i] = fmaf(-c, zi2, -s * z[1 + D + i]);
    }
  }
  if (ORDER == 4) {
#pragma unroll
    for (int m = 0; m < D * (D - 1) / 2; ++m) {
      const float zz = z[1 + mixed_i(D, m)] * z[1 + mixed_j(D, m)];
      sj[1 + 2 * D + m] = fmaf(-s, zz, c * z[1 + 2 * D + m]);
      cj[1 + 2 * D + m] = fmaf(-c, zz, -s * z[1 + 2 * D + m]);
    }
  }
  if (ORDER == 3) {
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) q = fmaf(lc[i] * z[1 + i], z[1 + i], q);
    sj[1 + D] = fmaf(-s, q, c * z[1 + D]);
    cj[1 + D] = fmaf(-c, q, -s * z[1 + D]);
  }
}

// reverse of act_jets: given z[J] (pre-activation jets) and ab[J] (grad wrt activation jets),
// overwrite ab with the grad wrt z.
template <int D, int ORDER>
__device__ __forceinline__ void act_jets_bwd(int act, const float* z, float* ab, const float* lc) {
  if (act == MPINN_ACT_NONE) return;
  float s0, s1, s2, s3;
  act_coeffs<(ORDER >= 2)>(act, z[0], s0, s1, s2, s3);
  float zb0 = ab[0] * s1;
  if (ORDER == 3) {
    // a_L = s2 q + s1 z_L with q = sum_i c_i z_i^2
    const float aL = ab[1 + D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float zi = z[1 + i];
      zb0 = fmaf(ab[1 + i] * s2, zi, zb0);
      zb0 = fmaf(aL * lc[i] * s3, zi * zi, zb0);
      ab[1 + i] = fmaf(2.0f * lc[i] * aL * s2, zi, ab[1 + i] * s1);
    }
    zb0 = fmaf(aL * s2, z[1 + D], zb0);
    ab[1 + D] = aL * s1;
    ab[0] = zb0;
    return;
  }
  float mixb[D > 0 ? D : 1];   // contributions of the mixed channels to the first-derivative channels
#pragma unroll
  for (int i = 0; i < D; ++i) mixb[i] = 0.0f;
  if (ORDER == 4) {
#pragma unroll
    for (int m = 0; m < D * (D - 1) / 2; ++m) {
      const int i = mixed_i(D, m), j = mixed_j(D, m);
      const float aij = ab[1 + 2 * D + m], zi = z[1 + i], zj = z[1 + j];
      mixb[i] = fmaf(aij * s2, zj, mixb[i]);
      mixb[j] = fmaf(aij * s2, zi, mixb[j]);
      zb0 = fmaf(aij, fmaf(s3 * zi, zj, s2 * z[1 + 2 * D + m]), zb0);
      ab[1 + 2 * D + m] = aij * s1;
    }
  }
  if (ORDER >= 1) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float zi = z[1 + i];
      float zbi = fmaf(ab[1 + i], s1, mixb[i]);
      zb0 = fmaf(ab[1 + i] * s2, zi, zb0);
      if (ORDER == 2 || ORDER == 4) {
        const float aii = ab[1 + D + i];
        zbi = fmaf(2.0f * aii * s2, zi, zbi);
        zb0 = fmaf(aii, fmaf(s3 * zi, zi, s2 * z[1 + D + i]), zb0);
        ab[1 + D + i] = aii * s1;
      }
      ab[1 + i] = zbi;
    }
  }
  ab[0] = zb0;
}

}  // namespace mpinn
