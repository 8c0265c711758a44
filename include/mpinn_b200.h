/* mpinn_b200.h -- C ABI of the B200-native PINN train-step library (libmpinn_b200.so).
 *
 * The reference (multipinn) is pure Python on PyTorch and has no FFI of its own; the "operator
 * API" this library sits behind is the set of Python call signatures on its hot path.  Each
 * entry point below names the reference interface it replaces (paths relative to the reference
 * tree).  The Python host (multipinn_b200/_lib.py) binds these with ctypes; see INTEGRATION.md
 * for the stub a multipinn maintainer would add.
 *
 * Conventions: every pointer is a DEVICE pointer unless it says "host"; all arrays are fp32,
 * dense, row-major; `stream` is a cudaStream_t passed as void*; every call is asynchronous on
 * that stream and returns 0 on success or a non-zero code with a message available through
 * mpinn_last_error().  No entry point allocates device memory.
 *
 * Jet channels of a network output, per point:  0 = value, 1..d = d/dx_i, d+1..2d = d2/dx_i^2;
 * J = 1 + order*d for order in {0,1,2}.
 * order 3 (MPINN_ORDER_LAPLACE) carries the value, the d first derivatives and ONE second-order channel
 * L = sum_i c_i d2/dx_i^2 with caller-supplied constant weights c (J = d + 2): what a residual needs when it
 * uses second derivatives only through a Laplacian / a single u_xx (reference examples/poisson_2D_1C/problem.py:24-37,
 * examples/burgers_1D/problem.py:15-18).  It is exact, not an approximation: L obeys its own layer recursion
 *   a_L = s''(z) sum_i c_i z_i^2 + s'(z) z_L.
 */
#ifndef MPINN_B200_H
#define MPINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPINN_ABI_VERSION 3
#define MPINN_MAX_OPS 64

enum { MPINN_OP_LINEAR = 0, MPINN_OP_FOURIER = 1 };
enum { MPINN_ACT_NONE = 0, MPINN_ACT_SINE = 1, MPINN_ACT_GELU = 2 };
enum { MPINN_FLAG_SAVE_SKIP = 1, MPINN_FLAG_ADD_SKIP = 2, MPINN_FLAG_TRAINABLE = 4 };
enum { MPINN_ENGINE_AUTO = 0, MPINN_ENGINE_FFMA = 1, MPINN_ENGINE_TCGEN05 = 2 };
#define MPINN_ORDER_LAPLACE 3
#define MPINN_ORDER_HESSIAN 4   /* value, d first, d diagonal second, then the d(d-1)/2 mixed second derivatives
                                   (pairs i<j in the order (0,1), (0,2), (1,2)): J = 1 + 2d + d(d-1)/2.  What
                                   grad(u_x, arg) returns in its off-diagonal columns in the reference
                                   (condition/diff.py:47).  FP32-FMA engine only. */

/* One layer of the network "program" (host struct, passed by pointer to host memory).
 *   LINEAR : z = W h + b (+ skip), a = act(z); W is [out_dim,in_dim] at params+w_off, b at params+b_off
 *            (the nn.Linear layout: reference neural_network/feedforward_neural_network.py:28-37)
 *   FOURIER: a = [sin(xB), cos(xB)], B is [in_dim,out_dim/2] at params+w_off
 *            (reference neural_network/fourier_features.py:30-40)
 *   flags  : SAVE_SKIP remembers this op's input; ADD_SKIP adds it before the activation
 *            (reference neural_network/light_residual_block.py:33-38); TRAINABLE on a FOURIER op asks the reverse
 *            pass for dL/dB (FourierEncoding(is_trainable=True), fourier_features.py:30-32; FP32-FMA engine)   */
typedef struct mpinn_op {
  int32_t kind, in_dim, out_dim, w_off, b_off, act, flags, reserved;
} mpinn_op_t;

int mpinn_abi_version(void);
const char* mpinn_last_error(void);

/* Bytes of device scratch the forward pass needs to stash pre-activations for the reverse pass,
 * and bytes of workspace the reverse pass needs (per-CTA partial parameter gradients). */
size_t mpinn_stash_bytes(const mpinn_op_t* ops_host, int n_ops, int d, int order, int64_t n_points);
size_t mpinn_backward_workspace_bytes(const mpinn_op_t* ops_host, int n_ops, int64_t n_params);
/* Bytes of the packed-weight buffer (transposed / padded / hi-lo split copies of the weights).
 * The buffer must be 16-byte aligned (its operand images are fetched by the TMA engine). */
size_t mpinn_packed_bytes(const mpinn_op_t* ops_host, int n_ops);

/* Re-pack the flat parameter vector into the kernels' operand layouts.  Call once after every
 * parameter update (one small launch). */
int mpinn_pack_params(const mpinn_op_t* ops_host, int n_ops, const float* params, void* packed,
                      void* stream);

/* Forward jets of the whole network.
 * Replaces `model(arg)` + the nested `grad(u,arg)` sweeps:
 *   reference neural_network/feedforward_neural_network.py:39-52 (FNN.forward),
 *   residual_neural_network.py:114-126, fourier_features.py:73-76, condition/diff.py:25-47.
 *   x    [n,d]      collocation points
 *   out  [J,n,o]    jets of the o network outputs (o = out_dim of the last op)
 *   stash           mpinn_stash_bytes() bytes, or NULL for a value-only/inference pass (nothing is written for a
 *                   reverse pass; the tcgen05 engine does this at order 0 for programs without skip connections,
 *                   MPINN_ENGINE_AUTO otherwise runs the FP32-FMA engine)
 *   lap_coef_host   HOST pointer to the d weights c_i of the second-order channel when order == 3, else NULL */
int mpinn_jet_forward(const mpinn_op_t* ops_host, int n_ops, const float* params,
                      const void* packed, const float* x, int64_t n, int d, int order,
                      const float* lap_coef_host, float* out, void* stash, int engine, void* stream);

/* Reverse pass: parameter gradient of a scalar loss given dL/d(out jets).
 * Replaces `loss.backward()` through the double-backward graph (reference trainer/trainer.py:141).
 *   dout    [J,n,o]  gradient w.r.t. the forward's `out`
 *   dparams [P]      ACCUMULATED into (+=); P = number of floats in `params`.  The last layer's bias entries (the
 *                    column sums of dout[0], usually the largest entries of the gradient) are summed in double
 *                    precision and are identical from run to run; the rest is fp32 with per-CTA partial slabs
 *   workspace        mpinn_backward_workspace_bytes() bytes (partial slabs; its last 64 bytes hold those sums)
 *   dx      [n,d]    written if non-NULL (gradient w.r.t. the points: the `points.grad` the reference's
 *                    `batch_loss.backward()` leaves behind), else skipped; with a FOURIER first op on the FP32-FMA
 *                    engine only (forward and reverse with MPINN_ENGINE_FFMA: the tcgen05 reverse kernel refuses it) */
int mpinn_jet_backward(const mpinn_op_t* ops_host, int n_ops, const float* params,
                       const void* packed, const float* x, int64_t n, int d, int order,
                       const float* lap_coef_host, const float* dout, const void* stash, float* dparams,
                       int64_t n_params, float* dx, void* workspace, size_t workspace_bytes, int engine,
                       void* stream);

/* 1 if `engine` (MPINN_ENGINE_*) can run this program at (d, order), else 0 (host-only query). */
int mpinn_engine_supports(const mpinn_op_t* ops_host, int n_ops, int d, int order, int engine);

/* Diagnostic: one tcgen05.mma product through the same descriptor / layout / issue helpers the jet kernels use, so
 * that the tensor-core conventions can be pinned in isolation by tests.
 *   mode 0: c[128][64] = a[128][64] * b[64][64]^T, 3xTF32, K-major operands (forward pass)
 *   mode 2: the same product as BF16x3 (three bf16 pieces per operand, six piece products), K-major (reverse pass)
 *   mode 3: c[128][64] = stacked(a)^T b with a, b [128][64], BF16x3, MN-major operands; rows 0..63 plus rows
 *           64..127 of c give a^T b (weight-gradient GEMM of the reverse pass)
 *   mode 4: the same product with the two leading b pieces N-stacked: c[128][128], the four 64 x 64 quadrants add up
 *           to a^T b (the weight-gradient GEMM when at most four jet channels leave 128 accumulator columns free)
 * (mode 1, an MN-major TF32 variant, was removed: it needs a swizzle layout the kernels do not use.) */
int mpinn_selftest_umma(int mode, const float* a, const float* b, float* c, void* stream);

/* Diagnostic: when the library is built with -DMPINN_TC_TIMING=1 the tcgen05 reverse kernel accumulates
 * clock64() deltas per code segment (CTA 0, threads 0 and 224: 2 x 16 counters), and so does the forward kernel
 * (counters 32..63); this synchronises the device, copies up to n <= 64 counters to out_host and clears them.
 * All zeros in a normal build. */
int mpinn_debug_counters(long long* out_host, int n);

/* sum_i r[i*stride]^2 added into *accum (device scalar).
 * Replaces `torch.sum(r**2)` (reference PINN/PINN.py:59,95,123). */
int mpinn_sumsq(const float* r, int64_t n, int64_t stride, float* accum, void* stream);

/* Fused Adam over flat buffers; g is multiplied by grad_scale first (1/world for DP).
 * Replaces `optimizer.step()` with torch.optim.Adam(lr, betas, eps) (reference
 * trainer/trainer.py:150,235,335; examples/<name>/configs/config.yaml:41-44).  `step` is the
 * 1-based step count of this update. */
int mpinn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    int64_t n, int64_t step, float lr, float beta1, float beta2, float eps,
                    float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Residual programs.  The user's residual callable (reference examples/<name>/problem.py, evaluated per
 * batch by Condition.get_residual, reference condition/condition.py:62-71) is traced once into a
 * straight-line SSA program over per-point scalars: instruction i defines register i.
 *   CONST a        r = consts[a]                       LOADX a      r = x[n][a]
 *   LOADJ a b      r = jets[a][j][n][k], b = j*o + k   (block a = the a-th model(arg) call)
 *   LOADD a        r = data[n][a]                      (per-point auxiliary data of ConditionExtra, e.g. boundary
 *                                                       normals: reference condition/condition.py:117-215)
 *   ADD SUB MUL DIV MAX MIN a b ;  NEG SIN COS EXP TANH RELU ABS SQRT LOG a ;  POWI a b (integer exponent b) ;
 *   POWF a b (exponent = constant register b) ;  GT GE a b (1.0 / 0.0 masks: `p > c`) ;
 *   SEL a b c   r = reg[a] != 0 ? reg[b] : reg[c]   (torch.where, `p[p > c] = 0`: reference
 *                                                    examples/gas_dynamics_1D/problem.py:33-37)
 * jet_reg[a][j*o+k] is the register that loads that channel, or -1 if the residual does not read it. */
enum {
  MPINN_RP_CONST = 0, MPINN_RP_LOADX = 1, MPINN_RP_LOADJ = 2, MPINN_RP_LOADD = 3,
  MPINN_RP_ADD = 10, MPINN_RP_SUB = 11, MPINN_RP_MUL = 12, MPINN_RP_DIV = 13, MPINN_RP_MAX = 14, MPINN_RP_MIN = 15,
  MPINN_RP_NEG = 20, MPINN_RP_SIN = 21, MPINN_RP_COS = 22, MPINN_RP_EXP = 23, MPINN_RP_TANH = 24,
  MPINN_RP_RELU = 25, MPINN_RP_ABS = 26, MPINN_RP_SQRT = 27, MPINN_RP_LOG = 28,
  MPINN_RP_POWI = 30, MPINN_RP_POWF = 31,
  MPINN_RP_GT = 40, MPINN_RP_GE = 41, MPINN_RP_SEL = 42
};
#define MPINN_RP_MAX_INSTR 192
#define MPINN_RP_MAX_CONST 64
#define MPINN_RP_MAX_OUT 8
#define MPINN_RP_MAX_BLOCKS 8
#define MPINN_RP_MAX_DATA 16
#define MPINN_RP_MAX_JO 32

typedef struct mpinn_rinstr { int16_t op, a, b, c; } mpinn_rinstr_t;
typedef struct mpinn_rprog {
  int32_t n_instr, n_const, n_out, n_blocks, d, n_data, reserved0, reserved1;
  int32_t block_J[MPINN_RP_MAX_BLOCKS], block_o[MPINN_RP_MAX_BLOCKS];
  int16_t out_reg[MPINN_RP_MAX_OUT];
  int16_t jet_reg[MPINN_RP_MAX_BLOCKS][MPINN_RP_MAX_JO];
  mpinn_rinstr_t instr[MPINN_RP_MAX_INSTR];
  float consts[MPINN_RP_MAX_CONST];
} mpinn_rprog_t;

/* Evaluate the residual components r_k[n] of one condition at n points and, in the same launch,
 *   sums[k]    += sum_n r_k[n]^2                                (replaces torch.sum(r**2), reference PINN/PINN.py:123)
 *   djets[a]    = d/d(jets[a]) of  sum_k weights[k] * sum_n r_k[n]^2   (the seeds of mpinn_jet_backward; replaces the
 *                 autograd graph between the residual list and the network outputs)
 *   values[k][n] = r_k[n]                                       (optional: what Condition.get_residual returns;
 *                                                                 values_point_major != 0 stores values[n][k] instead)
 * point_weights (device [n_out][n], or NULL): per-point loss weights lambda_k[n] -- the loss whose seeds are written
 * becomes sum_k weights[k] sum_n lambda_k[n] r_k[n]^2 and wsums[k] (device [n_out], or NULL) += sum_n lambda_k[n] r_k[n]^2
 * (reference regularization/pointwise.py:84-100: a bias-free Linear(N, 1) applied to r**2); sums stays unweighted.
 * data is the device [n, n_data] row-major block LOADD reads (NULL when prog->n_data == 0).
 * prog_host and the two pointer tables are HOST memory; jets[a] / djets[a] are device [J_a, n, o_a] blocks.
 * djets, weights (device [n_out]), sums (device [n_out]) and values may each be NULL.  jet_stride (0 = n) is the
 * number of points of the jet blocks when they are longer than this launch's n: several conditions of the same
 * jet order share one forward / reverse launch over their concatenated points, and each condition's program runs on
 * its slice (jets[a] / djets[a] then point at the slice's first point).  A program without model
 * calls (n_blocks == 0, jets_host NULL) is allowed: it maps point columns to point columns, which is how the points
 * of a second model call such as model(arg + 1) (reference examples/regression_1D/problem.py:17-35) or a shifted
 * boundary (examples/burgers_1D/problem.py:20-25) are produced on the device. */
int mpinn_residual_program(const mpinn_rprog_t* prog_host, const float* x, const float* data, int64_t n,
                           const float* const* jets_host, float* const* djets_host,
                           const float* weights, float* sums, float* values, int values_point_major,
                           int64_t jet_stride, const float* point_weights, float* wsums, void* stream);

/* Adam with its scalar state on the device, so that a captured CUDA graph of the train step can be replayed
 * while the learning rate and the step count change: state = {lr, step, step_size, sqrt(1 - beta2^step)}.
 * mpinn_adam_tick increments state[1] and refreshes state[2..3] (bias corrections in double precision);
 * mpinn_adam_step_dev then applies the same update as mpinn_adam_step.  The caller writes state[0] (lr). */
int mpinn_adam_tick(float* state, float beta1, float beta2, void* stream);
int mpinn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* state, float beta1, float beta2, float eps, float grad_scale,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPINN_B200_H */
