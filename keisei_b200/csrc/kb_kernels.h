// kb_kernels.h — internal (C++) launcher prototypes shared between translation units.
// Every launcher enqueues on the caller's stream, allocates nothing, and returns KB_OK or a
// negative error code after kb_set_error().
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "conv_epilogue.cuh"

// ---- conv_simt.cu ----
int kbk_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cinp, int dtype, cudaStream_t st);
int kbk_pack_obs(const float* obs, void* out, int B, int Cin, int Cinp, int dtype, cudaStream_t st);
int kbk_conv3x3_simt(const void* in, const void* w, void* out, int B, int Cin, int Cout, int dtype,
                     const ConvEpi& epi, cudaStream_t st);
int kbk_conv3x3_wgrad_simt(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true,
                           int dtype, cudaStream_t st);

// ---- pack_batched.cu: all weight re-packs of a network in one launch per 256 jobs ----
#define KB_PACK_CONV 0    // s0 = w (Cout,Cin,3,3) fp32; d0 = wf [Cout][9][Cinp], d1 = wd [Cinp][9][Cout] or null; n0..n2 = Cout, Cin, Cinp
#define KB_PACK_LINEAR 1  // s0 = w (N,K) fp32; d0 = bf16 [Np][Kp]; n0..n3 = N, K, Np, Kp
#define KB_PACK_BN 2      // s0..s3 = weight, bias, running_mean, running_var; d0 = a, d1 = b; count = C
struct PackJob {
  const float *s0, *s1, *s2, *s3;
  void *d0, *d1;
  long long count;        // elements of d0 this job writes
  int kind, dtype, n0, n1, n2, n3;
  float eps;
  int pad_;
};
int kbk_pack_batched(const PackJob* jobs, int n_jobs, cudaStream_t st);

// ---- conv_tc.cu (tcgen05 / TMEM / TMA, bf16) ----
int kbk_conv3x3_tc_supported(int Cin, int Cout, int dtype);
int kbk_conv3x3_se_tail_supported(int Cin, int Cout, int S, int dtype);   // fused evaluation tail in the CTA-pair kernel
int kbk_conv3x3_tc_mode(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi,
                        int num_sms, int mode, cudaStream_t st);
int kbk_conv3x3_tc_single(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi,
                          int num_sms, cudaStream_t st);
int kbk_conv3x3_tc(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi,
                   int num_sms, cudaStream_t st);
long long kbk_conv3x3_wgrad_tc_ws_bytes(int Cin, int Cout, int num_sms);
int kbk_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true, float* ws,
                         long long ws_bytes, int num_sms, cudaStream_t st);

// ---- gemm_tc.cu (tcgen05 Linear / 1x1-conv layers, bf16) ----
int kbk_pack_linear_weight(const float* w, void* out_bf16, int N, int K, int Np, int Kp, cudaStream_t st);
int kbk_cast_rows_bf16(const float* in, void* out_bf16, long long rows, int cols, int ld, cudaStream_t st);
// Y[m][n] = act((X[m][:] . W[n][:]) * scale[n] + bias[n]); X bf16 [M][Kp], W bf16 [Np][Kp] (zero padded);
// out_f32 [M][ld_f] (n < N) and/or out_bf16 [M][ld_b] (n < nb_store, zeros for n >= N; optionally board-pitched rows)
int kbk_linear_tc(const void* x, long long M, int Kp, const void* w, int N, int Np, const float* scale, const float* bias,
                  int relu, float* out_f32, long long ld_f, void* out_bf, long long ld_b, int nb_store, int group_rows,
                  long long group_pitch, int num_sms, cudaStream_t st);

// ---- gemm_simt.cu ----
struct GemmArgs {
  // C[M,N] = epilogue( prologue(op(A))[M,K] * op(B)[K,N] )
  const void* A; int a_dtype; long long lda; int transA;  // op(A)[m][k] = transA ? A[k*lda+m] : A[m*lda+k]
  int a_group_rows; long long a_group_pitch;               // if >0: row m lives at (m/gr)*pitch + (m%gr)*lda
  const float* a_pa; const float* a_pb; int a_relu;        // prologue: v = v*pa[k]+pb[k]; relu
  const void* B; int b_dtype; long long ldb; int transB;   // op(B)[k][n] = transB ? B[n*ldb+k] : B[k*ldb+n]
  void* C; int c_dtype; long long ldc;
  int c_group_rows; long long c_group_pitch;
  const float* bias;                                       // [N] or null
  int relu;
  const float* mask_src; long long ld_mask;                // v = mask_src[m][n] > 0 ? v : 0
  int M, N, K;
  int splitk;                                              // >1: fp32 atomicAdd into C (C pre-zeroed / accumulating)
  int tf32;                                                // inner product as TF32 tensor-core MMAs (bf16 / AMP path only)
};
// gpool_mlp_tc.cu: the whole global-pool-bias MLP (Linear 3C->128, ReLU, Linear 128->256) as one tcgen05 kernel
int kbk_gpool_mlp_tc_supported(int C, int G);
int kbk_gpool_mlp_tc(const void* x_bf16, int B, int K, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                     float* gh_out, float* g_out, int num_sms, cudaStream_t st);
int kbk_gemm(const GemmArgs& g, cudaStream_t st);
// Grouped launches: between begin and end (per host thread) kbk_gemm / kbk_colsum calls are COLLECTED (up to 4 + 4) and
// run as one kernel at the next flush / end — for problems that are independent of one another.
int kbk_gemm_group_begin();
int kbk_gemm_group_flush(cudaStream_t st);
int kbk_gemm_group_end(cudaStream_t st);
// out[n] += sum_m X[m][n]  (X may be board-pitched like GemmArgs.A)
int kbk_colsum(const void* X, int dtype, long long ldx, int group_rows, long long group_pitch, int M, int N,
               float* out, cudaStream_t st);

// ---- blocks.cu (BN / SE / pool elementwise + reductions) ----
struct ApplyArgs {
  // out = relu( (z*a[c]+b[c]) * sigmoid(se[b][c]) + se[b][C+c] + res ) + gbias[b][c]; optional pool stats of out
  const void* z; const float* a; const float* b;  // a/b null -> identity
  const float* se;      // [B][2C] (scale logits, shift) or null
  const void* res;      // [B][81][C] or null
  const float* gbias;   // [B][C] or null
  void* out;
  float* pool;          // [B][3C] or null
  float* ties;          // [B][C] or null: number of pixels equal to the board max (needs pool)
  void* pool_bf;        // [B][3C] bf16 or null: copy of pool for the tcgen05 Linear layers
  int B, C, dtype;
};
int kbk_apply(const ApplyArgs& a, cudaStream_t st);
// eval-mode folded BN: a = w/sqrt(rv+eps), b = bias - rm*a
int kbk_bn_eval_affine(const float* w, const float* bias, const float* rm, const float* rv, float eps, int C,
                       float* a, float* b, cudaStream_t st);
// training-mode BN finalize from double sums; updates running stats (momentum, unbiased var) and nbt
int kbk_bn_finalize(double* sums /*[2][C], zeroed afterwards*/, double count, const float* w, const float* bias,
                    const float* running_mean, const float* running_var, float* rm_out, float* rv_out, long long* nbt,
                    float momentum, float eps, int C, float* a, float* b, float* mean, float* invstd, cudaStream_t st);
// out[r][c] = in[r][c]*a[c] + b[c]  (fp32 [rows][C]; the SE squeeze input from the board means)
int kbk_affine_rows(const float* in, const float* a, const float* b, float* out, void* out_bf16, long long rows, int C,
                    cudaStream_t st);
// per-channel sum / sum of squares over rows of a [M][C] fp32 matrix (policy head BN)
int kbk_rows_stats(const float* x, long long M, int C, double* sums, cudaStream_t st);

// ---- backward elementwise ----
struct BlockBwdArgs {
  int B, C, dtype;
  const void* dxp;   // dL/d(block output) [B][81][C]
  const void* xp;    // block output (post-ReLU)
  const void* z2;    // raw conv2 output
  const float* a2; const float* b2;       // BN2 affine (a = gamma*invstd, b = beta - mean*a)
  const float* se;   // [B][2C]
  float* s_du;       // [B][C]  sum_p du
  float* s_duz;      // [B][C]  sum_p du*z2
};
int kbk_block_bwd_reduce(const BlockBwdArgs& a, cudaStream_t st);  // pass A
// SE backward glue: dse[b][c] = (a2*s_duz + b2*s_du) * sig'(scale), dse[b][C+c] = s_du
int kbk_se_bwd_prep(const float* s_du, const float* s_duz, const float* a2, const float* b2, const float* se,
                    float* dse, int B, int C, cudaStream_t st);
// BN2 backward channel sums: sum1[c] = sum_b (sig*s_du + dmean), sum2[c] = sum_b (sig*s_duz + dmean*bmean2)
// dmean = d(se_in)*a2 is NOT pre-scaled here: pass the gradient wrt zhat2's board mean, i.e. dse_in; bmean2 = mean_p z2
int kbk_bn2_bwd_sums(const float* s_du, const float* s_duz, const float* se, const float* dmean,
                     const float* bmean2, int B, int C, double* sums /*[2][C]*/, cudaStream_t st);
// BN backward finalize: from sums (sum dzh, sum dzh*z) -> m1, m2', dgamma, dbeta
int kbk_bn_bwd_finalize(double* sums /*[2][C], zeroed afterwards*/, double count, const float* w, const float* mean,
                        const float* invstd, float* k1, float* k2, float* k3, float* dgamma, float* dbeta, int C,
                        cudaStream_t st, double* sums_local = nullptr);
// ---- se_bwd.cu: the whole squeeze-excite branch backward of one block in one launch (C <= 256, S in {4,8,16,32}) ----
int kbk_se_mlp_bwd_supported(int C, int S);
int kbk_se_mlp_bwd(const float* s_du, const float* s_duz, const float* a2, const float* b2, const float* se,
                   const float* seh, const float* se_in, const float* bmean2, const float* W1, const float* W2,
                   float* dse_in, float* dW1, float* db1, float* dW2, float* db2, double* sums, int B, int C, int S,
                   int num_sms, cudaStream_t st);
struct PassBArgs {
  int B, C, dtype;
  const void* dxp; const void* xp; const void* z2;
  const float* se; const float* dse_in;  // [B][C] d(se_in) (gradient wrt the SE squeeze input)
  const float* k1; const float* k2; const float* k3;  // dz = k1*dzh - k2*z - k3  (per channel)
  void* dz2;
};
int kbk_block_bwd_dz2(const PassBArgs& a, cudaStream_t st);  // pass B
// pass C: dz = k1*dzh - k2*z - k3 in place over dzh
int kbk_bn_bwd_apply(void* dzh_inout, const void* z, const float* k1, const float* k2, const float* k3,
                     long long rows, int C, int dtype, cudaStream_t st);
struct PassDArgs {
  int B, C, dtype;
  const void* dxc;   // dgrad result (may be null -> 0)
  const void* dxp;   // upstream grad of the block output (may be null)
  const void* xp;    // block output (ReLU mask for dxp), may be null -> no mask
  const void* x;     // block input (for pool backward)
  const float* pool; // [B][3C] saved mean/max/std of x
  const float* dpool;// [B][3C] grad wrt pool (may be null)
  const float* ties; // [B][C] tie counts of the board max from the forward (null -> counted here)
  void* dx;          // out
  // Hand-off to the producing block (whose output is x): when mask_out != 0 the result is stored as
  // du = dx * [x > 0] (the gradient after that block's final ReLU), and with z_next (that block's raw conv2 output)
  // the per-(board, channel) sums of du and du * z_next come out as well — the producing block then needs no
  // separate reduction pass and never re-reads its output for the mask.
  int mask_out;
  const void* z_next; // [B][81][C] or null
  float* s_du;        // [B][C]  (with z_next)
  float* s_duz;       // [B][C]
};
int kbk_block_bwd_dx(const PassDArgs& a, cudaStream_t st);  // pass D
// stem / generic: dzh = dy * (y > 0), channel sums of dzh and dzh*z   (y = post-activation, z = raw)
int kbk_relu_bwd_stats(const void* dy, const void* y, const void* z, void* dzh, long long rows, int C, int dtype,
                       double* sums, cudaStream_t st);
// data-gradient tail when it is NOT fused in the conv epilogue: d = d * [z*ma+mb > 0] in place, channel sums of the
// masked d and d*z, board sums of the unmasked d (the gpool-bias gradient). Needs kbk_mask_bwd_stats_supported(C).
int kbk_mask_bwd_stats_supported(int C);
int kbk_mask_bwd_stats(void* d_inout, const void* z, const float* ma, const float* mb, float* board_sum, int B, int C,
                       int dtype, double* sums, cudaStream_t st);
// Two-pass variant of the pair above (one tensor pass less per residual block): the statistics kernel leaves `d`
// unmasked and the BatchNorm-backward apply recomputes the mask from z: dz = k1 * (d * [z*ma+mb > 0]) - k2*z - k3.
int kbk_mask_bwd_stats_ro_supported(int C);
int kbk_mask_bwd_stats_ro(const void* d, const void* z, const float* ma, const float* mb, float* board_sum, int B, int C,
                          int dtype, double* sums, cudaStream_t st);
int kbk_bn_bwd_apply_masked_supported(long long rows, int C);
int kbk_bn_bwd_apply_masked(void* d_inout, const void* z, const float* k1, const float* k2, const float* k3, const float* ma,
                            const float* mb, long long rows, int C, int dtype, cudaStream_t st);
// fp32 [M][C] variant for the policy head (mask by act > 0)
int kbk_relu_bwd_stats_f32(float* d_inout, const float* act, const float* z, long long M, int C, double* sums,
                           cudaStream_t st);
int kbk_fill_zero(void* p, size_t bytes, cudaStream_t st);

// ---- se_apply.cu: SE MLP + scale/shift + residual + ReLU + next-block pool statistics, TMA-bulk staged (bf16) ----
struct SeApplyArgs {
  const __nv_bfloat16* z;    // [B][81][C] conv2 output (raw when a/b are given, BN already folded when they are null)
  const __nv_bfloat16* res;  // [B][81][C] block input
  __nv_bfloat16* out;        // [B][81][C] block output
  const float* a; const float* b;            // [C] BN2 affine or null
  const float* bmean;        // [B][C] board means of z (before the a/b affine)
  const float* w1; const float* b1;          // se_fc1 [S][C], [S]
  const float* w2; const float* b2;          // se_fc2 [2C][S], [2C]
  float* se_in_out;          // [B][C] or null   (saved for backward)
  float* seh_out;            // [B][S] or null
  float* se_out;             // [B][2C] REQUIRED (scratch in eval): se_raw != 0 -> raw logits (saved for backward),
                             //         se_raw == 0 -> the scale half holds sigmoid(scale)
  int se_raw;
  float* pool;               // [B][3C] mean, max, std of out
  __nv_bfloat16* pool_bf;    // [B][3C] or null
  float* ties;               // [B][C] or null: pixels equal to the board max
  int B, C, S;
};
int kbk_se_apply_supported(int C, int S);
int kbk_se_apply(const SeApplyArgs& a, int num_sms, cudaStream_t st);
int kbk_se_apply_variant(const SeApplyArgs& a, int variant /*0 default, 1 TMA-staged, 2 column pair*/, int num_sms, cudaStream_t st);
// se_apply_col.cu: the same contract as a thread-per-channel MLP kernel + one column-layout streaming pass (default)
int kbk_se_apply_col_supported(int C, int S);
int kbk_se_apply_col(const SeApplyArgs& a, int num_sms, cudaStream_t st);

// ---- resnet_heads.cu (plain ResNet policy / value head front ends, reference models/resnet.py:49-59,76-84) ----
// raw [B*81][3] fp32 = x [B][81][C] . {policy_conv rows 0,1; value_conv}; sums (optional) double[6] =
// {sum p0, sum p1, sumsq p0, sumsq p1, sum v, sumsq v} (policy as a [2][2] BN block, value as a [2][1] block at +4)
int kbk_resnet_head_conv(const void* x, int dtype, const float* wp, const float* wv, float* raw, int B, int C, double* sums,
                         cudaStream_t st);
// BN affine + ReLU + NCHW flatten -> p_flat [B][162] (+ bf16 copy with pitch_bf >= 162, pad zeroed), v_flat [B][81]
int kbk_resnet_head_act(const float* raw, const float* ap, const float* bp, const float* av, const float* bv, float* p_flat,
                        void* p_flat_bf, int pitch_bf, float* v_flat, int B, cudaStream_t st);
// ReLU mask + un-flatten -> d3 [B*81][3]; sums double[6] = {sum d p0, sum d p1, sum d*raw p0, sum d*raw p1, sum d v, sum d*raw v}
int kbk_resnet_head_bwd_act(const float* dp_flat, const float* dv_flat, const float* raw, const float* ap, const float* bp,
                            const float* av, const float* bv, float* d3, int B, double* sums, cudaStream_t st);
// BN backward (k1,k2,k3 at kp/kv + {0, kstride, 2*kstride}), dx [B][81][C] written, dwp [2][C] / dwv [C] accumulated
int kbk_resnet_head_bwd_x(const void* x, int dtype, const float* d3, const float* raw, const float* kp, const float* kv,
                          int kstride, const float* wp, const float* wv, void* dx, float* dwp, float* dwv, int B, int C,
                          cudaStream_t st);
int kbk_tanh_fwd(const float* in, float* out, float* out2, long long n, cudaStream_t st);
int kbk_tanh_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st);
