// conv_api.cu — C-ABI entry points for a single 3x3 convolution (unit tests, profiling).
#include "kb_common.cuh"
#include "kb_kernels.h"
#include "../../include/keisei_b200.h"

extern "C" int kb_conv3x3_forward(const void* in, const void* w, void* out, int B, int Cin, int Cout, int dtype,
                                  int backend, const float* scale, const float* shift, int relu, const float* gbias,
                                  double* ch_sums, float* board_mean, float* pool, int num_sms, cudaStream_t stream) {
  KB_CHECK_ARG(in && w && out, "kb_conv3x3_forward: null pointer");
  KB_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0, "kb_conv3x3_forward: bad shape");
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "kb_conv3x3_forward: bad dtype");
  ConvEpi e; memset(&e, 0, sizeof(e));
  e.scale = scale; e.shift = shift; e.relu = relu; e.gbias = gbias;
  if (ch_sums) { e.ch_sum = ch_sums; e.ch_sumsq = ch_sums + Cout; }
  e.board_sum = board_mean; e.board_scale = 1.f / 81.f; e.pool = pool;
  if (backend >= 1 && backend <= 3) {   // 1: tcgen05, kernel chosen automatically; 2: single-CTA kernel; 3: CTA-pair (cta_group::2) kernel
    KB_CHECK_ARG(B >= 3, "kb_conv3x3_forward: tcgen05 path needs at least 3 boards");
    KB_CHECK_ARG(kbk_conv3x3_tc_supported(Cin, Cout, dtype), "kb_conv3x3_forward: tcgen05 path needs bf16, Cin%%64==0, Cout%%128==0 (got %d,%d,dtype %d)", Cin, Cout, dtype);
    return kbk_conv3x3_tc_mode(in, w, out, B, Cin, Cout, e, num_sms, backend - 1, stream);
  }
  return kbk_conv3x3_simt(in, w, out, B, Cin, Cout, dtype, e, stream);
}

// conv2 of a GlobalPoolBiasBlock in evaluation mode with the whole block tail fused into the epilogue (CTA-pair kernel)
extern "C" int kb_conv3x3_se_tail(const void* in, const void* w, void* out, int B, int Cin, int Cout, const float* scale,
                                  const float* shift, const void* res, const float* se_w1, const float* se_b1,
                                  const float* se_w2, const float* se_b2, int S, float* pool, void* pool_bf16, int num_sms,
                                  cudaStream_t stream) {
  KB_CHECK_ARG(in && w && out && scale && shift && res && se_w1 && se_b1 && se_w2 && se_b2 && pool, "kb_conv3x3_se_tail: null pointer");
  KB_CHECK_ARG(B >= 3, "kb_conv3x3_se_tail: needs at least 3 boards");
  KB_CHECK_ARG(kbk_conv3x3_se_tail_supported(Cin, Cout, S, KB_BF16), "kb_conv3x3_se_tail: needs Cout == 256, S == 16, Cin %% 64 == 0 (got %d, %d, %d)", Cout, S, Cin);
  ConvEpi e; memset(&e, 0, sizeof(e));
  e.scale = scale; e.shift = shift; e.res = res; e.se_w1 = se_w1; e.se_b1 = se_b1; e.se_w2 = se_w2; e.se_b2 = se_b2;
  e.pool = pool; e.pool_bf = pool_bf16;
  return kbk_conv3x3_tc_mode(in, w, out, B, Cin, Cout, e, num_sms, 2, stream);
}

extern "C" long long kb_conv3x3_wgrad_ws_bytes(int Cin, int Cout, int num_sms) {
  return kbk_conv3x3_wgrad_tc_ws_bytes(Cin, Cout, num_sms);
}

extern "C" int kb_conv3x3_wgrad(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true,
                                int dtype, int backend, void* ws, long long ws_bytes, int num_sms, cudaStream_t stream) {
  KB_CHECK_ARG(x && dy && dw, "kb_conv3x3_wgrad: null pointer");
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "kb_conv3x3_wgrad: bad dtype");
  if (backend == 1) {
    KB_CHECK_ARG(kbk_conv3x3_tc_supported(Cin, Cout, dtype) && Cin <= 256, "kb_conv3x3_wgrad: tcgen05 path unsupported for this shape/dtype");
    return kbk_conv3x3_wgrad_tc(x, dy, dw, B, Cin, Cout, Cin_true, (float*)ws, ws_bytes, num_sms, stream);
  }
  return kbk_conv3x3_wgrad_simt(x, dy, dw, B, Cin, Cout, Cin_true, dtype, stream);
}

extern "C" int kb_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cinp, int dtype,
                                   cudaStream_t stream) {
  KB_CHECK_ARG(w && wf && Cinp >= Cin, "kb_pack_conv_weight: bad arguments");
  return kbk_pack_conv_weight(w, wf, wd, Cout, Cin, Cinp, dtype, stream);
}

extern "C" int kb_pack_linear_weight(const float* w, void* out_bf16, int N, int K, int Np, int Kp, cudaStream_t stream) {
  KB_CHECK_ARG(w && out_bf16 && Np >= N && Kp >= K, "kb_pack_linear_weight: bad arguments");
  return kbk_pack_linear_weight(w, out_bf16, N, K, Np, Kp, stream);
}

extern "C" int kb_linear_tc(const void* x_bf16, long long M, int Kp, const void* w_bf16, int N, int Np, const float* scale,
                            const float* bias, int relu, float* out_f32, long long ld_f, void* out_bf16, long long ld_b,
                            int nb_store, int group_rows, long long group_pitch, int num_sms, cudaStream_t stream) {
  KB_CHECK_ARG(x_bf16 && w_bf16 && (out_f32 || out_bf16), "kb_linear_tc: null pointer");
  return kbk_linear_tc(x_bf16, M, Kp, w_bf16, N, Np, scale, bias, relu, out_f32, ld_f, out_bf16, ld_b, nb_store, group_rows,
                       group_pitch, num_sms, stream);
}

extern "C" int kb_se_block_tail(const void* z, const void* res, void* out, const float* bn_a, const float* bn_b,
                                const float* board_mean, const float* w1, const float* b1, const float* w2, const float* b2,
                                float* se_in_out, float* seh_out, float* se_out, int se_raw, float* pool, void* pool_bf16,
                                float* ties, int B, int C, int S, int num_sms, cudaStream_t stream) {
  return kb_se_block_tail_variant(z, res, out, bn_a, bn_b, board_mean, w1, b1, w2, b2, se_in_out, seh_out, se_out, se_raw, pool,
                                  pool_bf16, ties, B, C, S, 0, num_sms, stream);
}

extern "C" int kb_se_block_tail_variant(const void* z, const void* res, void* out, const float* bn_a, const float* bn_b,
                                        const float* board_mean, const float* w1, const float* b1, const float* w2,
                                        const float* b2, float* se_in_out, float* seh_out, float* se_out, int se_raw,
                                        float* pool, void* pool_bf16, float* ties, int B, int C, int S, int variant,
                                        int num_sms, cudaStream_t stream) {
  KB_CHECK_ARG(variant >= 0 && variant <= 2, "kb_se_block_tail_variant: variant must be 0, 1 or 2");
  KB_CHECK_ARG(kbk_se_apply_supported(C, S), "kb_se_block_tail: needs bf16 tiles with C %% 8 == 0, 64 <= C <= 256 (got C=%d S=%d)", C, S);
  SeApplyArgs a; memset(&a, 0, sizeof(a));
  a.z = (const __nv_bfloat16*)z; a.res = (const __nv_bfloat16*)res; a.out = (__nv_bfloat16*)out; a.a = bn_a; a.b = bn_b;
  a.bmean = board_mean; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.se_in_out = se_in_out; a.seh_out = seh_out; a.se_out = se_out; a.se_raw = se_raw;
  a.pool = pool; a.pool_bf = (__nv_bfloat16*)pool_bf16; a.ties = ties; a.B = B; a.C = C; a.S = S;
  return kbk_se_apply_variant(a, variant, num_sms, stream);
}
