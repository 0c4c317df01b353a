// resnet.cu — forward / backward schedule of the plain ResNet baseline (reference
// keisei/training/models/resnet.py:25-84: `ResidualBlock`, `ResNetModel`; BASELINE.json configs[3])
// as a stream-ordered sequence of this library's kernels. Pure host code: no allocation, no global
// state; the caller owns parameters, packed weights and the workspace.
//
// Parameter table order == PyTorch registration order of the reference model:
//   input_conv.weight, input_bn.{weight,bias}, per block conv1.weight, bn1.{weight,bias}, conv2.weight,
//   bn2.{weight,bias}; then policy_conv.weight, policy_bn.{weight,bias}, policy_fc.{weight,bias},
//   value_conv.weight, value_bn.{weight,bias}, value_fc1.{weight,bias}, value_fc2.{weight,bias}.
// Buffer table order: input_bn.{running_mean,running_var,num_batches_tracked}, per block bn1.*, bn2.*,
//   then policy_bn.*, value_bn.*.
// The trunk reuses the SE-ResNet's convolution / BatchNorm / residual kernels; the heads (1x1 conv to
// 2 + 1 channels, BN, ReLU, NCHW flatten) are resnet_heads.cu; policy_fc (162 -> 11259) runs on the
// tcgen05 Linear kernel in bf16 and on the SIMT GEMM in fp32.
#include <stdlib.h>
#include "kb_common.cuh"
#include "kb_kernels.h"
#include "schedule_common.cuh"
#include "../../include/keisei_b200.h"

namespace {

using namespace kbs;

constexpr float kBnEps = 1e-5f, kBnMomentum = 0.1f;
constexpr int kA = 81 * 139;       // 11,259 actions
constexpr int kPF = 162;           // policy_fc input features (2 channels x 81)
constexpr int kPFk = 192;          // ... padded to the tcgen05 K block
constexpr int kANp = 11264;        // 11,259 padded to 128 output features

struct RDims {
  int L, C, C0, C0p, B, dtype;
  size_t esz;
  long long M;
  size_t act() const { return (size_t)B * 81 * C * esz; }
};
RDims make_dims(const kb_resnet_desc* d, int B, int dtype) {
  RDims m;
  m.L = d->num_layers; m.C = d->hidden_size; m.C0 = d->obs_channels; m.C0p = ((d->obs_channels + 63) / 64) * 64;
  m.B = B; m.dtype = dtype; m.esz = dtype == KB_F32 ? 4 : 2; m.M = (long long)B * 81;
  return m;
}
inline int pi_blk(int i, int j) { return 3 + 6 * i + j; }
inline int pi_head(const RDims& m, int j) { return 3 + 6 * m.L + j; }
inline int bi_blk(int i, int j) { return 3 + 6 * i + j; }
inline int bi_head(const RDims& m, int j) { return 3 + 6 * m.L + j; }
// BN layer numbering: 0 = input_bn, 1+2i = bn1, 2+2i = bn2 of block i, 2L+1 = policy_bn, 2L+2 = value_bn
inline int n_bn(const RDims& m) { return 2 * m.L + 3; }

struct WPack {
  void* stem_wf; char* blocks; size_t conv_bytes; float* bn_eval; void* fc_bf; size_t total;
  int C;
  void* wf(int i, int conv) const { return blocks + ((size_t)i * 4 + conv * 2) * conv_bytes; }
  void* wd(int i, int conv) const { return blocks + ((size_t)i * 4 + conv * 2 + 1) * conv_bytes; }
  float* bn_a(int l) const { return bn_eval + (size_t)l * 2 * C; }
  float* bn_b(int l) const { return bn_eval + (size_t)l * 2 * C + C; }
};
WPack make_wpack(const RDims& m, void* base) {
  WPack w; Bump b(base);
  w.C = m.C;
  w.conv_bytes = (((size_t)m.C * 9 * m.C * m.esz) + 1023) & ~(size_t)1023;
  w.stem_wf = b.take((size_t)m.C * 9 * m.C0p * m.esz);
  b.off = (b.off + 1023) & ~(size_t)1023;
  w.blocks = (char*)b.take((size_t)m.L * 4 * w.conv_bytes);
  w.bn_eval = b.f32((size_t)n_bn(m) * 2 * m.C);
  b.off = (b.off + 1023) & ~(size_t)1023;
  w.fc_bf = m.dtype == KB_BF16 ? b.take((size_t)kANp * kPFk * 2) : nullptr;
  w.total = b.off + 256;
  return w;
}

struct BlockWs { void *z1, *a1, *z2, *xout; };
struct Ws {
  double* dsums;
  void *obs_p, *p_flat_bf;
  float *raw3, *p_flat, *v_flat, *vh, *vpre, *bn;
  void *z0, *x0; BlockWs* blk; float* vout;
  void *ea, *eb, *ey1, *ey2;
  void *d0, *d1, *d2, *d4;  // d4: fourth rotating gradient buffer (side-stream weight gradients, see model.cu)
  float *d3, *dp_flat, *dv_flat, *dvh, *dvpre, *k123;
  float* wg_ws; long long wg_ws_bytes;
  int C; size_t total;
  float* bn_a(int l) const { return bn + (size_t)l * 4 * C; }
  float* bn_b(int l) const { return bn + (size_t)l * 4 * C + C; }
  float* bn_mean(int l) const { return bn + (size_t)l * 4 * C + 2 * C; }
  float* bn_invstd(int l) const { return bn + (size_t)l * 4 * C + 3 * C; }
};
void carve(const RDims& m, void* base, int training, Ws& w, BlockWs* storage) {
  Bump b(base);
  const size_t B = m.B;
  w.C = m.C; w.blk = storage;
  w.dsums = (double*)b.take(2 * (size_t)(m.C > 4 ? m.C : 4) * sizeof(double));
  w.obs_p = b.take(B * 81 * m.C0p * m.esz);
  w.raw3 = b.f32((size_t)m.M * 3);
  w.p_flat = b.f32(B * kPF);
  w.p_flat_bf = b.take(B * kPFk * 2);
  w.v_flat = b.f32(B * 81);
  w.vh = b.f32(B * m.C);
  w.vpre = b.f32(B);
  w.bn = b.f32((size_t)n_bn(m) * 4 * m.C);
  if (training) {
    w.z0 = b.take(m.act()); w.x0 = b.take(m.act());
    for (int i = 0; i < m.L; ++i) {
      BlockWs bw; bw.z1 = b.take(m.act()); bw.a1 = b.take(m.act()); bw.z2 = b.take(m.act()); bw.xout = b.take(m.act());
      if (storage) storage[i] = bw;
    }
    w.vout = b.f32(B);
    w.d0 = b.take(m.act()); w.d1 = b.take(m.act()); w.d2 = b.take(m.act()); w.d4 = b.take(m.act());
    w.d3 = b.f32((size_t)m.M * 3); w.dp_flat = b.f32(B * kPF); w.dv_flat = b.f32(B * 81); w.dvh = b.f32(B * m.C);
    w.dvpre = b.f32(B); w.k123 = b.f32(2 * 3 * (size_t)m.C);
    w.wg_ws_bytes = (m.dtype == KB_BF16 && m.C % 128 == 0 && m.C <= 256) ? kbk_conv3x3_wgrad_tc_ws_bytes(m.C, m.C, 148 * 2) : 0;
    w.wg_ws = w.wg_ws_bytes ? (float*)b.take((size_t)w.wg_ws_bytes) : nullptr;
    w.ea = w.eb = w.ey1 = w.ey2 = nullptr;
  } else {
    w.ea = b.take(m.act()); w.eb = b.take(m.act()); w.ey1 = b.take(m.act()); w.ey2 = b.take(m.act());
    w.z0 = w.x0 = nullptr; w.vout = nullptr; w.d0 = w.d1 = w.d2 = w.d4 = nullptr;
    w.d3 = w.dp_flat = w.dv_flat = w.dvh = w.dvpre = w.k123 = nullptr; w.wg_ws = nullptr; w.wg_ws_bytes = 0;
  }
  w.total = b.off + 256;
}

int conv3x3(const RDims& m, const void* in, const void* wgt, void* out, int Cin, int Cout, const ConvEpi& e, int use_tc,
            int num_sms, cudaStream_t st) {
  if (use_tc && m.B >= 3 && kbk_conv3x3_tc_supported(Cin, Cout, m.dtype)) return kbk_conv3x3_tc(in, wgt, out, m.B, Cin, Cout, e, num_sms, st);
  return kbk_conv3x3_simt(in, wgt, out, m.B, Cin, Cout, m.dtype, e, st);
}
int wgrad3x3(const RDims& m, const void* x, const void* dy, float* dw, int Cin, int Cout, int Cin_true, int use_tc,
             int num_sms, float* wg_ws, long long wg_ws_bytes, cudaStream_t st) {
  if (use_tc && m.B >= 3 && Cin <= 256 && wg_ws != nullptr && kbk_conv3x3_tc_supported(Cin, Cout, m.dtype))
    return kbk_conv3x3_wgrad_tc(x, dy, dw, m.B, Cin, Cout, Cin_true, wg_ws, wg_ws_bytes, num_sms, st);
  return kbk_conv3x3_wgrad_simt(x, dy, dw, m.B, Cin, Cout, Cin_true, m.dtype, st);
}

int check_desc(const kb_resnet_desc* d) {
  KB_CHECK_ARG(d != nullptr, "null model descriptor");
  KB_CHECK_ARG(d->num_layers >= 0 && d->num_layers <= 1024, "num_layers out of range");
  KB_CHECK_ARG(d->hidden_size >= 4 && d->hidden_size % 4 == 0 && d->hidden_size <= 1024,
               "hidden_size=%d must be a multiple of 4 in [4,1024]", d->hidden_size);
  KB_CHECK_ARG(d->obs_channels >= 1 && d->obs_channels <= 128, "bad obs_channels");
  return KB_OK;
}

}  // namespace

extern "C" long long kb_resnet_num_params(const kb_resnet_desc* d) { return d ? 15 + 6LL * d->num_layers : -1; }
extern "C" long long kb_resnet_num_buffers(const kb_resnet_desc* d) { return d ? 9 + 6LL * d->num_layers : -1; }

extern "C" long long kb_resnet_wpack_bytes(const kb_resnet_desc* d, int dtype) {
  if (check_desc(d) != KB_OK) return -1;
  return (long long)make_wpack(make_dims(d, 1, dtype), nullptr).total;
}

extern "C" long long kb_resnet_workspace_bytes(const kb_resnet_desc* d, int B, int training, int dtype) {
  if (check_desc(d) != KB_OK || B < 0) return -1;
  Ws w;
  carve(make_dims(d, B, dtype), nullptr, training, w, nullptr);
  return (long long)w.total;
}

extern "C" int kb_resnet_pack_weights(const kb_resnet_desc* d, const void* const* params, const void* const* buffers,
                                      int dtype, void* wpack, long long wpack_bytes, cudaStream_t st) {
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  const RDims m = make_dims(d, 1, dtype);
  WPack w = make_wpack(m, wpack);
  KB_CHECK_ARG(wpack && (size_t)wpack_bytes >= w.total, "wpack buffer too small: %lld < %zu", wpack_bytes, w.total);
  KB_TRY(kbk_pack_conv_weight((const float*)params[0], w.stem_wf, nullptr, m.C, m.C0, m.C0p, dtype, st));
  auto bn = [&](int layer, int pw, int bbase, int C) {
    return kbk_bn_eval_affine((const float*)params[pw], (const float*)params[pw + 1], (const float*)buffers[bbase],
                              (const float*)buffers[bbase + 1], kBnEps, C, w.bn_a(layer), w.bn_b(layer), st);
  };
  KB_TRY(bn(0, 1, 0, m.C));
  for (int i = 0; i < m.L; ++i) {
    KB_TRY(kbk_pack_conv_weight((const float*)params[pi_blk(i, 0)], w.wf(i, 0), w.wd(i, 0), m.C, m.C, m.C, dtype, st));
    KB_TRY(kbk_pack_conv_weight((const float*)params[pi_blk(i, 3)], w.wf(i, 1), w.wd(i, 1), m.C, m.C, m.C, dtype, st));
    KB_TRY(bn(1 + 2 * i, pi_blk(i, 1), bi_blk(i, 0), m.C));
    KB_TRY(bn(2 + 2 * i, pi_blk(i, 4), bi_blk(i, 3), m.C));
  }
  KB_TRY(bn(2 * m.L + 1, pi_head(m, 1), bi_head(m, 0), 2));
  KB_TRY(bn(2 * m.L + 2, pi_head(m, 6), bi_head(m, 3), 1));
  if (w.fc_bf) KB_TRY(kbk_pack_linear_weight((const float*)params[pi_head(m, 3)], w.fc_bf, kA, kPF, kANp, kPFk, st));
  return KB_OK;
}

extern "C" int kb_resnet_forward(const kb_resnet_desc* d, const void* const* params, void* const* buffers, float* new_stats,
                                 const void* wpack, const float* obs, int B, int training, int dtype, void* workspace,
                                 long long ws_bytes, void* policy_out, long long policy_pitch, float* value_out, int use_tc,
                                 int num_sms, cudaStream_t st) {
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  KB_CHECK_ARG(B >= 1, "batch must be >= 1");
  KB_CHECK_ARG(policy_pitch >= kA, "policy pitch %lld < 11259", policy_pitch);
  KB_CHECK_ARG(params && buffers && wpack && obs && workspace && policy_out && value_out, "null pointer");
  const RDims m = make_dims(d, B, dtype);
  const WPack wp = make_wpack(m, const_cast<void*>(wpack));
  BlockWs* blks = (BlockWs*)alloca(sizeof(BlockWs) * (m.L > 0 ? m.L : 1));
  Ws w;
  carve(m, workspace, training, w, blks);
  KB_CHECK_ARG((size_t)ws_bytes >= w.total, "workspace too small: %lld < %zu", ws_bytes, w.total);
  const int C = m.C;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto BUF = [&](int i) { return (float*)buffers[i]; };
  const double count = (double)m.M;
  const int LPOL = 2 * m.L + 1, LVAL = 2 * m.L + 2;
  auto bn_fin = [&](double* sums, int layer, int pw, int bbase, int Cl) {
    float* rm_out = new_stats ? new_stats + (size_t)layer * 2 * C : BUF(bbase);
    float* rv_out = new_stats ? new_stats + (size_t)layer * 2 * C + C : BUF(bbase + 1);
    long long* nbt = new_stats ? nullptr : (long long*)buffers[bbase + 2];
    return kbk_bn_finalize(sums, count, P(pw), P(pw + 1), BUF(bbase), BUF(bbase + 1), rm_out, rv_out, nbt, kBnMomentum, kBnEps,
                           Cl, w.bn_a(layer), w.bn_b(layer), w.bn_mean(layer), w.bn_invstd(layer), st);
  };
  auto apply = [&](const void* z, const float* a, const float* b, const void* res, void* out) {
    ApplyArgs ap; memset(&ap, 0, sizeof(ap));
    ap.z = z; ap.a = a; ap.b = b; ap.res = res; ap.out = out; ap.B = B; ap.C = C; ap.dtype = dtype;
    return kbk_apply(ap, st);
  };

  KB_TRY(kbk_fill_zero(w.dsums, 2 * (size_t)(C > 4 ? C : 4) * sizeof(double), st));
  KB_TRY(kbk_pack_obs(obs, w.obs_p, B, m.C0, m.C0p, dtype, st));

  // ---- stem: relu(bn(conv(obs)))  (resnet.py:73) ----
  void* x_cur;
  if (training) {
    ConvEpi e = epi_base();
    e.ch_sum = w.dsums; e.ch_sumsq = w.dsums + C;
    KB_TRY(conv3x3(m, w.obs_p, wp.stem_wf, w.z0, m.C0p, C, e, use_tc, num_sms, st));
    KB_TRY(bn_fin(w.dsums, 0, 1, 0, C));
    KB_TRY(apply(w.z0, w.bn_a(0), w.bn_b(0), nullptr, w.x0));
    x_cur = w.x0;
  } else {
    ConvEpi e = epi_base();
    e.scale = wp.bn_a(0); e.shift = wp.bn_b(0); e.relu = 1;
    KB_TRY(conv3x3(m, w.obs_p, wp.stem_wf, w.ea, m.C0p, C, e, use_tc, num_sms, st));
    x_cur = w.ea;
  }
  // ---- residual tower (resnet.py:33-37): relu(bn2(conv2(relu(bn1(conv1(x))))) + x) ----
  for (int i = 0; i < m.L; ++i) {
    const int l1 = 1 + 2 * i, l2 = 2 + 2 * i;
    if (training) {
      BlockWs& bw = blks[i];
      ConvEpi e = epi_base();
      e.ch_sum = w.dsums; e.ch_sumsq = w.dsums + C;
      KB_TRY(conv3x3(m, x_cur, wp.wf(i, 0), bw.z1, C, C, e, use_tc, num_sms, st));
      KB_TRY(bn_fin(w.dsums, l1, pi_blk(i, 1), bi_blk(i, 0), C));
      KB_TRY(apply(bw.z1, w.bn_a(l1), w.bn_b(l1), nullptr, bw.a1));
      KB_TRY(conv3x3(m, bw.a1, wp.wf(i, 1), bw.z2, C, C, e, use_tc, num_sms, st));
      KB_TRY(bn_fin(w.dsums, l2, pi_blk(i, 4), bi_blk(i, 3), C));
      KB_TRY(apply(bw.z2, w.bn_a(l2), w.bn_b(l2), x_cur, bw.xout));
      x_cur = bw.xout;
    } else {
      ConvEpi e = epi_base();
      e.scale = wp.bn_a(l1); e.shift = wp.bn_b(l1); e.relu = 1;
      KB_TRY(conv3x3(m, x_cur, wp.wf(i, 0), w.ey1, C, C, e, use_tc, num_sms, st));
      ConvEpi e2 = epi_base();
      e2.scale = wp.bn_a(l2); e2.shift = wp.bn_b(l2);
      KB_TRY(conv3x3(m, w.ey1, wp.wf(i, 1), w.ey2, C, C, e2, use_tc, num_sms, st));
      void* xout = (x_cur == w.ea) ? w.eb : w.ea;
      KB_TRY(apply(w.ey2, nullptr, nullptr, x_cur, xout));
      x_cur = xout;
    }
  }
  // ---- heads (resnet.py:76-84) ----
  const float *ap, *bp, *av, *bv;
  KB_TRY(kbk_resnet_head_conv(x_cur, dtype, P(pi_head(m, 0)), P(pi_head(m, 5)), w.raw3, B, C, training ? w.dsums : nullptr, st));
  if (training) {
    KB_TRY(bn_fin(w.dsums, LPOL, pi_head(m, 1), bi_head(m, 0), 2));
    KB_TRY(bn_fin(w.dsums + 4, LVAL, pi_head(m, 6), bi_head(m, 3), 1));
    ap = w.bn_a(LPOL); bp = w.bn_b(LPOL); av = w.bn_a(LVAL); bv = w.bn_b(LVAL);
  } else {
    ap = wp.bn_a(LPOL); bp = wp.bn_b(LPOL); av = wp.bn_a(LVAL); bv = wp.bn_b(LVAL);
  }
  const bool fc_tc = use_tc && dtype == KB_BF16 && wp.fc_bf != nullptr;
  KB_TRY(kbk_resnet_head_act(w.raw3, ap, bp, av, bv, w.p_flat, fc_tc ? w.p_flat_bf : nullptr, kPFk, w.v_flat, B, st));
  if (fc_tc) {
    KB_TRY(kbk_linear_tc(w.p_flat_bf, B, kPFk, wp.fc_bf, kA, kANp, nullptr, P(pi_head(m, 4)), 0, nullptr, 0, policy_out,
                         policy_pitch, kA, 0, 0, num_sms, st));
  } else {
    KB_TRY(linear_fwd(w.p_flat, KB_F32, kPF, B, kPF, P(pi_head(m, 3)), kA, P(pi_head(m, 4)), 0, policy_out, dtype, policy_pitch, st));
  }
  KB_TRY(linear_fwd(w.v_flat, KB_F32, 81, B, 81, P(pi_head(m, 8)), C, P(pi_head(m, 9)), 1, w.vh, KB_F32, C, st));
  KB_TRY(linear_fwd(w.vh, KB_F32, C, B, C, P(pi_head(m, 10)), 1, P(pi_head(m, 11)), 0, w.vpre, KB_F32, 1, st));
  KB_TRY(kbk_tanh_fwd(w.vpre, value_out, training ? w.vout : nullptr, B, st));
  return KB_OK;
}

extern "C" int kb_resnet_backward(const kb_resnet_desc* d, const void* const* params, const void* wpack, int B, int dtype,
                                  void* workspace, long long ws_bytes, const void* dpolicy, long long policy_pitch,
                                  const float* dvalue, void* const* grads, int use_tc, int num_sms, cudaStream_t st) {
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  KB_CHECK_ARG(B >= 1 && policy_pitch >= kA, "bad shape");
  KB_CHECK_ARG(params && wpack && workspace && dpolicy && dvalue && grads, "null pointer");
  const RDims m = make_dims(d, B, dtype);
  const WPack wp = make_wpack(m, const_cast<void*>(wpack));
  BlockWs* blks = (BlockWs*)alloca(sizeof(BlockWs) * (m.L > 0 ? m.L : 1));
  Ws w;
  carve(m, workspace, 1, w, blks);
  KB_CHECK_ARG((size_t)ws_bytes >= w.total, "workspace too small: %lld < %zu", ws_bytes, w.total);
  Tf32Scope tf32_scope(dtype == KB_BF16 && use_tc ? 1 : 0);  // head backward GEMMs on the tensor cores (AMP path only)
  const int C = m.C;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto G = [&](int i) { return (float*)grads[i]; };
  const double count = (double)m.M;
  const int LPOL = 2 * m.L + 1, LVAL = 2 * m.L + 2;
  float *k1 = w.k123, *k2 = w.k123 + C, *k3 = w.k123 + 2 * C;  // trunk / policy BN coefficients
  float* kv = w.k123 + 3 * C;                                   // value BN coefficients (same stride C)
  const void* x_last = m.L > 0 ? blks[m.L - 1].xout : w.x0;

  // ---- policy_fc: dW = dlogits^T p_flat, db = colsum, dp_flat = dlogits W ----
  {
    GemmArgs g = gemm_base();
    g.A = dpolicy; g.a_dtype = dtype; g.lda = policy_pitch; g.transA = 1;
    g.B = w.p_flat; g.b_dtype = KB_F32; g.ldb = kPF; g.transB = 0;
    g.C = G(pi_head(m, 3)); g.c_dtype = KB_F32; g.ldc = kPF; g.M = kA; g.N = kPF; g.K = B;
    g.splitk = kb_ceil_div(B, 512) < 2 ? 2 : kb_ceil_div(B, 512);
    KB_TRY(kbk_gemm(g, st));
    KB_TRY(kbk_colsum(dpolicy, dtype, policy_pitch, 0, 0, B, kA, G(pi_head(m, 4)), st));
    KB_TRY(linear_bwd_x(dpolicy, dtype, policy_pitch, B, kA, P(pi_head(m, 3)), kPF, w.dp_flat, KB_F32, kPF, nullptr, 0, 0, st));
  }
  // ---- value head: tanh, fc2, ReLU, fc1 ----
  KB_TRY(kbk_tanh_bwd(dvalue, w.vout, w.dvpre, B, st));
  KB_TRY(linear_bwd_w(w.dvpre, KB_F32, 1, w.vh, KB_F32, C, B, 1, C, G(pi_head(m, 10)), G(pi_head(m, 11)), st));
  KB_TRY(linear_bwd_x(w.dvpre, KB_F32, 1, B, 1, P(pi_head(m, 10)), C, w.dvh, KB_F32, C, w.vh, C, 0, st));
  KB_TRY(linear_bwd_w(w.dvh, KB_F32, C, w.v_flat, KB_F32, 81, B, C, 81, G(pi_head(m, 8)), G(pi_head(m, 9)), st));
  KB_TRY(linear_bwd_x(w.dvh, KB_F32, C, B, C, P(pi_head(m, 8)), 81, w.dv_flat, KB_F32, 81, nullptr, 0, 0, st));
  // ---- head front ends: ReLU mask, BN backward (2 + 1 channels), 1x1 conv weight / data gradients ----
  KB_TRY(kbk_resnet_head_bwd_act(w.dp_flat, w.dv_flat, w.raw3, w.bn_a(LPOL), w.bn_b(LPOL), w.bn_a(LVAL), w.bn_b(LVAL), w.d3, B,
                                 w.dsums, st));
  KB_TRY(kbk_bn_bwd_finalize(w.dsums, count, P(pi_head(m, 1)), w.bn_mean(LPOL), w.bn_invstd(LPOL), k1, k2, k3,
                             G(pi_head(m, 1)), G(pi_head(m, 2)), 2, st));
  KB_TRY(kbk_bn_bwd_finalize(w.dsums + 4, count, P(pi_head(m, 6)), w.bn_mean(LVAL), w.bn_invstd(LVAL), kv, kv + C, kv + 2 * C,
                             G(pi_head(m, 6)), G(pi_head(m, 7)), 1, st));
  // weight-gradient convolutions on the side stream, exactly as in the SE-ResNet schedule (model.cu): each one is
  // released just before its data-gradient conv; events per block: [0]/[1] operand ready (main -> side),
  // [2]/[3] conv2 / conv1 weight gradient done (side -> main); a fourth buffer takes pass D's output
  const bool overlap = m.L > 0 && bwd_overlap_enabled(st);
  SideStream& side = side_stream_tls();
  if (overlap) {
    int dev = 0;
    KB_CUDA_CHECK(cudaGetDevice(&dev));
    KB_TRY(side.ensure(dev, 4 * m.L + 1));
  }
  cudaStream_t wst = overlap ? side.s : st;
  auto ev = [&](int blk, int k) { return side.ev[4 * blk + k]; };
  void *cur = w.d0, *t1 = w.d1, *t2 = w.d2, *t3 = w.d4;
  KB_TRY(kbk_resnet_head_bwd_x(x_last, dtype, w.d3, w.raw3, k1, kv, C, P(pi_head(m, 0)), P(pi_head(m, 5)), cur, G(pi_head(m, 0)),
                               G(pi_head(m, 5)), B, C, st));

  // ---- residual tower, last block first ----
  for (int i = m.L - 1; i >= 0; --i) {
    BlockWs& bw = blks[i];
    const void* x_in = i > 0 ? blks[i - 1].xout : w.x0;
    const int l1 = 1 + 2 * i, l2 = 2 + 2 * i;
    // du = dx' * [x' > 0] -> BN2 backward -> dz2
    KB_TRY(kbk_relu_bwd_stats(cur, bw.xout, bw.z2, t1, m.M, C, dtype, w.dsums, st));
    KB_TRY(kbk_bn_bwd_finalize(w.dsums, count, P(pi_blk(i, 4)), w.bn_mean(l2), w.bn_invstd(l2), k1, k2, k3, G(pi_blk(i, 4)),
                               G(pi_blk(i, 5)), C, st));
    KB_TRY(kbk_bn_bwd_apply(t1, bw.z2, k1, k2, k3, m.M, C, dtype, st));
    if (overlap) { KB_CUDA_CHECK(cudaEventRecord(ev(i, 0), st)); KB_CUDA_CHECK(cudaStreamWaitEvent(wst, ev(i, 0), 0)); }
    KB_TRY(wgrad3x3(m, bw.a1, t1, G(pi_blk(i, 3)), C, C, C, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, wst));
    if (overlap) KB_CUDA_CHECK(cudaEventRecord(ev(i, 2), wst));
    ConvEpi e = epi_base();
    KB_TRY(conv3x3(m, t1, wp.wd(i, 1), t2, C, C, e, use_tc, num_sms, st));
    // ReLU mask of a1, BN1 backward -> dz1
    // a1 = relu(bn1(z1)): the mask is a function of z1 and the BN1 affine, so a1 is not re-read (3 passes instead of 4)
    // ... and the masked gradient is not written back either: the dz1 pass recomputes the mask (2 + 3 passes)
    static int mask_recompute_env = -1;
    if (mask_recompute_env < 0) { const char* me = getenv("KB_MASK_RECOMPUTE"); mask_recompute_env = (me && me[0] == '0') ? 0 : 1; }
    const bool mask_pending = mask_recompute_env && kbk_mask_bwd_stats_ro_supported(C) && kbk_bn_bwd_apply_masked_supported(m.M, C);
    if (mask_pending) KB_TRY(kbk_mask_bwd_stats_ro(t2, bw.z1, w.bn_a(l1), w.bn_b(l1), nullptr, B, C, dtype, w.dsums, st));
    else if (kbk_mask_bwd_stats_supported(C)) KB_TRY(kbk_mask_bwd_stats(t2, bw.z1, w.bn_a(l1), w.bn_b(l1), nullptr, B, C, dtype, w.dsums, st));
    else KB_TRY(kbk_relu_bwd_stats(t2, bw.a1, bw.z1, t2, m.M, C, dtype, w.dsums, st));
    KB_TRY(kbk_bn_bwd_finalize(w.dsums, count, P(pi_blk(i, 1)), w.bn_mean(l1), w.bn_invstd(l1), k1, k2, k3, G(pi_blk(i, 1)),
                               G(pi_blk(i, 2)), C, st));
    if (mask_pending) KB_TRY(kbk_bn_bwd_apply_masked(t2, bw.z1, k1, k2, k3, w.bn_a(l1), w.bn_b(l1), m.M, C, dtype, st));
    else KB_TRY(kbk_bn_bwd_apply(t2, bw.z1, k1, k2, k3, m.M, C, dtype, st));
    if (overlap) { KB_CUDA_CHECK(cudaEventRecord(ev(i, 1), st)); KB_CUDA_CHECK(cudaStreamWaitEvent(wst, ev(i, 1), 0)); }
    KB_TRY(wgrad3x3(m, x_in, t2, G(pi_blk(i, 0)), C, C, C, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, wst));
    if (overlap) {
      KB_CUDA_CHECK(cudaEventRecord(ev(i, 3), wst));
      KB_CUDA_CHECK(cudaStreamWaitEvent(st, ev(i, 2), 0));  // t1 (dz2) is about to be overwritten
    }
    KB_TRY(conv3x3(m, t2, wp.wd(i, 0), t1, C, C, e, use_tc, num_sms, st));
    // dx = data gradient + skip branch (du recomputed from dx' and the ReLU mask of x'); written to the fourth buffer:
    // t2 (dz1) is still read by this block's conv1 weight gradient, t3 was the previous block's dz1
    PassDArgs pd; memset(&pd, 0, sizeof(pd));
    pd.B = B; pd.C = C; pd.dtype = dtype; pd.dxc = t1; pd.dxp = cur; pd.xp = bw.xout; pd.dx = t3;
    if (overlap && i + 1 < m.L) KB_CUDA_CHECK(cudaStreamWaitEvent(st, ev(i + 1, 3), 0));
    KB_TRY(kbk_block_bwd_dx(pd, st));
    void* nc = t3; t3 = t2; t2 = t1; t1 = cur; cur = nc;
  }
  if (overlap) {  // join before the stem's weight gradient reuses the partial-tile workspace
    KB_CUDA_CHECK(cudaEventRecord(side.ev[4 * m.L], wst));
    KB_CUDA_CHECK(cudaStreamWaitEvent(st, side.ev[4 * m.L], 0));
  }
  // ---- stem ----
  KB_TRY(kbk_relu_bwd_stats(cur, w.x0, w.z0, t1, m.M, C, dtype, w.dsums, st));
  KB_TRY(kbk_bn_bwd_finalize(w.dsums, count, P(1), w.bn_mean(0), w.bn_invstd(0), k1, k2, k3, G(1), G(2), C, st));
  KB_TRY(kbk_bn_bwd_apply(t1, w.z0, k1, k2, k3, m.M, C, dtype, st));
  KB_TRY(wgrad3x3(m, w.obs_p, t1, G(0), m.C0p, C, m.C0, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, st));
  return KB_OK;
}
