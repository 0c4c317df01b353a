// model.cu — the SE-ResNet forward / backward schedule (reference se_resnet.py:68-90, :132-159)
// as a stream-ordered sequence of this library's kernels. Pure host code: no allocation, no
// global state; the caller owns parameters, packed weights and the workspace.
//
// Parameter table order == PyTorch registration order of the reference model (so Adam state and
// strict state_dict loads line up): input_conv.weight, input_bn.{weight,bias}, then per block
// conv1.weight, bn1.{weight,bias}, conv2.weight, bn2.{weight,bias}, global_fc.{0,2}.{weight,bias},
// se_fc1.{weight,bias}, se_fc2.{weight,bias}; then policy_conv1.weight, policy_bn1.{weight,bias},
// policy_conv2.{weight,bias}, value_fc{1,2}.{weight,bias}, score_fc{1,2}.{weight,bias}.
// Buffer table order: input_bn.{running_mean,running_var,num_batches_tracked}, per block bn1.*, bn2.*,
// then policy_bn1.*.
#include <stdlib.h>
#include <vector>
#include "kb_common.cuh"
#include "kb_kernels.h"
#include "schedule_common.cuh"
#include "../../include/keisei_b200.h"

namespace {

using namespace kbs;

constexpr float kBnEps = 1e-5f, kBnMomentum = 0.1f;

struct Dims {
  int nb, C, S, G, Pc, V, S2, C0, C0p, B, dtype;
  size_t esz;
  long long M;  // B*81
  size_t act() const { return (size_t)B * 81 * C * esz; }
};

Dims make_dims(const kb_seresnet_desc* d, int B, int dtype) {
  Dims m;
  m.nb = d->num_blocks; m.C = d->channels; m.S = d->se_hidden; m.G = d->gpool_channels; m.Pc = d->policy_channels;
  m.V = d->value_fc; m.S2 = d->score_fc; m.C0 = d->obs_channels; m.C0p = ((d->obs_channels + 63) / 64) * 64;
  m.B = B; m.dtype = dtype; m.esz = dtype == KB_F32 ? 4 : 2; m.M = (long long)B * 81;
  return m;
}

inline int pi_blk(int i, int j) { return 3 + 14 * i + j; }
inline int pi_head(const Dims& m, int j) { return 3 + 14 * m.nb + j; }
inline int bi_blk(int i, int j) { return 3 + 6 * i + j; }
inline int bi_pol(const Dims& m, int j) { return 3 + 6 * m.nb + j; }

// packed-weight buffer: conv weights in the activation dtype, then eval-mode BN affines (fp32)
struct WPack {
  char* base;
  void *stem_wf;
  size_t conv_bytes, stem_bytes;
  char* blocks;     // per block: conv1 wf, conv1 wd, conv2 wf, conv2 wd
  float* bn_eval;   // [2*nb+2][2][Cmax]
  int Cmax;
  // bf16 zero-padded Linear / 1x1-conv weights for the tcgen05 GEMM ([Np][Kp], Np % 128 == 0, Kp % 64 == 0)
  char* lin; size_t lin_blk_bytes, o_g1, o_g2, o_s1, o_s2, o_heads, o_p1, o_p2, o_v1, o_v2, o_c1, o_c2;
  int Gp, Gk, Cp, Sp, Sk, C2p, Pp, Pk, Vp, Vk, Qp, Qk;
  void* lin_blk(int i, size_t off) const { return lin + (size_t)i * lin_blk_bytes + off; }
  void* lin_head(size_t off) const { return lin + o_heads + off; }
  size_t total;
  void* wf(int i, int conv) const { return blocks + ((size_t)i * 4 + conv * 2) * conv_bytes; }
  void* wd(int i, int conv) const { return blocks + ((size_t)i * 4 + conv * 2 + 1) * conv_bytes; }
  float* bn_a(int layer) const { return bn_eval + (size_t)layer * 2 * Cmax; }
  float* bn_b(int layer) const { return bn_eval + (size_t)layer * 2 * Cmax + Cmax; }
};
// BN layer numbering: 0 = input_bn, 1+2i = block i bn1, 2+2i = block i bn2, 2nb+1 = policy_bn1
WPack make_wpack(const Dims& m, void* base) {
  WPack w;
  w.base = (char*)base;
  Bump b(base);
  w.stem_bytes = (size_t)m.C * 9 * m.C0p * m.esz;
  w.conv_bytes = (((size_t)m.C * 9 * m.C * m.esz) + 1023) & ~(size_t)1023;
  w.stem_wf = b.take(w.stem_bytes);
  b.off = (b.off + 1023) & ~(size_t)1023;
  w.blocks = (char*)b.take((size_t)m.nb * 4 * w.conv_bytes);
  w.Cmax = m.C > m.Pc ? m.C : m.Pc;
  w.bn_eval = b.f32((size_t)(2 * m.nb + 2) * 2 * w.Cmax);
  auto up = [](int v, int a) { return (v + a - 1) / a * a; };
  w.Gp = up(m.G, 128); w.Gk = up(m.G, 64); w.Cp = up(m.C, 128); w.Sp = up(m.S, 128); w.Sk = up(m.S, 64);
  w.C2p = up(2 * m.C, 128); w.Pp = up(m.Pc, 128); w.Pk = up(m.Pc, 64); w.Vp = up(m.V, 128); w.Vk = up(m.V, 64);
  w.Qp = up(m.S2, 128); w.Qk = up(m.S2, 64);
  const int K3 = up(3 * m.C, 64), Kc = up(m.C, 64);
  size_t o = 0;
  auto place = [&](size_t& slot, size_t rows, size_t cols) { slot = o; o += (rows * cols * 2 + 1023) & ~(size_t)1023; };
  place(w.o_g1, w.Gp, K3); place(w.o_g2, w.Cp, w.Gk); place(w.o_s1, w.Sp, Kc); place(w.o_s2, w.C2p, w.Sk);
  w.lin_blk_bytes = o;
  o = 0;
  place(w.o_p1, w.Pp, Kc); place(w.o_p2, 256, w.Pk); place(w.o_v1, w.Vp, K3); place(w.o_v2, 128, w.Vk);
  place(w.o_c1, w.Qp, K3); place(w.o_c2, 128, w.Qk);
  const size_t heads_bytes = o;
  b.off = (b.off + 1023) & ~(size_t)1023;
  w.lin = (char*)b.take((size_t)m.nb * w.lin_blk_bytes + heads_bytes);
  w.o_heads = (size_t)m.nb * w.lin_blk_bytes;
  w.total = b.off + 256;
  return w;
}

struct BlockWs {
  void *z1, *a1, *z2, *xout;
  float *gh, *bmean2, *se_in, *seh, *se;
};

struct Ws {
  void *obs_p, *z0, *x0;
  BlockWs* blk;            // host array, nb entries (filled by carve)
  float* pools;            // [nb+1][B][3C]
  float* tie_counts;       // [nb+1][B][C]: pixels equal to the board max (amax backward)
  float* bn;               // [2nb+2][4][Cmax]: a, b, mean, invstd
  float *g, *p1raw, *p1act, *vh, *sh;
  double* dsums;           // [2][Cmax]
  double* dsums_local;     // [2][Cmax]: this rank's sums under SyncBatchNorm (dgamma / dbeta)
  // eval-only ping-pong
  void *ea, *eb, *ey1, *ey2;
  float *epool_a, *epool_b;
  // backward temporaries
  void *d0, *d1, *d2, *d3;
  float *s_du, *s_duz, *dse_in, *dg, *dse, *dseh, *dgh, *dpool, *k123, *dp1, *dvh, *dsh;
  float* wg_ws; long long wg_ws_bytes;  // tcgen05 weight-gradient partial tiles
  void *pool_bf, *gh_bf, *sein_bf, *seh_bf, *vh_bf, *sh_bf, *p1act_bf;  // bf16 operands of the tcgen05 Linear layers
  int Cmax;
  size_t total;
  float* pool(const Dims& m, int i) const { return pools + (size_t)i * m.B * 3 * m.C; }
  float* ties(const Dims& m, int i) const { return tie_counts + (size_t)i * m.B * m.C; }
  float* bn_a(int l) const { return bn + (size_t)l * 4 * Cmax; }
  float* bn_b(int l) const { return bn + (size_t)l * 4 * Cmax + Cmax; }
  float* bn_mean(int l) const { return bn + (size_t)l * 4 * Cmax + 2 * Cmax; }
  float* bn_invstd(int l) const { return bn + (size_t)l * 4 * Cmax + 3 * Cmax; }
};

void carve(const Dims& m, void* base, int training, Ws& w, BlockWs* blk_storage) {
  Bump b(base);
  const size_t B = m.B;
  w.Cmax = m.C > m.Pc ? m.C : m.Pc;
  w.blk = blk_storage;
  w.dsums = (double*)b.take(2 * (size_t)w.Cmax * sizeof(double));
  w.dsums_local = (double*)b.take(2 * (size_t)w.Cmax * sizeof(double));
  w.obs_p = b.take(B * 81 * m.C0p * m.esz);
  w.g = b.f32(B * m.C);
  w.p1raw = b.f32((size_t)m.M * m.Pc);
  w.p1act = b.f32((size_t)m.M * m.Pc);
  w.vh = b.f32(B * m.V);
  w.sh = b.f32(B * m.S2);
  w.bn = b.f32((size_t)(2 * m.nb + 2) * 4 * w.Cmax);
  {
    auto up64 = [](int v) { return (size_t)((v + 63) / 64 * 64); };
    w.pool_bf = b.take(B * up64(3 * m.C) * 2); w.gh_bf = b.take(B * up64(m.G) * 2); w.sein_bf = b.take(B * up64(m.C) * 2);
    w.seh_bf = b.take(B * up64(m.S) * 2); w.vh_bf = b.take(B * up64(m.V) * 2); w.sh_bf = b.take(B * up64(m.S2) * 2);
    w.p1act_bf = b.take((size_t)m.M * up64(m.Pc) * 2);
  }
  if (training) {
    w.z0 = b.take(m.act()); w.x0 = b.take(m.act());
    w.pools = b.f32((size_t)(m.nb + 1) * B * 3 * m.C);
    w.tie_counts = b.f32((size_t)(m.nb + 1) * B * m.C);
    for (int i = 0; i < m.nb; ++i) {
      BlockWs bw;
      bw.z1 = b.take(m.act()); bw.a1 = b.take(m.act()); bw.z2 = b.take(m.act()); bw.xout = b.take(m.act());
      bw.gh = b.f32(B * m.G); bw.bmean2 = b.f32(B * m.C); bw.se_in = b.f32(B * m.C); bw.seh = b.f32(B * m.S);
      bw.se = b.f32(B * 2 * m.C);
      if (blk_storage) blk_storage[i] = bw;
    }
    w.d0 = b.take(m.act()); w.d1 = b.take(m.act()); w.d2 = b.take(m.act()); w.d3 = b.take(m.act());
    w.s_du = b.f32(B * m.C); w.s_duz = b.f32(B * m.C); w.dse_in = b.f32(B * m.C); w.dg = b.f32(B * m.C);
    w.dse = b.f32(B * 2 * m.C); w.dseh = b.f32(B * m.S); w.dgh = b.f32(B * m.G); w.dpool = b.f32(B * 3 * m.C);
    w.k123 = b.f32(3 * (size_t)w.Cmax);
    w.dp1 = b.f32((size_t)m.M * m.Pc);
    w.dvh = b.f32(B * m.V); w.dsh = b.f32(B * m.S2);
    w.wg_ws_bytes = (m.dtype == KB_BF16 && m.C % 128 == 0 && m.C <= 256) ? kbk_conv3x3_wgrad_tc_ws_bytes(m.C, m.C, 148 * 2) : 0;
    w.wg_ws = w.wg_ws_bytes ? (float*)b.take((size_t)w.wg_ws_bytes) : nullptr;
    w.ea = w.eb = w.ey1 = w.ey2 = nullptr; w.epool_a = w.epool_b = nullptr;
  } else {
    w.ea = b.take(m.act()); w.eb = b.take(m.act()); w.ey1 = b.take(m.act()); w.ey2 = b.take(m.act());
    w.epool_a = b.f32(B * 3 * m.C); w.epool_b = b.f32(B * 3 * m.C);
    BlockWs bw;
    bw.z1 = bw.a1 = bw.z2 = bw.xout = nullptr;
    bw.gh = b.f32(B * m.G); bw.bmean2 = b.f32(B * m.C); bw.se_in = nullptr; bw.seh = b.f32(B * m.S); bw.se = b.f32(B * 2 * m.C);
    if (blk_storage) blk_storage[0] = bw;
    w.z0 = w.x0 = nullptr; w.pools = nullptr; w.tie_counts = nullptr;
    w.d0 = w.d1 = w.d2 = w.d3 = nullptr;
    w.s_du = w.s_duz = w.dse_in = w.dg = w.dse = w.dseh = w.dgh = w.dpool = w.k123 = w.dp1 = w.dvh = w.dsh = nullptr;
    w.wg_ws = nullptr; w.wg_ws_bytes = 0;
  }
  w.total = b.off + 256;
}

int conv3x3(const Dims& m, const void* in, const void* wgt, void* out, int Cin, int Cout, const ConvEpi& e, int use_tc,
            int num_sms, cudaStream_t st) {
  if (use_tc && m.B >= 3 && kbk_conv3x3_tc_supported(Cin, Cout, m.dtype)) return kbk_conv3x3_tc(in, wgt, out, m.B, Cin, Cout, e, num_sms, st);
  return kbk_conv3x3_simt(in, wgt, out, m.B, Cin, Cout, m.dtype, e, st);
}
int wgrad3x3(const Dims& m, const void* x, const void* dy, float* dw, int Cin, int Cout, int Cin_true, int use_tc,
             int num_sms, float* wg_ws, long long wg_ws_bytes, cudaStream_t st) {
  if (use_tc && m.B >= 3 && Cin <= 256 && wg_ws != nullptr && kbk_conv3x3_tc_supported(Cin, Cout, m.dtype))
    return kbk_conv3x3_wgrad_tc(x, dy, dw, m.B, Cin, Cout, Cin_true, wg_ws, wg_ws_bytes, num_sms, st);
  return kbk_conv3x3_wgrad_simt(x, dy, dw, m.B, Cin, Cout, Cin_true, m.dtype, st);
}

// tcgen05 Linear path: bf16 activations, every K a multiple of 64 after padding the small hidden sizes
bool lin_tc(const Dims& m, int use_tc) { return use_tc && m.dtype == KB_BF16 && m.C % 64 == 0 && m.B >= 1; }

// Gradient-bucket hook of the backward schedule (data-parallel overlap, reference katago_loop.py:498-504: DDP overlaps
// its 25 MB gradient buckets with the backward): called on the host, per thread, right after the kernels that produce a
// contiguous run of parameter gradients have been enqueued — the heads first, then every residual block from the last
// to the first. The callee records events on the two streams and launches its collective behind them.
thread_local kb_bucket_hook g_bucket_hook = nullptr;
thread_local void* g_bucket_user = nullptr;

int check_desc(const kb_seresnet_desc* d) {
  KB_CHECK_ARG(d != nullptr, "null model descriptor");
  KB_CHECK_ARG(d->num_blocks >= 0 && d->num_blocks <= 1024, "num_blocks out of range");
  KB_CHECK_ARG(d->channels >= 4 && d->channels % 4 == 0 && d->channels <= 1024, "channels=%d must be a multiple of 4 in [4,1024]", d->channels);
  KB_CHECK_ARG(d->se_hidden >= 1 && d->gpool_channels >= 1 && d->policy_channels >= 1 && d->policy_channels <= 1024 &&
               d->value_fc >= 1 && d->score_fc >= 1 && d->obs_channels >= 1 && d->obs_channels <= 128, "bad model descriptor");
  return KB_OK;
}

}  // namespace

extern "C" int kb_seresnet_set_bucket_hook(kb_bucket_hook hook, void* user) {
  g_bucket_hook = hook;
  g_bucket_user = user;
  return KB_OK;
}

extern "C" long long kb_seresnet_num_params(const kb_seresnet_desc* d) { return d ? 16 + 14LL * d->num_blocks : -1; }
extern "C" long long kb_seresnet_num_buffers(const kb_seresnet_desc* d) { return d ? 6 + 6LL * d->num_blocks : -1; }

extern "C" long long kb_seresnet_wpack_bytes(const kb_seresnet_desc* d, int dtype) {
  if (check_desc(d) != KB_OK) return -1;
  const Dims m = make_dims(d, 1, dtype);
  return (long long)make_wpack(m, nullptr).total;
}

extern "C" long long kb_seresnet_workspace_bytes(const kb_seresnet_desc* d, int B, int training, int dtype) {
  if (check_desc(d) != KB_OK || B < 0) return -1;
  const Dims m = make_dims(d, B, dtype);
  Ws w;
  carve(m, nullptr, training, w, nullptr);
  return (long long)w.total;
}

// Repack conv weights into the kernels' layouts (activation dtype) and fold eval-mode BatchNorm.
extern "C" int kb_seresnet_pack_weights(const kb_seresnet_desc* d, const void* const* params, const void* const* buffers,
                                        int dtype, void* wpack, long long wpack_bytes, cudaStream_t st) {
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  const Dims m = make_dims(d, 1, dtype);
  WPack w = make_wpack(m, wpack);
  KB_CHECK_ARG(wpack && (size_t)wpack_bytes >= w.total, "wpack buffer too small: %lld < %zu", wpack_bytes, w.total);
  // one job list, a handful of launches (pack_batched.cu) instead of ~8 launches per block
  const int max_jobs = 2 + 8 * m.nb + 1 + 6;
  std::vector<PackJob> jobs_v((size_t)max_jobs);
  PackJob* jobs = jobs_v.data();
  int nj = 0;
  auto conv = [&](const void* wsrc, void* wf, void* wd, int Cout, int Cin, int Cinp) {
    PackJob j; memset(&j, 0, sizeof(j));
    j.kind = KB_PACK_CONV; j.dtype = dtype; j.s0 = (const float*)wsrc; j.d0 = wf; j.d1 = wd; j.n0 = Cout; j.n1 = Cin; j.n2 = Cinp;
    j.count = (long long)Cout * 9 * Cinp;
    jobs[nj++] = j;
  };
  auto bn = [&](int layer, int pw, int bbase, int C) {
    PackJob j; memset(&j, 0, sizeof(j));
    j.kind = KB_PACK_BN; j.s0 = (const float*)params[pw]; j.s1 = (const float*)params[pw + 1];
    j.s2 = (const float*)buffers[bbase]; j.s3 = (const float*)buffers[bbase + 1];
    j.d0 = w.bn_a(layer); j.d1 = w.bn_b(layer); j.count = C; j.eps = kBnEps;
    jobs[nj++] = j;
  };
  auto lin = [&](const void* wsrc, void* out, int N, int K, int Np, int Kp) {
    PackJob j; memset(&j, 0, sizeof(j));
    j.kind = KB_PACK_LINEAR; j.s0 = (const float*)wsrc; j.d0 = out; j.n0 = N; j.n1 = K; j.n2 = Np; j.n3 = Kp;
    j.count = (long long)Np * Kp;
    jobs[nj++] = j;
  };
  conv(params[0], w.stem_wf, nullptr, m.C, m.C0, m.C0p);
  bn(0, 1, 0, m.C);
  for (int i = 0; i < m.nb; ++i) {
    conv(params[pi_blk(i, 0)], w.wf(i, 0), w.wd(i, 0), m.C, m.C, m.C);
    conv(params[pi_blk(i, 3)], w.wf(i, 1), w.wd(i, 1), m.C, m.C, m.C);
    bn(1 + 2 * i, pi_blk(i, 1), bi_blk(i, 0), m.C);
    bn(2 + 2 * i, pi_blk(i, 4), bi_blk(i, 3), m.C);
  }
  bn(2 * m.nb + 1, pi_head(m, 1), bi_pol(m, 0), m.Pc);
  if (dtype == KB_BF16 && m.C % 64 == 0) {
    const int K3 = 3 * m.C;
    for (int i = 0; i < m.nb; ++i) {
      lin(params[pi_blk(i, 6)], w.lin_blk(i, w.o_g1), m.G, K3, w.Gp, K3);
      lin(params[pi_blk(i, 8)], w.lin_blk(i, w.o_g2), m.C, m.G, w.Cp, w.Gk);
      lin(params[pi_blk(i, 10)], w.lin_blk(i, w.o_s1), m.S, m.C, w.Sp, m.C);
      lin(params[pi_blk(i, 12)], w.lin_blk(i, w.o_s2), 2 * m.C, m.S, w.C2p, w.Sk);
    }
    lin(params[pi_head(m, 0)], w.lin_head(w.o_p1), m.Pc, m.C, w.Pp, m.C);
    lin(params[pi_head(m, 3)], w.lin_head(w.o_p2), 139, m.Pc, 256, w.Pk);
    lin(params[pi_head(m, 5)], w.lin_head(w.o_v1), m.V, K3, w.Vp, K3);
    lin(params[pi_head(m, 7)], w.lin_head(w.o_v2), 3, m.V, 128, w.Vk);
    lin(params[pi_head(m, 9)], w.lin_head(w.o_c1), m.S2, K3, w.Qp, K3);
    lin(params[pi_head(m, 11)], w.lin_head(w.o_c2), 1, m.S2, 128, w.Qk);
  }
  return kbk_pack_batched(jobs, nj, st);
}

extern "C" int kb_seresnet_forward(const kb_seresnet_desc* d, const void* const* params, void* const* buffers,
                                   float* new_stats, const void* wpack, const float* obs, int B, int training, int dtype, void* workspace,
                                   long long ws_bytes, void* policy_out, long long policy_pitch, float* value_out,
                                   float* score_out, int use_tc, int num_sms, cudaStream_t st) {
  return kb_seresnet_forward_sync(d, params, buffers, new_stats, wpack, obs, B, training, dtype, workspace, ws_bytes, policy_out,
                                  policy_pitch, value_out, score_out, use_tc, num_sms, nullptr, nullptr, 1, st);
}

// SyncBatchNorm variant (reference katago_loop.py:494-508 wraps the model in SyncBatchNorm + DDP by default): `hook`
// all-reduces (sum) the per-channel double sums across `world` ranks between each convolution and its BatchNorm
// finalize, so every rank normalises with the statistics of the GLOBAL batch. hook == NULL / world == 1: local.
extern "C" int kb_seresnet_forward_sync(const kb_seresnet_desc* d, const void* const* params, void* const* buffers,
                                        float* new_stats, const void* wpack, const float* obs, int B, int training, int dtype,
                                        void* workspace, long long ws_bytes, void* policy_out, long long policy_pitch,
                                        float* value_out, float* score_out, int use_tc, int num_sms, kb_allreduce_hook hook,
                                        void* hook_user, int world, cudaStream_t st) {
  KB_CHECK_ARG(world >= 1, "world must be >= 1");
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  KB_CHECK_ARG(B >= 1, "batch must be >= 1");
  KB_CHECK_ARG(policy_pitch >= 81LL * 139, "policy pitch %lld < 11259", policy_pitch);
  KB_CHECK_ARG(params && buffers && wpack && obs && workspace && policy_out && value_out && score_out, "null pointer");
  const Dims m = make_dims(d, B, dtype);
  const WPack wp = make_wpack(m, const_cast<void*>(wpack));
  BlockWs* blks = (BlockWs*)alloca(sizeof(BlockWs) * (m.nb > 0 ? m.nb : 1));
  Ws w;
  carve(m, workspace, training, w, blks);
  KB_CHECK_ARG((size_t)ws_bytes >= w.total, "workspace too small: %lld < %zu", ws_bytes, w.total);
  const int C = m.C;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto BUF = [&](int i) { return (float*)buffers[i]; };
  const double count = (double)m.M * world;
  const int LP = 2 * m.nb + 1;  // policy BN layer id
  // training-mode BatchNorm: new running statistics go to `new_stats` ([layer][2][Cmax]) when given
  // (functional variant for autograd frameworks) or in place into `buffers` (num_batches_tracked too).
  auto bn_fin = [&](int layer, int pw, int bbase, int Cl) {
    float* rm_out = new_stats ? new_stats + (size_t)layer * 2 * w.Cmax : BUF(bbase);
    float* rv_out = new_stats ? new_stats + (size_t)layer * 2 * w.Cmax + w.Cmax : BUF(bbase + 1);
    long long* nbt = new_stats ? nullptr : (long long*)buffers[bbase + 2];
    if (hook && world > 1) {
      const int hr = hook(hook_user, w.dsums, 2LL * Cl, (kb_stream_t)st);
      if (hr != 0) { kb_set_error("BatchNorm all-reduce hook failed (rc=%d)", hr); return KB_ERR_INVALID; }
    }
    return kbk_bn_finalize(w.dsums, count, P(pw), P(pw + 1), BUF(bbase), BUF(bbase + 1), rm_out, rv_out, nbt, kBnMomentum,
                           kBnEps, Cl, w.bn_a(layer), w.bn_b(layer), w.bn_mean(layer), w.bn_invstd(layer), st);
  };

  KB_TRY(kbk_fill_zero(w.dsums, 2 * (size_t)w.Cmax * sizeof(double), st));
  KB_TRY(kbk_pack_obs(obs, w.obs_p, B, m.C0, m.C0p, dtype, st));

  // ---- stem: x0 = relu(bn(conv(obs))), plus its global-pool statistics ----
  void* x_cur; float* pool_cur;
  if (training) {
    ConvEpi e = epi_base();
    e.ch_sum = w.dsums; e.ch_sumsq = w.dsums + C;
    KB_TRY(conv3x3(m, w.obs_p, wp.stem_wf, w.z0, m.C0p, C, e, use_tc, num_sms, st));
    KB_TRY(bn_fin(0, 1, 0, C));
    ApplyArgs a; memset(&a, 0, sizeof(a));
    a.z = w.z0; a.a = w.bn_a(0); a.b = w.bn_b(0); a.out = w.x0; a.pool = w.pool(m, 0); a.ties = w.ties(m, 0); a.B = B; a.C = C; a.dtype = dtype;
    KB_TRY(kbk_apply(a, st));
    x_cur = w.x0; pool_cur = w.pool(m, 0);
  } else {
    ConvEpi e = epi_base();
    e.scale = wp.bn_a(0); e.shift = wp.bn_b(0); e.relu = 1; e.pool = w.epool_a;
    KB_TRY(conv3x3(m, w.obs_p, wp.stem_wf, w.ea, m.C0p, C, e, use_tc, num_sms, st));
    x_cur = w.ea; pool_cur = w.epool_a;
  }

  const bool ltc = lin_tc(m, use_tc);
  // SE MLP + scale/shift + residual + ReLU + pool statistics as one TMA-bulk-staged kernel (se_apply.cu)
  static int no_se_apply = -1;
  if (no_se_apply < 0) { const char* e = getenv("KB_NO_SE_APPLY"); no_se_apply = (e && e[0] == '1') ? 1 : 0; }
  const bool fuse_se = ltc && !no_se_apply && kbk_se_apply_supported(C, m.S);
  // evaluation: the whole block tail (SE MLP, scale / shift, residual, ReLU, pool statistics) rides in conv2's epilogue
  static int fused_tail_env = -1;
  if (fused_tail_env < 0) { const char* e = getenv("KB_FUSED_TAIL"); fused_tail_env = (e && e[0] == '0') ? 0 : 1; }
  const bool fuse_tail = !training && ltc && fused_tail_env && B >= 3 && kbk_conv3x3_se_tail_supported(C, C, m.S, dtype);
  static int fused_gfc_env = -1;
  if (fused_gfc_env < 0) { const char* e = getenv("KB_FUSED_GFC"); fused_gfc_env = (e && e[0] == '0') ? 0 : 1; }
  const bool fused_gfc = fused_gfc_env && kbk_gpool_mlp_tc_supported(C, m.G);
  bool pool_bf_ready = false;  // the producing apply kernel also writes the bf16 copy of the pool statistics
  // ---- residual tower ----
  for (int i = 0; i < m.nb; ++i) {
    BlockWs& bw = training ? blks[i] : blks[0];
    const int l1 = 1 + 2 * i, l2 = 2 + 2 * i;
    // global-pool bias from the block INPUT: g = W2 relu(W1 pool + b1) + b2   (se_resnet.py:73-78)
    if (ltc && fused_gfc) {
      // both Linear layers in one tcgen05 kernel, the hidden layer never leaves the SM (gpool_mlp_tc.cu)
      if (!pool_bf_ready) KB_TRY(kbk_cast_rows_bf16(pool_cur, w.pool_bf, B, 3 * C, 3 * C, st));
      KB_TRY(kbk_gpool_mlp_tc(w.pool_bf, B, 3 * C, wp.lin_blk(i, wp.o_g1), P(pi_blk(i, 7)), wp.lin_blk(i, wp.o_g2), P(pi_blk(i, 9)),
                              training ? bw.gh : nullptr, w.g, num_sms, st));
    } else if (ltc) {
      if (!pool_bf_ready) KB_TRY(kbk_cast_rows_bf16(pool_cur, w.pool_bf, B, 3 * C, 3 * C, st));
      KB_TRY(kbk_linear_tc(w.pool_bf, B, 3 * C, wp.lin_blk(i, wp.o_g1), m.G, wp.Gp, nullptr, P(pi_blk(i, 7)), 1, bw.gh, m.G,
                           w.gh_bf, wp.Gk, wp.Gk, 0, 0, num_sms, st));
      KB_TRY(kbk_linear_tc(w.gh_bf, B, wp.Gk, wp.lin_blk(i, wp.o_g2), C, wp.Cp, nullptr, P(pi_blk(i, 9)), 0, w.g, C, nullptr, 0, 0,
                           0, 0, num_sms, st));
    } else {
      KB_TRY(linear_fwd(pool_cur, KB_F32, 3 * C, B, 3 * C, P(pi_blk(i, 6)), m.G, P(pi_blk(i, 7)), 1, bw.gh, KB_F32, m.G, st));
      KB_TRY(linear_fwd(bw.gh, KB_F32, m.G, B, m.G, P(pi_blk(i, 8)), C, P(pi_blk(i, 9)), 0, w.g, KB_F32, C, st));
    }
    const void* z2; void* xout; float* pool_next; const float* se_in;
    if (training) {
      ConvEpi e = epi_base();
      e.ch_sum = w.dsums; e.ch_sumsq = w.dsums + C;
      KB_TRY(conv3x3(m, x_cur, wp.wf(i, 0), bw.z1, C, C, e, use_tc, num_sms, st));
      KB_TRY(bn_fin(l1, pi_blk(i, 1), bi_blk(i, 0), C));
      ApplyArgs a; memset(&a, 0, sizeof(a));
      a.z = bw.z1; a.a = w.bn_a(l1); a.b = w.bn_b(l1); a.gbias = w.g; a.out = bw.a1; a.B = B; a.C = C; a.dtype = dtype;
      KB_TRY(kbk_apply(a, st));
      ConvEpi e2 = epi_base();
      e2.ch_sum = w.dsums; e2.ch_sumsq = w.dsums + C; e2.board_sum = bw.bmean2; e2.board_scale = 1.f / 81.f;
      KB_TRY(conv3x3(m, bw.a1, wp.wf(i, 1), bw.z2, C, C, e2, use_tc, num_sms, st));
      KB_TRY(bn_fin(l2, pi_blk(i, 4), bi_blk(i, 3), C));
      // SE squeeze input: mean_p(bn2(z2)) = a2 * mean_p(z2) + b2
      if (!fuse_se) KB_TRY(kbk_affine_rows(bw.bmean2, w.bn_a(l2), w.bn_b(l2), bw.se_in, ltc ? w.sein_bf : nullptr, B, C, st));
      z2 = bw.z2; xout = bw.xout; pool_next = w.pool(m, i + 1); se_in = bw.se_in;
    } else {
      ConvEpi e = epi_base();
      e.scale = wp.bn_a(l1); e.shift = wp.bn_b(l1); e.relu = 1; e.gbias = w.g;
      KB_TRY(conv3x3(m, x_cur, wp.wf(i, 0), w.ey1, C, C, e, use_tc, num_sms, st));
      if (fuse_tail) {
        void* xo = (x_cur == w.ea) ? w.eb : w.ea;
        float* pn = (pool_cur == w.epool_a) ? w.epool_b : w.epool_a;
        ConvEpi e2 = epi_base();
        e2.scale = wp.bn_a(l2); e2.shift = wp.bn_b(l2); e2.res = x_cur;
        e2.se_w1 = P(pi_blk(i, 10)); e2.se_b1 = P(pi_blk(i, 11)); e2.se_w2 = P(pi_blk(i, 12)); e2.se_b2 = P(pi_blk(i, 13));
        e2.pool = pn; e2.pool_bf = w.pool_bf;
        KB_TRY(kbk_conv3x3_tc_mode(w.ey1, wp.wf(i, 1), xo, B, C, C, e2, num_sms, 2, st));
        pool_bf_ready = true;
        x_cur = xo; pool_cur = pn;
        continue;
      }
      ConvEpi e2 = epi_base();
      e2.scale = wp.bn_a(l2); e2.shift = wp.bn_b(l2); e2.board_sum = bw.bmean2; e2.board_scale = 1.f / 81.f;
      e2.board_bf = (ltc && !fuse_se) ? w.sein_bf : nullptr;
      KB_TRY(conv3x3(m, w.ey1, wp.wf(i, 1), w.ey2, C, C, e2, use_tc, num_sms, st));
      z2 = w.ey2; xout = (x_cur == w.ea) ? w.eb : w.ea; pool_next = (pool_cur == w.epool_a) ? w.epool_b : w.epool_a;
      se_in = bw.bmean2;  // BN affine already applied in the conv epilogue
    }
    if (fuse_se) {
      SeApplyArgs sa; memset(&sa, 0, sizeof(sa));
      sa.z = (const bf16*)z2; sa.res = (const bf16*)x_cur; sa.out = (bf16*)xout;
      sa.a = training ? w.bn_a(l2) : nullptr; sa.b = training ? w.bn_b(l2) : nullptr;
      sa.bmean = bw.bmean2;
      sa.w1 = P(pi_blk(i, 10)); sa.b1 = P(pi_blk(i, 11)); sa.w2 = P(pi_blk(i, 12)); sa.b2 = P(pi_blk(i, 13));
      sa.se_in_out = training ? bw.se_in : nullptr; sa.seh_out = training ? bw.seh : nullptr; sa.se_out = bw.se; sa.se_raw = training ? 1 : 0;
      sa.pool = pool_next; sa.pool_bf = (bf16*)w.pool_bf; sa.ties = training ? w.ties(m, i + 1) : nullptr;
      sa.B = B; sa.C = C; sa.S = m.S;
      KB_TRY(kbk_se_apply(sa, num_sms, st));
      pool_bf_ready = true;
      x_cur = xout; pool_cur = pool_next;
      continue;
    }
    // SE excite: (scale, shift) = W2 relu(W1 se_in + b1) + b2   (se_resnet.py:83-86)
    if (ltc) {
      KB_TRY(kbk_linear_tc(w.sein_bf, B, C, wp.lin_blk(i, wp.o_s1), m.S, wp.Sp, nullptr, P(pi_blk(i, 11)), 1, bw.seh, m.S,
                           w.seh_bf, wp.Sk, wp.Sk, 0, 0, num_sms, st));
      KB_TRY(kbk_linear_tc(w.seh_bf, B, wp.Sk, wp.lin_blk(i, wp.o_s2), 2 * C, wp.C2p, nullptr, P(pi_blk(i, 13)), 0, bw.se, 2 * C,
                           nullptr, 0, 0, 0, 0, num_sms, st));
    } else {
      KB_TRY(linear_fwd(se_in, KB_F32, C, B, C, P(pi_blk(i, 10)), m.S, P(pi_blk(i, 11)), 1, bw.seh, KB_F32, m.S, st));
      KB_TRY(linear_fwd(bw.seh, KB_F32, m.S, B, m.S, P(pi_blk(i, 12)), 2 * C, P(pi_blk(i, 13)), 0, bw.se, KB_F32, 2 * C, st));
    }
    // x' = relu(bn2(z2) * sigmoid(scale) + shift + x), plus the pool statistics of x' for the next consumer
    ApplyArgs a; memset(&a, 0, sizeof(a));
    a.z = z2; a.a = training ? w.bn_a(l2) : nullptr; a.b = training ? w.bn_b(l2) : nullptr; a.se = bw.se; a.res = x_cur;
    a.out = xout; a.pool = pool_next; a.ties = training ? w.ties(m, i + 1) : nullptr; a.B = B; a.C = C; a.dtype = dtype;
    a.pool_bf = ltc ? w.pool_bf : nullptr;  // consumed by the next block's global_fc GEMM
    KB_TRY(kbk_apply(a, st));
    pool_bf_ready = ltc;
    x_cur = xout; pool_cur = pool_next;
  }

  // ---- policy head: 1x1 conv -> BN -> ReLU -> 1x1 conv + bias, written NHWC into the padded logits buffer ----
  const int Pc = m.Pc, M = (int)m.M;
  if (ltc) {
    const float *pa = wp.bn_a(LP), *pb = wp.bn_b(LP);
    if (training) {
      KB_TRY(kbk_linear_tc(x_cur, m.M, C, wp.lin_head(wp.o_p1), Pc, wp.Pp, nullptr, nullptr, 0, w.p1raw, Pc, nullptr, 0, 0, 0, 0,
                           num_sms, st));
      KB_TRY(kbk_rows_stats(w.p1raw, m.M, Pc, w.dsums, st));
      KB_TRY(bn_fin(LP, pi_head(m, 1), bi_pol(m, 0), Pc));
      ApplyArgs a; memset(&a, 0, sizeof(a));
      a.z = w.p1raw; a.a = w.bn_a(LP); a.b = w.bn_b(LP); a.out = w.p1act; a.B = B; a.C = Pc; a.dtype = KB_F32;
      KB_TRY(kbk_apply(a, st));
      KB_TRY(kbk_cast_rows_bf16(w.p1act, w.p1act_bf, m.M, Pc, wp.Pk, st));
    } else {
      // eval: folded BatchNorm + ReLU ride in the GEMM epilogue
      KB_TRY(kbk_linear_tc(x_cur, m.M, C, wp.lin_head(wp.o_p1), Pc, wp.Pp, pa, pb, 1, nullptr, 0, w.p1act_bf, wp.Pk, wp.Pk, 0, 0,
                           num_sms, st));
    }
    KB_TRY(kbk_linear_tc(w.p1act_bf, m.M, wp.Pk, wp.lin_head(wp.o_p2), 139, 256, nullptr, P(pi_head(m, 4)), 0, nullptr, 0,
                         policy_out, 139, 139, 81, policy_pitch, num_sms, st));
    // value / score heads stay fp32 end to end (4 tiny GEMMs on the fp32 pool statistics): their logits are the
    // smallest-magnitude outputs of the network and carry the 2e-2 bf16 parity bar with the least headroom
    KB_TRY(linear_fwd(pool_cur, KB_F32, 3 * C, B, 3 * C, P(pi_head(m, 5)), m.V, P(pi_head(m, 6)), 1, w.vh, KB_F32, m.V, st));
    KB_TRY(linear_fwd(w.vh, KB_F32, m.V, B, m.V, P(pi_head(m, 7)), 3, P(pi_head(m, 8)), 0, value_out, KB_F32, 3, st));
    KB_TRY(linear_fwd(pool_cur, KB_F32, 3 * C, B, 3 * C, P(pi_head(m, 9)), m.S2, P(pi_head(m, 10)), 1, w.sh, KB_F32, m.S2, st));
    KB_TRY(linear_fwd(w.sh, KB_F32, m.S2, B, m.S2, P(pi_head(m, 11)), 1, P(pi_head(m, 12)), 0, score_out, KB_F32, 1, st));
    return KB_OK;
  }
  KB_TRY(linear_fwd(x_cur, dtype, C, M, C, P(pi_head(m, 0)), Pc, nullptr, 0, w.p1raw, KB_F32, Pc, st));
  const float *pa, *pb;
  if (training) {
    KB_TRY(kbk_rows_stats(w.p1raw, m.M, Pc, w.dsums, st));
    KB_TRY(bn_fin(LP, pi_head(m, 1), bi_pol(m, 0), Pc));
    pa = w.bn_a(LP); pb = w.bn_b(LP);
  } else {
    pa = wp.bn_a(LP); pb = wp.bn_b(LP);
  }
  {
    ApplyArgs a; memset(&a, 0, sizeof(a));
    a.z = w.p1raw; a.a = pa; a.b = pb; a.out = w.p1act; a.B = B; a.C = Pc; a.dtype = KB_F32;
    KB_TRY(kbk_apply(a, st));
    GemmArgs g = gemm_base();
    g.A = w.p1act; g.a_dtype = KB_F32; g.lda = Pc;
    g.B = P(pi_head(m, 3)); g.b_dtype = KB_F32; g.ldb = Pc; g.transB = 1;
    g.C = policy_out; g.c_dtype = dtype; g.ldc = 139; g.c_group_rows = 81; g.c_group_pitch = policy_pitch;
    g.bias = P(pi_head(m, 4)); g.M = M; g.N = 139; g.K = Pc;
    KB_TRY(kbk_gemm(g, st));
  }
  // ---- value / score heads on the shared global pool of the trunk output ----
  KB_TRY(linear_fwd(pool_cur, KB_F32, 3 * C, B, 3 * C, P(pi_head(m, 5)), m.V, P(pi_head(m, 6)), 1, w.vh, KB_F32, m.V, st));
  KB_TRY(linear_fwd(w.vh, KB_F32, m.V, B, m.V, P(pi_head(m, 7)), 3, P(pi_head(m, 8)), 0, value_out, KB_F32, 3, st));
  KB_TRY(linear_fwd(pool_cur, KB_F32, 3 * C, B, 3 * C, P(pi_head(m, 9)), m.S2, P(pi_head(m, 10)), 1, w.sh, KB_F32, m.S2, st));
  KB_TRY(linear_fwd(w.sh, KB_F32, m.S2, B, m.S2, P(pi_head(m, 11)), 1, P(pi_head(m, 12)), 0, score_out, KB_F32, 1, st));
  return KB_OK;
}

// Backward of the training-mode forward that filled `workspace`. `grads` must be pre-zeroed fp32
// buffers shaped like the parameters (same table order).
extern "C" int kb_seresnet_backward(const kb_seresnet_desc* d, const void* const* params, const void* wpack, int B,
                                    int dtype, void* workspace, long long ws_bytes, const void* dpolicy,
                                    long long policy_pitch, const float* dvalue, const float* dscore,
                                    void* const* grads, int use_tc, int num_sms, cudaStream_t st) {
  return kb_seresnet_backward_sync(d, params, wpack, B, dtype, workspace, ws_bytes, dpolicy, policy_pitch, dvalue, dscore, grads,
                                   use_tc, num_sms, nullptr, nullptr, 1, st);
}

// Backward of kb_seresnet_forward_sync: the BatchNorm-backward sums (sum dz, sum dz*z) are all-reduced through `hook`
// for the data gradient; dgamma / dbeta keep this rank's share (gradient averaging adds the ranks up afterwards).
extern "C" int kb_seresnet_backward_sync(const kb_seresnet_desc* d, const void* const* params, const void* wpack, int B,
                                         int dtype, void* workspace, long long ws_bytes, const void* dpolicy,
                                         long long policy_pitch, const float* dvalue, const float* dscore, void* const* grads,
                                         int use_tc, int num_sms, kb_allreduce_hook hook, void* hook_user, int world,
                                         cudaStream_t st) {
  KB_CHECK_ARG(world >= 1, "world must be >= 1");
  KB_TRY(check_desc(d));
  KB_CHECK_ARG(dtype == KB_F32 || dtype == KB_BF16, "bad dtype");
  KB_CHECK_ARG(B >= 1 && policy_pitch >= 81LL * 139, "bad shape");
  KB_CHECK_ARG(params && wpack && workspace && dpolicy && dvalue && dscore && grads, "null pointer");
  const Dims m = make_dims(d, B, dtype);
  const WPack wp = make_wpack(m, const_cast<void*>(wpack));
  BlockWs* blks = (BlockWs*)alloca(sizeof(BlockWs) * (m.nb > 0 ? m.nb : 1));
  Ws w;
  carve(m, workspace, 1, w, blks);
  KB_CHECK_ARG((size_t)ws_bytes >= w.total, "workspace too small: %lld < %zu", ws_bytes, w.total);
  const int C = m.C, Pc = m.Pc, M = (int)m.M;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto G = [&](int i) { return (float*)grads[i]; };
  const double count = (double)m.M * world;
  const int LP = 2 * m.nb + 1;
  float *k1 = w.k123, *k2 = w.k123 + w.Cmax, *k3 = w.k123 + 2 * w.Cmax;
  const bool sync = hook != nullptr && world > 1;
  Tf32Scope tf32_scope(dtype == KB_BF16 && use_tc ? 1 : 0);  // MLP / head backward GEMMs on the tensor cores (AMP path only)
  // BatchNorm backward finalize; under SyncBatchNorm the sums are all-reduced first (local copy kept for dgamma / dbeta)
  auto bn_bwd_fin = [&](int pw, int layer, float* dgamma, float* dbeta, int Cl) -> int {
    if (sync) {
      KB_CUDA_CHECK(cudaMemcpyAsync(w.dsums_local, w.dsums, 2 * (size_t)Cl * sizeof(double), cudaMemcpyDeviceToDevice, st));
      const int hr = hook(hook_user, w.dsums, 2LL * Cl, (kb_stream_t)st);
      if (hr != 0) { kb_set_error("BatchNorm all-reduce hook failed (rc=%d)", hr); return KB_ERR_INVALID; }
    }
    return kbk_bn_bwd_finalize(w.dsums, count, P(pw), w.bn_mean(layer), w.bn_invstd(layer), k1, k2, k3, dgamma, dbeta, Cl, st,
                               sync ? w.dsums_local : nullptr);
  };
  const void* x_last = m.nb > 0 ? blks[m.nb - 1].xout : w.x0;
  const float* pool_f = w.pool(m, m.nb);

  // ---- policy head ----
  {
    GemmArgs g = gemm_base();  // dWp2[139][Pc] += dlogits^T * p1act
    g.A = dpolicy; g.a_dtype = dtype; g.lda = 139; g.transA = 1; g.a_group_rows = 81; g.a_group_pitch = policy_pitch;
    g.B = w.p1act; g.b_dtype = KB_F32; g.ldb = Pc; g.transB = 0;
    g.C = G(pi_head(m, 3)); g.c_dtype = KB_F32; g.ldc = Pc; g.M = 139; g.N = Pc; g.K = M;
    g.splitk = kb_ceil_div(M, 2048) < 2 ? 2 : kb_ceil_div(M, 2048);
    KB_TRY(kbk_gemm(g, st));
    KB_TRY(kbk_colsum(dpolicy, dtype, 139, 81, policy_pitch, M, 139, G(pi_head(m, 4)), st));
    GemmArgs h = gemm_base();  // dp1[M][Pc] = dlogits * Wp2
    h.A = dpolicy; h.a_dtype = dtype; h.lda = 139; h.a_group_rows = 81; h.a_group_pitch = policy_pitch;
    h.B = P(pi_head(m, 3)); h.b_dtype = KB_F32; h.ldb = Pc; h.transB = 0;
    h.C = w.dp1; h.c_dtype = KB_F32; h.ldc = Pc; h.M = M; h.N = Pc; h.K = 139;
    KB_TRY(kbk_gemm(h, st));
  }
  KB_TRY(kbk_relu_bwd_stats_f32(w.dp1, w.p1act, w.p1raw, m.M, Pc, w.dsums, st));
  KB_TRY(bn_bwd_fin(pi_head(m, 1), LP, G(pi_head(m, 1)), G(pi_head(m, 2)), Pc));
  KB_TRY(kbk_bn_bwd_apply(w.dp1, w.p1raw, k1, k2, k3, m.M, Pc, KB_F32, st));
  KB_TRY(linear_bwd_w(w.dp1, KB_F32, Pc, x_last, dtype, C, M, Pc, C, G(pi_head(m, 0)), nullptr, st));
  KB_TRY(linear_bwd_x(w.dp1, KB_F32, Pc, M, Pc, P(pi_head(m, 0)), C, w.d1, dtype, C, nullptr, 0, 0, st));
  // ---- value / score heads ----
  KB_TRY(linear_bwd_w(dvalue, KB_F32, 3, w.vh, KB_F32, m.V, B, 3, m.V, G(pi_head(m, 7)), G(pi_head(m, 8)), st));
  KB_TRY(linear_bwd_x(dvalue, KB_F32, 3, B, 3, P(pi_head(m, 7)), m.V, w.dvh, KB_F32, m.V, w.vh, m.V, 0, st));
  KB_TRY(linear_bwd_w(w.dvh, KB_F32, m.V, pool_f, KB_F32, 3 * C, B, m.V, 3 * C, G(pi_head(m, 5)), G(pi_head(m, 6)), st));
  KB_TRY(linear_bwd_x(w.dvh, KB_F32, m.V, B, m.V, P(pi_head(m, 5)), 3 * C, w.dpool, KB_F32, 3 * C, nullptr, 0, 0, st));
  KB_TRY(linear_bwd_w(dscore, KB_F32, 1, w.sh, KB_F32, m.S2, B, 1, m.S2, G(pi_head(m, 11)), G(pi_head(m, 12)), st));
  KB_TRY(linear_bwd_x(dscore, KB_F32, 1, B, 1, P(pi_head(m, 11)), m.S2, w.dsh, KB_F32, m.S2, w.sh, m.S2, 0, st));
  KB_TRY(linear_bwd_w(w.dsh, KB_F32, m.S2, pool_f, KB_F32, 3 * C, B, m.S2, 3 * C, G(pi_head(m, 9)), G(pi_head(m, 10)), st));
  KB_TRY(linear_bwd_x(w.dsh, KB_F32, m.S2, B, m.S2, P(pi_head(m, 9)), 3 * C, w.dpool, KB_F32, 3 * C, nullptr, 0, 1, st));
  // Weight-gradient convolutions on the side stream (see SideStream). Events per block: [0] dz2 ready (main -> side),
  // [1] dz1 ready (main -> side), [2] conv2 weight gradient done, [3] conv1 weight gradient done (side -> main).
  const bool overlap = m.nb > 0 && bwd_overlap_enabled(st);
  // Each weight gradient is released just BEFORE its data-gradient conv (measured: 214 ms per 8192-sample step against
  // 220 ms when released after it and 224 ms on one stream); KB_BWD_WG_FIRST=0 selects the other order.
  static int wg_first = -1;
  if (wg_first < 0) { const char* e = getenv("KB_BWD_WG_FIRST"); wg_first = (e && e[0] == '0') ? 0 : 1; }
  SideStream& side = side_stream_tls();
  if (overlap) {
    int dev = 0;
    KB_CUDA_CHECK(cudaGetDevice(&dev));
    KB_TRY(side.ensure(dev, 6 * m.nb + 1));
  }
  cudaStream_t wst = overlap ? side.s : st;
  // The global-pool-bias MLP backward of a block (two grouped launches of small TF32 GEMMs, 2 x 68 us at 8192 samples and
  // 2 x 25 us at 1024: latency- / L2-bound) has one consumer on the main chain, the block's dx pass (dpool), so it CAN run
  // on a second side stream under the HBM-bound dz1 pass and the conv1 data gradient: events [4] dg ready (main ->
  // gst), [5] MLP backward done (gst -> main, before the dx pass). Measured on B200 (same box, interleaved): SLOWER —
  // 197.5 / 199.3 ms against 192.7 / 193.9 ms per 8192-sample step, 29.6 / 29.8 against 29.0 / 29.1 ms at 1024 samples —
  // the co-running kernels take more from the dz1 pass and the convolutions than the GEMMs' own time. Off unless
  // KB_BWD_MLP_SIDE=1.
  static int mlp_side_env = -1;
  if (mlp_side_env < 0) { const char* e = getenv("KB_BWD_MLP_SIDE"); mlp_side_env = (e && e[0] == '1') ? 1 : 0; }
  const bool mlp_side = overlap && mlp_side_env;
  cudaStream_t gst = mlp_side ? side.s2 : st;
  auto ev = [&](int blk, int k) { return side.ev[6 * blk + k]; };
  auto bucket_ready = [&](int bucket, int first_param, int n_params) -> int {
    if (g_bucket_hook == nullptr) return KB_OK;
    const int hr = g_bucket_hook(g_bucket_user, bucket, first_param, n_params, (kb_stream_t)st, (kb_stream_t)wst);
    if (hr != 0) { kb_set_error("gradient bucket hook failed (rc=%d)", hr); return KB_ERR_INVALID; }
    return KB_OK;
  };
  KB_TRY(bucket_ready(m.nb, pi_head(m, 0), 13));   // policy / value / score head gradients are complete
  // dL/dx_last = policy path + global-pool backward
  void *cur = w.d0, *t1 = w.d1, *t2 = w.d2, *t3 = w.d3;
  {
    PassDArgs a; memset(&a, 0, sizeof(a));
    a.B = B; a.C = C; a.dtype = dtype; a.dxc = w.d1; a.x = x_last; a.pool = pool_f; a.dpool = w.dpool; a.ties = w.ties(m, m.nb); a.dx = cur;
    // stored as du = dx * [x_last > 0] together with the last block's board sums (see PassDArgs)
    a.mask_out = 1;
    if (m.nb > 0) { a.z_next = blks[m.nb - 1].z2; a.s_du = w.s_du; a.s_duz = w.s_duz; }
    KB_TRY(kbk_block_bwd_dx(a, st));
  }

  // ---- residual tower, last block first ----
  for (int i = m.nb - 1; i >= 0; --i) {
    BlockWs& bw = blks[i];
    const void* x_in = i > 0 ? blks[i - 1].xout : w.x0;
    const float* pool_in = w.pool(m, i);
    const int l1 = 1 + 2 * i, l2 = 2 + 2 * i;
    // `cur` already holds du = dx' * [x' > 0] and w.s_du / w.s_duz its per-(board, channel) sums: both were produced
    // by the consumer's data-gradient pass (PassDArgs.mask_out / z_next), so this block never re-reads x' for the mask
    if (kbk_se_mlp_bwd_supported(C, m.S)) {
      // one launch: SE MLP backward (both weight gradients, dse_in) + the BatchNorm-2 backward statistics
      KB_TRY(kbk_se_mlp_bwd(w.s_du, w.s_duz, w.bn_a(l2), w.bn_b(l2), bw.se, bw.seh, bw.se_in, bw.bmean2, P(pi_blk(i, 10)),
                            P(pi_blk(i, 12)), w.dse_in, G(pi_blk(i, 10)), G(pi_blk(i, 11)), G(pi_blk(i, 12)), G(pi_blk(i, 13)),
                            w.dsums, B, C, m.S, num_sms, st));
    } else {
      KB_TRY(kbk_se_bwd_prep(w.s_du, w.s_duz, w.bn_a(l2), w.bn_b(l2), bw.se, w.dse, B, C, st));
      // SE MLP backward
      KB_TRY(linear_bwd_w(w.dse, KB_F32, 2 * C, bw.seh, KB_F32, m.S, B, 2 * C, m.S, G(pi_blk(i, 12)), G(pi_blk(i, 13)), st));
      KB_TRY(linear_bwd_x(w.dse, KB_F32, 2 * C, B, 2 * C, P(pi_blk(i, 12)), m.S, w.dseh, KB_F32, m.S, bw.seh, m.S, 0, st));
      KB_TRY(linear_bwd_w(w.dseh, KB_F32, m.S, bw.se_in, KB_F32, C, B, m.S, C, G(pi_blk(i, 10)), G(pi_blk(i, 11)), st));
      KB_TRY(linear_bwd_x(w.dseh, KB_F32, m.S, B, m.S, P(pi_blk(i, 10)), C, w.dse_in, KB_F32, C, nullptr, 0, 0, st));
      // BN2 backward statistics from board-level sums, then dz2
      KB_TRY(kbk_bn2_bwd_sums(w.s_du, w.s_duz, bw.se, w.dse_in, bw.bmean2, B, C, w.dsums, st));
    }
    KB_TRY(bn_bwd_fin(pi_blk(i, 4), l2, G(pi_blk(i, 4)), G(pi_blk(i, 5)), C));
    PassBArgs pb; memset(&pb, 0, sizeof(pb));
    pb.B = B; pb.C = C; pb.dtype = dtype; pb.dxp = cur; pb.xp = nullptr; pb.z2 = bw.z2; pb.se = bw.se; pb.dse_in = w.dse_in;
    pb.k1 = k1; pb.k2 = k2; pb.k3 = k3; pb.dz2 = t1;
    KB_TRY(kbk_block_bwd_dz2(pb, st));
    // conv2: data gradient first (it heads the dependency chain), then the weight gradient — on the side stream it is
    // released only when the data-gradient conv has finished, so the two tensor-bound kernels never compete for the
    // SMs and the weight gradient runs under the HBM-bound passes that follow (mask/statistics, MLP backward, dz1)
    auto conv2_wgrad = [&]() -> int {
      if (overlap) { KB_CUDA_CHECK(cudaEventRecord(ev(i, 0), st)); KB_CUDA_CHECK(cudaStreamWaitEvent(wst, ev(i, 0), 0)); }
      KB_TRY(wgrad3x3(m, bw.a1, t1, G(pi_blk(i, 3)), C, C, C, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, wst));
      if (overlap) KB_CUDA_CHECK(cudaEventRecord(ev(i, 2), wst));
      return KB_OK;
    };
    if (!overlap || wg_first) KB_TRY(conv2_wgrad());
    // The fused epilogue (mask + statistics inside the conv) is latency-bound on its 2-byte mask loads for the
    // tcgen05 kernel (profiles/): there the tail runs as one vectorised pass after a plain data-gradient conv.
    const bool tc_conv = use_tc && B >= 3 && kbk_conv3x3_tc_supported(C, C, dtype);
    static int fused_env = -1;
    if (fused_env < 0) { const char* fe = getenv("KB_FUSED_DGRAD"); fused_env = (fe && fe[0] == '1') ? 1 : 0; }
    // The ReLU mask is a function of z1 alone (a1 = relu(bn1(z1)) + gbias), so the statistics pass does not write the
    // masked gradient back: the dz1 pass below recomputes the mask (KB_MASK_RECOMPUTE=0 restores the three-pass form).
    static int mask_recompute_env = -1;
    if (mask_recompute_env < 0) { const char* me = getenv("KB_MASK_RECOMPUTE"); mask_recompute_env = (me && me[0] == '0') ? 0 : 1; }
    bool mask_pending = false;
    if (tc_conv && kbk_mask_bwd_stats_supported(C) && !fused_env) {
      ConvEpi e = epi_base();
      KB_TRY(conv3x3(m, t1, wp.wd(i, 1), t2, C, C, e, use_tc, num_sms, st));
      if (overlap && !wg_first) KB_TRY(conv2_wgrad());
      if (mask_recompute_env && kbk_mask_bwd_stats_ro_supported(C) && kbk_bn_bwd_apply_masked_supported(m.M, C)) {
        KB_TRY(kbk_mask_bwd_stats_ro(t2, bw.z1, w.bn_a(l1), w.bn_b(l1), w.dg, B, C, dtype, w.dsums, st));
        mask_pending = true;
      } else {
        KB_TRY(kbk_mask_bwd_stats(t2, bw.z1, w.bn_a(l1), w.bn_b(l1), w.dg, B, C, dtype, w.dsums, st));
      }
    } else {
      ConvEpi e = epi_base();
      e.mask_src = bw.z1; e.mask_a = w.bn_a(l1); e.mask_b = w.bn_b(l1);
      e.ch_sum = w.dsums; e.ch_dot = w.dsums + C; e.board_sum = w.dg; e.board_scale = 1.f;
      KB_TRY(conv3x3(m, t1, wp.wd(i, 1), t2, C, C, e, use_tc, num_sms, st));
      if (overlap && !wg_first) KB_TRY(conv2_wgrad());
    }
    KB_TRY(bn_bwd_fin(pi_blk(i, 1), l1, G(pi_blk(i, 1)), G(pi_blk(i, 2)), C));
    // global-pool-bias MLP backward -> gradient wrt the pool statistics of the block input: four small GEMMs and two
    // bias column sums as TWO grouped launches (the members of a group are independent of one another)
    {
      if (mlp_side) { KB_CUDA_CHECK(cudaEventRecord(ev(i, 4), st)); KB_CUDA_CHECK(cudaStreamWaitEvent(gst, ev(i, 4), 0)); }
      GemmGroupScope grp(gst);
      KB_TRY(grp.status());
      KB_TRY(linear_bwd_w(w.dg, KB_F32, C, bw.gh, KB_F32, m.G, B, C, m.G, G(pi_blk(i, 8)), G(pi_blk(i, 9)), gst));
      KB_TRY(linear_bwd_x(w.dg, KB_F32, C, B, C, P(pi_blk(i, 8)), m.G, w.dgh, KB_F32, m.G, bw.gh, m.G, 0, gst));
      KB_TRY(grp.flush());   // dgh is complete before anything reads it
      KB_TRY(linear_bwd_w(w.dgh, KB_F32, m.G, pool_in, KB_F32, 3 * C, B, m.G, 3 * C, G(pi_blk(i, 6)), G(pi_blk(i, 7)), gst));
      KB_TRY(linear_bwd_x(w.dgh, KB_F32, m.G, B, m.G, P(pi_blk(i, 6)), 3 * C, w.dpool, KB_F32, 3 * C, nullptr, 0, 0, gst));
      KB_TRY(grp.close());
      if (mlp_side) KB_CUDA_CHECK(cudaEventRecord(ev(i, 5), gst));
    }
    // pass C: dz1 in place; conv1 weight + data gradients
    if (mask_pending) KB_TRY(kbk_bn_bwd_apply_masked(t2, bw.z1, k1, k2, k3, w.bn_a(l1), w.bn_b(l1), m.M, C, dtype, st));
    else KB_TRY(kbk_bn_bwd_apply(t2, bw.z1, k1, k2, k3, m.M, C, dtype, st));
    auto conv1_wgrad = [&]() -> int {  // side stream: released at this point of the main stream
      KB_CUDA_CHECK(cudaEventRecord(ev(i, 1), st));
      KB_CUDA_CHECK(cudaStreamWaitEvent(wst, ev(i, 1), 0));
      KB_TRY(wgrad3x3(m, x_in, t2, G(pi_blk(i, 0)), C, C, C, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, wst));
      KB_CUDA_CHECK(cudaEventRecord(ev(i, 3), wst));
      return KB_OK;
    };
    if (!overlap) KB_TRY(wgrad3x3(m, x_in, t2, G(pi_blk(i, 0)), C, C, C, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, st));
    else if (wg_first) KB_TRY(conv1_wgrad());
    if (overlap) KB_CUDA_CHECK(cudaStreamWaitEvent(st, ev(i, 2), 0));  // t1 (dz2) is about to be overwritten: its weight gradient must be done
    ConvEpi e1 = epi_base();
    KB_TRY(conv3x3(m, t2, wp.wd(i, 0), t1, C, C, e1, use_tc, num_sms, st));
    // conv1 weight gradient: released after the data-gradient conv, runs under pass D / the next block's SE backward
    if (overlap && !wg_first) KB_TRY(conv1_wgrad());
    // pass D: dx = dgrad + residual branch + global-pool backward
    PassDArgs pd; memset(&pd, 0, sizeof(pd));
    pd.B = B; pd.C = C; pd.dtype = dtype; pd.dxc = t1; pd.dxp = cur; pd.xp = nullptr; pd.x = x_in; pd.pool = pool_in;
    // dx goes to the fourth buffer: t2 (dz1) is still being read by this block's conv1 weight gradient on the side
    // stream. t3 was the previous block's dz1, so that block's weight gradient has to be done by now.
    pd.dpool = w.dpool; pd.ties = w.ties(m, i); pd.dx = t3;
    pd.mask_out = 1;  // hand du (masked by the producer's ReLU) to block i-1 / the stem
    if (i > 0) { pd.z_next = blks[i - 1].z2; pd.s_du = w.s_du; pd.s_duz = w.s_duz; }
    if (overlap && i + 1 < m.nb) KB_CUDA_CHECK(cudaStreamWaitEvent(st, ev(i + 1, 3), 0));
    if (mlp_side) KB_CUDA_CHECK(cudaStreamWaitEvent(st, ev(i, 5), 0));   // dpool (and the MLP's four gradients) are complete
    KB_TRY(kbk_block_bwd_dx(pd, st));
    void* nc = t3; t3 = t2; t2 = t1; t1 = cur; cur = nc;
    KB_TRY(bucket_ready(i, pi_blk(i, 0), 14));     // all 14 gradients of block i are enqueued (weight gradients on `wst`)
  }
  if (overlap) {  // join: the stem's weight gradient reuses the partial-tile workspace, and the caller sees one stream
    KB_CUDA_CHECK(cudaEventRecord(side.ev[6 * m.nb], wst));
    KB_CUDA_CHECK(cudaStreamWaitEvent(st, side.ev[6 * m.nb], 0));
  }

  // ---- stem ----
  KB_TRY(kbk_relu_bwd_stats(cur, w.x0, w.z0, t1, m.M, C, dtype, w.dsums, st));
  KB_TRY(bn_bwd_fin(1, 0, G(1), G(2), C));
  KB_TRY(kbk_bn_bwd_apply(t1, w.z0, k1, k2, k3, m.M, C, dtype, st));
  KB_TRY(wgrad3x3(m, w.obs_p, t1, G(0), m.C0p, C, m.C0, use_tc, num_sms, w.wg_ws, w.wg_ws_bytes, st));
  return KB_OK;
}
