// conv_epilogue.cuh — the fused epilogue shared by the SIMT and tcgen05 3x3 convolution kernels.
//
// Both kernels give every epilogue thread ONE output channel c and a run of output pixels of up to
// NB whole boards (81 pixels each). That makes every per-(board, channel) and per-channel reduction
// of the SE-ResNet block a thread-local register reduction (reference se_resnet.py:68-98):
//   * BatchNorm batch statistics  (sum, sum of squares per channel)        -> double atomics
//   * SE squeeze                  (mean over the 81 pixels of a board)     -> board_sum
//   * global-pool statistics      (mean / max / population std per board)  -> pool
//   * BN-backward statistics      (sum dz, sum dz*z per channel), ReLU mask and gpool-bias grad
// plus the elementwise tail: per-channel affine (folded eval BatchNorm), ReLU, per-(board,channel)
// bias (the global-pool bias, added after the ReLU).
#pragma once
#include "kb_common.cuh"

struct ConvEpi {
  const float* scale;     // [Cout] or null: v = acc*scale + shift
  const float* shift;     // [Cout]
  int relu;               // v = max(v, 0)
  const float* gbias;     // [B][Cout] or null: v += gbias[b][c]   (after the ReLU)
  const void* mask_src;   // [B][81][Cout] (activation dtype) or null: v = (mask_src*mask_a+mask_b > 0) ? v : 0
  const float* mask_a;    // [Cout]
  const float* mask_b;    // [Cout]
  double* ch_sum;         // [Cout] or null: += sum of stored v
  double* ch_sumsq;       // [Cout] or null: += sum of stored v^2
  double* ch_dot;         // [Cout] or null: += sum of stored v * mask_src
  float* board_sum;       // [B][Cout] or null: board_scale * sum over the board of v (taken BEFORE the mask when mask_src is set)
  float board_scale;      // e.g. 1/81 for the SE squeeze (mean), 1 for the gpool-bias gradient
  float* pool;            // [B][3*Cout] or null: mean, max, population std of stored v
};

#ifdef __CUDACC__
// Per-thread epilogue state for one output channel over NB boards.
template <typename T, int NB>
struct ConvEpiThread {
  const ConvEpi& e;
  const int c, Cout, B;
  float sc, sh, ma, mb;
  float s[NB], ss[NB], mx[NB], dot[NB], pre[NB];
  float k0[NB], ds[NB], dss[NB];  // shifted-data accumulators for a cancellation-free board variance
  __device__ __forceinline__ ConvEpiThread(const ConvEpi& e_, int c_, int Cout_, int B_)
      : e(e_), c(c_), Cout(Cout_), B(B_) {
    sc = e.scale ? e.scale[c] : 1.f;
    sh = e.scale ? e.shift[c] : 0.f;
    ma = e.mask_src ? e.mask_a[c] : 0.f;
    mb = e.mask_src ? e.mask_b[c] : 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) { s[j] = 0.f; ss[j] = 0.f; mx[j] = -INFINITY; dot[j] = 0.f; pre[j] = 0.f; k0[j] = 0.f; ds[j] = 0.f; dss[j] = 0.f; }
  }
  // one accumulator value: local board j (compile-time after unrolling), global board b, pixel p
  __device__ __forceinline__ void value(int j, int b, int p, float acc, T* __restrict__ out) {
    float v = acc;
    if (e.scale) v = fmaf(v, sc, sh);
    if (e.relu) v = fmaxf(v, 0.f);
    if (e.gbias) v += e.gbias[(size_t)b * Cout + c];
    const size_t idx = ((size_t)b * 81 + p) * Cout + c;
    float msrc = 0.f;
    if (e.mask_src) {
      pre[j] += v;
      msrc = kb_to_float<T>(((const T*)e.mask_src)[idx]);
      if (!(fmaf(msrc, ma, mb) > 0.f)) v = 0.f;
    }
    const T stored = kb_from_float<T>(v);
    out[idx] = stored;
    const float r = kb_to_float<T>(stored);
    s[j] += r;
    ss[j] = fmaf(r, r, ss[j]);
    mx[j] = fmaxf(mx[j], r);
    dot[j] = fmaf(r, msrc, dot[j]);
    if (e.pool) {
      if (p == 0) k0[j] = r;
      const float d = r - k0[j];
      ds[j] += d;
      dss[j] = fmaf(d, d, dss[j]);
    }
  }
  // after all pixels of local board j (global board b) have been fed
  __device__ __forceinline__ void board_done(int j, int b) {
    if (e.board_sum) e.board_sum[(size_t)b * Cout + c] = e.board_scale * (e.mask_src ? pre[j] : s[j]);
    if (e.pool) {
      const float mean = s[j] * (1.f / 81.f);
      const float dm = ds[j] * (1.f / 81.f);
      const float var = fmaxf(dss[j] * (1.f / 81.f) - dm * dm, 0.f);
      float* pr = e.pool + (size_t)b * 3 * Cout;
      pr[c] = mean; pr[Cout + c] = mx[j]; pr[2 * Cout + c] = sqrtf(var);
    }
  }
  // once per tile; nb_valid = number of boards actually fed
  __device__ __forceinline__ void finish(int nb_valid) {
    float a = 0.f, q = 0.f, d = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) if (j < nb_valid) { a += s[j]; q += ss[j]; d += dot[j]; }
    if (e.ch_sum) atomicAdd(&e.ch_sum[c], (double)a);
    if (e.ch_sumsq) atomicAdd(&e.ch_sumsq[c], (double)q);
    if (e.ch_dot) atomicAdd(&e.ch_dot[c], (double)d);
  }
};
#endif
