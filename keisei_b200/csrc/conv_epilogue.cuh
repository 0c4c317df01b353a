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
  void* board_bf;         // [B][Cout] bf16 or null: copy of board_sum (operand of the tcgen05 SE layer)
  float* pool;            // [B][3*Cout] or null: mean, max, population std of stored v
  // ---- fused evaluation tail of a GlobalPoolBiasBlock (conv3x3_tc2_kernel<.., kSeTail = true> only; se_resnet.py:83-90):
  // out = relu((acc*scale + shift) * sigmoid(se_scale) + se_shift + res), (se_scale, se_shift) = W2 relu(W1 mean + b1) + b2
  // with mean = board mean of acc*scale + shift; pool / pool_bf receive the global-pool statistics of `out`
  const void* res;        // [B][81][Cout] block input (activation dtype) or null
  const float* se_w1;     // [S][Cout]
  const float* se_b1;     // [S]
  const float* se_w2;     // [2*Cout][S]
  const float* se_b2;     // [2*Cout]
  void* pool_bf;          // [B][3*Cout] bf16 copy of pool or null (operand of the next block's tcgen05 global_fc)
};

// Feature bits: a kernel may be instantiated for a fixed feature set F (branches resolved at compile
// time — the tcgen05 epilogue is fully unrolled over 243 columns) or with kEpiDynamic, where the
// null-ness of the ConvEpi pointers decides at run time (the SIMT kernels).
enum : int {
  kEpiAffine = 1, kEpiRelu = 2, kEpiGbias = 4, kEpiMask = 8, kEpiSum = 16, kEpiSumSq = 32, kEpiDot = 64,
  kEpiBoard = 128, kEpiPool = 256, kEpiDynamic = -1
};
inline int conv_epi_features(const ConvEpi& e) {
  return (e.scale ? kEpiAffine : 0) | (e.relu ? kEpiRelu : 0) | (e.gbias ? kEpiGbias : 0) | (e.mask_src ? kEpiMask : 0) |
         (e.ch_sum ? kEpiSum : 0) | (e.ch_sumsq ? kEpiSumSq : 0) | (e.ch_dot ? kEpiDot : 0) | (e.board_sum ? kEpiBoard : 0) |
         (e.pool ? kEpiPool : 0);
}

#ifdef __CUDACC__
// Per-thread epilogue state for one output channel over NB boards.
template <typename T, int NB, int F>
struct ConvEpiThread {
  const ConvEpi& e;
  const int c, Cout, B;
  float sc, sh, ma, mb;
  float s[NB], ss[NB], mx[NB], dot[NB], pre[NB];
  float k0[NB], ds[NB], dss[NB];  // shifted-data accumulators for a cancellation-free board variance
  static __device__ __forceinline__ bool on(int bit, bool runtime) { return F == kEpiDynamic ? runtime : (F & bit) != 0; }
  __device__ __forceinline__ ConvEpiThread(const ConvEpi& e_, int c_, int Cout_, int B_)
      : e(e_), c(c_), Cout(Cout_), B(B_) {
    const bool aff = on(kEpiAffine, e.scale != nullptr), msk = on(kEpiMask, e.mask_src != nullptr);
    sc = aff ? e.scale[c] : 1.f;
    sh = aff ? e.shift[c] : 0.f;
    ma = msk ? e.mask_a[c] : 0.f;
    mb = msk ? e.mask_b[c] : 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) { s[j] = 0.f; ss[j] = 0.f; mx[j] = -INFINITY; dot[j] = 0.f; pre[j] = 0.f; k0[j] = 0.f; ds[j] = 0.f; dss[j] = 0.f; }
  }
  // element index of (board b, pixel p) for this thread's channel
  __device__ __forceinline__ size_t index(int b, int p) const { return ((size_t)b * 81 + p) * Cout + c; }
  // Per-board cursor: all 64-bit address arithmetic and the per-(board, channel) bias load happen once per
  // board here, so the per-pixel code is a handful of FP ops and one store at a constant row offset.
  T* optr;
  float gb;
  __device__ __forceinline__ void begin_board(int b, T* __restrict__ out) {
    optr = out + index(b, 0);
    gb = on(kEpiGbias, e.gbias != nullptr) ? e.gbias[(size_t)b * Cout + c] : 0.f;
  }
  // one accumulator value of the current board: local board j and pixel p are compile-time after unrolling.
  // `msrc` = the mask-source element of this pixel, pre-loaded by the caller; ignored when the mask is off.
  __device__ __forceinline__ void value(int j, int p, float acc, float msrc = 0.f) {
    float v = acc;
    if (on(kEpiAffine, e.scale != nullptr)) v = fmaf(v, sc, sh);
    if (on(kEpiRelu, e.relu != 0)) v = fmaxf(v, 0.f);
    if (on(kEpiGbias, e.gbias != nullptr)) v += gb;
    if (on(kEpiMask, e.mask_src != nullptr)) {
      pre[j] += v;
      if (!(fmaf(msrc, ma, mb) > 0.f)) v = 0.f;
    } else {
      msrc = 0.f;
    }
    const T stored = kb_from_float<T>(v);
    optr[(unsigned)p * (unsigned)Cout] = stored;
    const float r = kb_to_float<T>(stored);
    s[j] += r;
    if (on(kEpiSumSq, e.ch_sumsq != nullptr)) ss[j] = fmaf(r, r, ss[j]);
    if (on(kEpiDot, e.ch_dot != nullptr)) dot[j] = fmaf(r, msrc, dot[j]);
    if (on(kEpiPool, e.pool != nullptr)) {
      mx[j] = fmaxf(mx[j], r);
      if (p == 0) k0[j] = r;
      const float d = r - k0[j];
      ds[j] += d;
      dss[j] = fmaf(d, d, dss[j]);
    }
  }
  // after all pixels of local board j (global board b) have been fed
  __device__ __forceinline__ void board_done(int j, int b) {
    if (on(kEpiBoard, e.board_sum != nullptr)) {
      const float bs = e.board_scale * (on(kEpiMask, e.mask_src != nullptr) ? pre[j] : s[j]);
      e.board_sum[(size_t)b * Cout + c] = bs;
      if (e.board_bf) ((bf16*)e.board_bf)[(size_t)b * Cout + c] = __float2bfloat16_rn(bs);
    }
    if (on(kEpiPool, e.pool != nullptr)) {
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
    if (on(kEpiSum, e.ch_sum != nullptr)) atomicAdd(&e.ch_sum[c], (double)a);
    if (on(kEpiSumSq, e.ch_sumsq != nullptr)) atomicAdd(&e.ch_sumsq[c], (double)q);
    if (on(kEpiDot, e.ch_dot != nullptr)) atomicAdd(&e.ch_dot[c], (double)d);
  }
};
#endif
