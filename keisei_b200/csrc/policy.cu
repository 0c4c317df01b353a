// policy.cu — masked-policy kernels over the 11,259-action space (HBM-bound, one CTA per row).
//
//   kb_policy_sample     rollout:  mask -> softmax -> sample -> log-prob, + scalar value (K11, K12)
//                        reference: keisei/training/katago_ppo.py:589-613, value_adapter.py:79-96
//   kb_ppo_policy_fwd    update:   masked log-softmax, gather, entropy, clipped surrogate (K13, K14)
//                        reference: keisei/training/katago_ppo.py:33-43, :858-888
//   kb_ppo_policy_bwd    update:   d(loss)/d(logits), written once
//   kb_value_losses_*    update:   W/D/L cross-entropy (ignore_index=-1) + score MSE (K15)
//                        reference: keisei/training/katago_ppo.py:46-57, :910-912; value_adapter.py:98-126
//
// Each row is staged once into shared memory as fp32 masked logits (illegal -> -inf); every
// reduction afterwards runs out of shared memory, so HBM sees one read of the logits and one of
// the mask per pass. Logit rows may be padded (row_stride >= A); 16-byte vector loads are used
// when the row base is 16-byte aligned, scalar loads otherwise.
#include "kb_common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T> struct Vec4;  // 4 consecutive logits
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static constexpr int kAlign = 16;
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&o)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static constexpr int kAlign = 8;
};

// Legal-mask row in one of three storage kinds (include/keisei_b200.h: KB_MASK_*):
//   0 bytes  — the reference's (B, A) bool tensor, one byte per action (katago_ppo.py:160)
//   1 bits   — bit-packed, `pitch` 32-bit words per row, action i = bit (i & 31) of word (i >> 5): 1,408 B instead of
//              11,259 B per row (device-resident buffer)
//   2 none   — every action legal (supervised-learning policy cross-entropy, sl/trainer.py:147-149)
template <int MK>
struct MaskRow {
  const uint8_t* bytes;
  const uint32_t* words;
  __device__ __forceinline__ MaskRow(const void* mask, long long pitch, int row) {
    bytes = MK == 0 ? (const uint8_t*)mask + (size_t)row * (size_t)pitch : nullptr;
    words = MK == 1 ? (const uint32_t*)mask + (size_t)row * (size_t)pitch : nullptr;
  }
  __device__ __forceinline__ bool get(int i) const {
    if (MK == 2) return true;
    if (MK == 0) return bytes[i] != 0;
    return ((__ldg(words + (i >> 5)) >> (i & 31)) & 1u) != 0;
  }
  // 4 consecutive actions starting at i (i % 4 == 0) as a 4-bit set
  __device__ __forceinline__ uint32_t get4(int i) const {
    if (MK == 2) return 0xFu;
    if (MK == 1) return (__ldg(words + (i >> 5)) >> (i & 31)) & 0xFu;
    return (bytes[i] != 0 ? 1u : 0u) | (bytes[i + 1] != 0 ? 2u : 0u) | (bytes[i + 2] != 0 ? 4u : 0u) | (bytes[i + 3] != 0 ? 8u : 0u);
  }
};

// Stage one row: s_row[i] = legal ? logit : -inf. Returns per-thread (legal count, any NaN in raw logits).
// Four 4-logit groups per thread are loaded before any is consumed (memory-level parallelism: the kernel is HBM-bound).
template <typename T, int MK>
__device__ __forceinline__ void stage_row(const T* __restrict__ lrow, const MaskRow<MK>& mrow,
                                          int A, float* s_row, int& legal, int& has_nan) {
  legal = 0; has_nan = 0;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(lrow) % Vec4<T>::kAlign) == 0);
  const int A4 = vec_ok ? (A & ~3) : 0;
  constexpr int kUn = 4;
  for (int base = threadIdx.x * 4; base < A4; base += kThreads * 4 * kUn) {
    float v[kUn][4];
    uint32_t mk[kUn];
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      const int i = base + u * kThreads * 4;
      if (i < A4) { Vec4<T>::load(lrow + i, v[u]); mk[u] = mrow.get4(i); }
    }
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      const int i = base + u * kThreads * 4;
      if (i < A4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = (mk[u] >> j) & 1u;
          has_nan |= (v[u][j] != v[u][j]);
          legal += ok;
          s_row[i + j] = ok ? v[u][j] : -INFINITY;
        }
      }
    }
  }
  for (int i = A4 + threadIdx.x; i < A; i += kThreads) {
    const float v = kb_to_float<T>(lrow[i]);
    const bool ok = mrow.get(i);
    has_nan |= (v != v);
    legal += ok;
    s_row[i] = ok ? v : -INFINITY;
  }
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---------------------------------------------------------------------------------------------
// Rollout: sample + log-prob + scalar value
// ---------------------------------------------------------------------------------------------
template <typename T, int MK>
__global__ void __launch_bounds__(kThreads) policy_sample_kernel(
    const T* __restrict__ logits, long long row_stride, const void* __restrict__ mask, long long mask_pitch,
    const float* __restrict__ value_logits, const float* __restrict__ score_lead, float alpha,
    int A, unsigned long long seed, unsigned long long offset, int logprob_mode,
    const long long* __restrict__ forced_actions,
    long long* __restrict__ actions, float* __restrict__ logp_out, float* __restrict__ value_out,
    int* __restrict__ legal_count, int* __restrict__ flags) {
  extern __shared__ float s_row[];
  __shared__ float scratch[32];
  __shared__ float s_best[kThreads / 32];
  __shared__ int s_besti[kThreads / 32];
  const int row = blockIdx.x;
  const T* lrow = logits + (size_t)row * row_stride;
  const MaskRow<MK> mrow(mask, mask_pitch, row);
  int legal, has_nan;
  stage_row<T, MK>(lrow, mrow, A, s_row, legal, has_nan);
  __syncthreads();
  const int n_legal = (int)(kb_block_sum((float)legal, scratch) + 0.5f);
  if (threadIdx.x == 0) {
    legal_count[row] = n_legal;
    if (n_legal == 0) atomicAdd(&flags[0], 1);
  }
  // scalar value (K12): P(W) - P(L), optional score blend
  if (threadIdx.x == 32 && value_out != nullptr) {
    const float a = value_logits[row * 3 + 0], b = value_logits[row * 3 + 1], c = value_logits[row * 3 + 2];
    const float m = fmaxf(a, fmaxf(b, c));
    const float ea = expf(a - m), eb = expf(b - m), ec = expf(c - m);
    const float inv = 1.f / (ea + eb + ec);
    float v = ea * inv - ec * inv;
    if (alpha != 0.f && score_lead != nullptr) {
      const float s = fminf(fmaxf(score_lead[row], -1.f), 1.f);
      v = (1.f - alpha) * v + alpha * s;
    }
    value_out[row] = v;
  }
  if (n_legal == 0) {  // reference raises; keep outputs defined
    if (threadIdx.x == 0) { actions[row] = 0; logp_out[row] = __int_as_float(0x7fc00000); }
    return;
  }
  // pass 1: max, and Gumbel-max argmax among legal entries
  float m = -INFINITY, best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < A; i += kThreads) {
    const float l = s_row[i];
    if (l == -INFINITY) continue;  // illegal (or a legal -inf logit: probability exactly 0)
    m = fmaxf(m, l);
    const kb_philox4 r = kb_philox4x32_10(seed, (unsigned long long)row,
                                          (offset << 32) | (unsigned long long)(unsigned)i);
    const float u = kb_u32_to_unit(r.x);
    const float g = -__logf(-__logf(u));
    const float sc = l + g;
    if (sc > best || (sc == best && i < besti)) { best = sc; besti = i; }
  }
  // NaN logits on a legal entry: propagate through max like torch (m becomes NaN-free here, the
  // reference guard for NaN lives in update(); select_actions has none), keep going.
  m = kb_block_max(m, scratch);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_besti[threadIdx.x >> 5] = besti; }
  // pass 2: normaliser
  float s = 0.f;
  for (int i = threadIdx.x; i < A; i += kThreads) s += __expf(s_row[i] - m);  // exp(-inf)=0
  const float S = kb_block_sum(s, scratch);  // includes the __syncthreads that publishes s_best
  if (threadIdx.x < 32) {
    float b = threadIdx.x < kThreads / 32 ? s_best[threadIdx.x] : -INFINITY;
    int bi = threadIdx.x < kThreads / 32 ? s_besti[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, b, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
    }
    if (threadIdx.x == 0) { s_besti[0] = bi; }
  }
  __syncthreads();
  int a = s_besti[0];
  if (forced_actions != nullptr) {  // evaluate the log-prob of a caller-chosen action instead
    const long long fa = forced_actions[row];
    a = (fa >= 0 && fa < A) ? (int)fa : 0;
  }
  if (a == 0x7fffffff) {  // every legal logit was -inf/NaN: fall back to the first legal index
    int first = 0x7fffffff;
    for (int i = threadIdx.x; i < A; i += kThreads) if (mrow.get(i)) { first = min(first, i); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_besti[threadIdx.x >> 5] = first;
    __syncthreads();
    a = s_besti[0];
    for (int w = 1; w < kThreads / 32; ++w) a = min(a, s_besti[w]);
  }
  const float la = s_row[a];
  float lp;
  if (logprob_mode == 0) {
    // fp32 reference: Categorical(probs).log_prob = log(clamp(p, eps, 1-eps)), eps = finfo(float32).eps
    const float eps = 1.1920928955078125e-07f;
    const float p = expf(la - m) / S;
    lp = logf(fminf(fmaxf(p, eps), 1.f - eps));
  } else {
    // bf16 reference (autocast logits -> bf16 probs): probs rounded to bf16, renormalised in bf16,
    // clamped with eps = finfo(bfloat16).eps = 2^-7, log rounded to bf16.
    float q = 0.f;
    const float invS = 1.f / S;
    for (int i = threadIdx.x; i < A; i += kThreads) q += bf16_round(__expf(s_row[i] - m) * invS);
    const float Q = bf16_round(kb_block_sum(q, scratch));
    const float eps = 0.0078125f;
    const float p = bf16_round(bf16_round(expf(la - m) * invS) / Q);
    lp = bf16_round(logf(fminf(fmaxf(p, eps), 1.f - eps)));
  }
  if (threadIdx.x == 0) { actions[row] = (long long)a; logp_out[row] = lp; }
}

// ---------------------------------------------------------------------------------------------
// Rollout sampling over BIT-PACKED masks: one WARP per row, only the legal logits are touched.
// A shogi position has ~30-120 legal moves out of 11,259 actions: instead of staging the whole row (22.5 KB of bf16 logits
// + the mask) through shared memory and reducing over all of it, a warp reads the row's 352 mask words, walks their set
// bits and gathers just those logits (three short walks: max + Gumbel arg-max, normaliser, bf16-mode renormaliser; the
// re-reads hit L1). Same semantics and the same Philox keys as policy_sample_kernel: the drawn action is identical, the
// log-prob differs by summation order only. No shared memory, no block barriers, 8 rows per 256-thread CTA.
// ---------------------------------------------------------------------------------------------
// Fallback walk (rows with more legal entries than the shared-memory list holds, or an action space wider than the
// register-resident mask words): three bit-walks straight over the mask words, gathering from global memory each time.
template <typename T>
__device__ __noinline__ void sample_row_walk(
    const T* __restrict__ logits, long long row_stride, const uint32_t* __restrict__ mask, long long pitch,
    const float* __restrict__ value_logits, const float* __restrict__ score_lead, float alpha, int row,
    int A, unsigned long long seed, unsigned long long offset, int logprob_mode,
    const long long* __restrict__ forced_actions,
    long long* __restrict__ actions, float* __restrict__ logp_out, float* __restrict__ value_out,
    int* __restrict__ legal_count, int* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const T* lrow = logits + (size_t)row * row_stride;
  const uint32_t* words = mask + (size_t)row * pitch;
  const int nwords = (A + 31) >> 5;
  auto word_at = [&](int w) -> uint32_t {
    uint32_t bits = __ldg(words + w);
    if (w == nwords - 1 && (A & 31) != 0) bits &= (1u << (A & 31)) - 1u;   // padding bits are never actions
    return bits;
  };
  // walk 1: legal count, first legal index, max and Gumbel arg-max over the legal logits
  int legal = 0, first = 0x7fffffff, besti = 0x7fffffff;
  float m = -INFINITY, best = -INFINITY;
  for (int w = lane; w < nwords; w += 32) {
    uint32_t bits = word_at(w);
    legal += __popc(bits);
    if (bits != 0) first = min(first, w * 32 + __ffs(bits) - 1);
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      const float l = kb_to_float<T>(lrow[i]);
      if (l == -INFINITY) continue;  // a legal -inf logit: probability exactly 0
      m = fmaxf(m, l);
      const kb_philox4 r = kb_philox4x32_10(seed, (unsigned long long)row, (offset << 32) | (unsigned long long)(unsigned)i);
      const float sc = l - __logf(-__logf(kb_u32_to_unit(r.x)));
      if (sc > best || (sc == best && i < besti)) { best = sc; besti = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    legal += __shfl_xor_sync(0xffffffffu, legal, o);
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (lane == 0) {
    legal_count[row] = legal;
    if (legal == 0) atomicAdd(&flags[0], 1);
    if (value_out != nullptr) {   // scalar value (K12): P(W) - P(L), optional score blend
      const float a = value_logits[row * 3 + 0], b = value_logits[row * 3 + 1], c = value_logits[row * 3 + 2];
      const float mm = fmaxf(a, fmaxf(b, c));
      const float ea = expf(a - mm), eb = expf(b - mm), ec = expf(c - mm);
      const float inv = 1.f / (ea + eb + ec);
      float v = ea * inv - ec * inv;
      if (alpha != 0.f && score_lead != nullptr) v = (1.f - alpha) * v + alpha * fminf(fmaxf(score_lead[row], -1.f), 1.f);
      value_out[row] = v;
    }
  }
  if (legal == 0) {  // reference raises; keep outputs defined
    if (lane == 0) { actions[row] = 0; logp_out[row] = __int_as_float(0x7fc00000); }
    return;
  }
  int a = besti;
  if (forced_actions != nullptr) {
    const long long fa = forced_actions[row];
    a = (fa >= 0 && fa < A) ? (int)fa : 0;
  }
  if (a == 0x7fffffff) a = first;  // every legal logit was -inf / NaN: the first legal index
  // walk 2: normaliser
  float s = 0.f;
  for (int w = lane; w < nwords; w += 32) {
    uint32_t bits = word_at(w);
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      s += __expf(kb_to_float<T>(lrow[i]) - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float S = s;
  const bool a_legal = ((__ldg(words + (a >> 5)) >> (a & 31)) & 1u) != 0;
  const float la = a_legal ? kb_to_float<T>(lrow[a]) : -INFINITY;
  float lp;
  if (logprob_mode == 0) {
    const float eps = 1.1920928955078125e-07f;
    const float p = expf(la - m) / S;
    lp = logf(fminf(fmaxf(p, eps), 1.f - eps));
  } else {
    // bf16 reference semantics (see policy_sample_kernel): probabilities rounded to bf16 and renormalised in bf16
    float q = 0.f;
    const float invS = 1.f / S;
    for (int w = lane; w < nwords; w += 32) {
      uint32_t bits = word_at(w);
      while (bits) {
        const int i = w * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
        q += bf16_round(__expf(kb_to_float<T>(lrow[i]) - m) * invS);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float Q = bf16_round(q);
    const float eps = 0.0078125f;
    const float p = bf16_round(bf16_round(expf(la - m) * invS) / Q);
    lp = bf16_round(logf(fminf(fmaxf(p, eps), 1.f - eps)));
  }
  if (lane == 0) { actions[row] = (long long)a; logp_out[row] = lp; }
}

__device__ __forceinline__ float scalar_value_of(const float* __restrict__ value_logits, const float* __restrict__ score_lead,
                                                 float alpha, int row) {
  const float a = value_logits[row * 3 + 0], b = value_logits[row * 3 + 1], c = value_logits[row * 3 + 2];
  const float mm = fmaxf(a, fmaxf(b, c));
  const float ea = expf(a - mm), eb = expf(b - mm), ec = expf(c - mm);
  const float inv = 1.f / (ea + eb + ec);
  float v = ea * inv - ec * inv;
  if (alpha != 0.f && score_lead != nullptr) v = (1.f - alpha) * v + alpha * fminf(fmaxf(score_lead[row], -1.f), 1.f);
  return v;
}

// The kernel proper. Each lane loads its 11 mask words at once (one round trip), the warp compacts the legal action
// indices into a shared-memory list by a prefix sum over the popcounts, and the list is then processed 32 entries at a
// time with four independent gathers in flight per lane — two dependent memory round trips per row instead of one per
// legal action. The gathered logits stay in shared memory for the normaliser passes.
constexpr int kListCap = 640;        // a shogi position has at most 593 legal moves; longer rows take the walk
constexpr int kWordsPerLane = 11;    // 352 mask words (11,259 actions) held in registers

template <typename T>
__global__ void __launch_bounds__(kThreads) policy_sample_bits_kernel(
    const T* __restrict__ logits, long long row_stride, const uint32_t* __restrict__ mask, long long pitch,
    const float* __restrict__ value_logits, const float* __restrict__ score_lead, float alpha, int B,
    int A, unsigned long long seed, unsigned long long offset, int logprob_mode,
    const long long* __restrict__ forced_actions,
    long long* __restrict__ actions, float* __restrict__ logp_out, float* __restrict__ value_out,
    int* __restrict__ legal_count, int* __restrict__ flags) {
  __shared__ uint16_t s_idx[kThreads / 32][kListCap];
  __shared__ float s_val[kThreads / 32][kListCap];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * (kThreads / 32) + wid;
  if (row >= B) return;
  const int nwords = (A + 31) >> 5;
  const T* lrow = logits + (size_t)row * row_stride;
  const uint32_t* words = mask + (size_t)row * pitch;
  uint32_t wd[kWordsPerLane];
  int cnt = 0;
  const bool wide = nwords > 32 * kWordsPerLane || A > 65535;
  if (!wide) {
#pragma unroll
    for (int k = 0; k < kWordsPerLane; ++k) {
      const int w = lane + 32 * k;
      uint32_t bits = w < nwords ? __ldg(words + w) : 0u;
      if (w == nwords - 1 && (A & 31) != 0) bits &= (1u << (A & 31)) - 1u;   // padding bits are never actions
      wd[k] = bits;
      cnt += __popc(bits);
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int legal = __shfl_sync(0xffffffffu, incl, 31);
  if (wide || legal > kListCap) {   // warp-uniform
    sample_row_walk<T>(logits, row_stride, mask, pitch, value_logits, score_lead, alpha, row, A, seed, offset, logprob_mode,
                       forced_actions, actions, logp_out, value_out, legal_count, flags);
    return;
  }
  if (lane == 0) {
    legal_count[row] = legal;
    if (legal == 0) atomicAdd(&flags[0], 1);
    if (value_out != nullptr) value_out[row] = scalar_value_of(value_logits, score_lead, alpha, row);
  }
  if (legal == 0) {  // reference raises; keep outputs defined
    if (lane == 0) { actions[row] = 0; logp_out[row] = __int_as_float(0x7fc00000); }
    return;
  }
  // compaction: lane-major order (any fixed order gives the same arg-max; sums differ by rounding only)
  int pos = incl - cnt, first = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < kWordsPerLane; ++k) {
    uint32_t bits = wd[k];
    while (bits) {
      const int i = (lane + 32 * k) * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      first = min(first, i);
      s_idx[wid][pos++] = (uint16_t)i;
    }
  }
  __syncwarp();
  float m = -INFINITY, best = -INFINITY;
  int besti = 0x7fffffff;
  for (int j0 = 0; j0 < legal; j0 += 128) {
    int idx[4];
    float l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 32 + lane;
      idx[u] = j < legal ? (int)s_idx[wid][j] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) l[u] = idx[u] >= 0 ? kb_to_float<T>(lrow[idx[u]]) : -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (idx[u] < 0) continue;
      s_val[wid][j0 + u * 32 + lane] = l[u];
      if (l[u] == -INFINITY) continue;  // a legal -inf logit: probability exactly 0
      m = fmaxf(m, l[u]);
      const kb_philox4 r = kb_philox4x32_10(seed, (unsigned long long)row, (offset << 32) | (unsigned long long)(unsigned)idx[u]);
      const float sc = l[u] - __logf(-__logf(kb_u32_to_unit(r.x)));
      if (sc > best || (sc == best && idx[u] < besti)) { best = sc; besti = idx[u]; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  int a = besti;
  if (forced_actions != nullptr) {
    const long long fa = forced_actions[row];
    a = (fa >= 0 && fa < A) ? (int)fa : 0;
  }
  if (a == 0x7fffffff) a = first;  // every legal logit was -inf / NaN: the first legal index
  float s = 0.f;
  for (int j = lane; j < legal; j += 32) s += __expf(s_val[wid][j] - m);
  const float S = kb_warp_sum(s);
  const bool a_legal = ((__ldg(words + (a >> 5)) >> (a & 31)) & 1u) != 0;
  const float la = a_legal ? kb_to_float<T>(lrow[a]) : -INFINITY;
  float lp;
  if (logprob_mode == 0) {
    const float eps = 1.1920928955078125e-07f;
    const float p = expf(la - m) / S;
    lp = logf(fminf(fmaxf(p, eps), 1.f - eps));
  } else {
    // bf16 reference semantics (see policy_sample_kernel): probabilities rounded to bf16 and renormalised in bf16
    float q = 0.f;
    const float invS = 1.f / S;
    for (int j = lane; j < legal; j += 32) q += bf16_round(__expf(s_val[wid][j] - m) * invS);
    const float Q = bf16_round(kb_warp_sum(q));
    const float eps = 0.0078125f;
    const float p = bf16_round(bf16_round(expf(la - m) * invS) / Q);
    lp = bf16_round(logf(fminf(fmaxf(p, eps), 1.f - eps)));
  }
  if (lane == 0) { actions[row] = (long long)a; logp_out[row] = lp; }
}

// ---------------------------------------------------------------------------------------------
// Update: forward (per row) + reduction + backward
// ---------------------------------------------------------------------------------------------
template <typename T, int MK>
__global__ void __launch_bounds__(kThreads) ppo_policy_fwd_kernel(
    const T* __restrict__ logits, long long row_stride, const void* __restrict__ mask, long long mask_pitch,
    const long long* __restrict__ actions, int A, float* __restrict__ new_logp,
    float* __restrict__ row_entropy, float* __restrict__ row_lse, int* __restrict__ flags) {
  extern __shared__ float s_row[];
  __shared__ float scratch[32];
  const int row = blockIdx.x;
  const T* lrow = logits + (size_t)row * row_stride;
  const MaskRow<MK> mrow(mask, mask_pitch, row);
  int legal, has_nan;
  stage_row<T, MK>(lrow, mrow, A, s_row, legal, has_nan);
  __syncthreads();
  float m = -INFINITY;
  for (int i = threadIdx.x; i < A; i += kThreads) m = fmaxf(m, s_row[i]);
  m = kb_block_max(m, scratch);
  const int n_legal = (int)(kb_block_sum((float)legal, scratch) + 0.5f);
  const int any_nan = (int)(kb_block_sum((float)has_nan, scratch) + 0.5f);
  if (threadIdx.x == 0) {
    if (n_legal == 0) atomicAdd(&flags[0], 1);
    if (any_nan != 0) atomicAdd(&flags[1], 1);
  }
  if (n_legal == 0) {
    if (threadIdx.x == 0) { new_logp[row] = 0.f; row_entropy[row] = 0.f; row_lse[row] = 0.f; }
    return;
  }
  // S = sum exp(l-m);  U = sum exp(l-m)*(l-m)  (legal only; exp(-inf)=0 and the product is skipped)
  float s = 0.f, u = 0.f;
  for (int i = threadIdx.x; i < A; i += kThreads) {
    const float d = s_row[i] - m;
    if (d > -INFINITY) { const float e = __expf(d); s += e; u += e * d; }
  }
  const float S = kb_block_sum(s, scratch);
  const float U = kb_block_sum(u, scratch);
  if (threadIdx.x == 0) {
    const float logS = logf(S);
    const long long a = actions[row];
    const float la = (a >= 0 && a < A) ? s_row[a] : -INFINITY;
    new_logp[row] = (la - m) - logS;
    row_entropy[row] = logS - U / S;  // -sum p*logp over legal
    row_lse[row] = m + logS;
  }
}

// Single CTA: clipped-surrogate mean, entropy mean, and the per-row d(policy_loss)/d(new_logp).
__global__ void __launch_bounds__(1024) ppo_policy_reduce_kernel(
    const float* __restrict__ new_logp, const float* __restrict__ old_logp,
    const float* __restrict__ adv, const float* __restrict__ row_entropy, int B, float clip_eps,
    float* __restrict__ out2 /* [policy_loss, entropy] */, float* __restrict__ dlogp /* (B,) */) {
  __shared__ double scratch[32];
  double ls = 0.0, es = 0.0;
  const float lo = 1.f - clip_eps, hi = 1.f + clip_eps;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float ratio = expf(new_logp[i] - old_logp[i]);
    const float a = adv[i];
    const float s1 = ratio * a;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s2 = rc * a;
    ls += (double)fminf(s1, s2);
    es += (double)row_entropy[i];
    // torch.min backward: ties split evenly; clamp passes gradient when lo <= ratio <= hi
    const float inr = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    float g;
    if (s1 < s2) g = a; else if (s1 > s2) g = a * inr; else g = 0.5f * a + 0.5f * a * inr;
    if (!(s1 == s1) || !(s2 == s2)) g = __int_as_float(0x7fc00000);
    dlogp[i] = -(g * ratio) / (float)B;  // d(-mean(min))/d new_logp
  }
  const double L = kb_block_sum_d(ls, scratch);
  const double E = kb_block_sum_d(es, scratch);
  if (threadIdx.x == 0) { out2[0] = (float)(-L / (double)B); out2[1] = (float)(E / (double)B); }
}

// dlogits[i] = legal ? gP*dlogp[row]*(onehot - p_i) + gH/B * (-p_i*(logp_i + H_row)) : 0
template <typename T, int MK>
__global__ void __launch_bounds__(kThreads) ppo_policy_bwd_kernel(
    const T* __restrict__ logits, long long row_stride, const void* __restrict__ mask, long long mask_pitch,
    const long long* __restrict__ actions, int A, int B, const float* __restrict__ row_lse,
    const float* __restrict__ row_entropy, const float* __restrict__ dlogp,
    const float* __restrict__ g_policy, const float* __restrict__ g_entropy,
    T* __restrict__ dlogits, long long d_row_stride) {
  const int row = blockIdx.x;
  const T* lrow = logits + (size_t)row * row_stride;
  const MaskRow<MK> mrow(mask, mask_pitch, row);
  T* drow = dlogits + (size_t)row * d_row_stride;
  const float lse = row_lse[row], H = row_entropy[row];
  const float gl = g_policy[0] * dlogp[row];
  const float gh = g_entropy[0] / (float)B;
  const int a = (int)actions[row];
  for (int i = threadIdx.x; i < A; i += kThreads) {
    float d = 0.f;
    if (mrow.get(i)) {
      const float lp = kb_to_float<T>(lrow[i]) - lse;
      const float p = __expf(lp);
      d = gl * ((i == a ? 1.f : 0.f) - p);
      if (p > 0.f) d -= gh * p * (lp + H);
    }
    drow[i] = kb_from_float<T>(d);
  }
  for (int i = A + threadIdx.x; i < d_row_stride; i += kThreads) drow[i] = kb_from_float<T>(0.f);
}

// ---------------------------------------------------------------------------------------------
// Update kernels over BIT-PACKED masks (the product path: the rollout buffer stores packed masks and hands them over).
// Illegal actions contribute nothing to the softmax, so the forward is a pure streaming NaN scan of the raw logits (the
// reference's guard, katago_ppo.py:860-862, looks at every logit) plus a walk over the ~1 % legal entries, and the
// backward never reads an illegal logit: it writes zeros for them. No shared-memory row, 8 CTAs resident per SM.
// ---------------------------------------------------------------------------------------------
// NaN test on packed values without converting: |x| > inf  <=>  |x| + (all mantissa bits) carries into the sign position.
template <typename T> __device__ __forceinline__ uint32_t nan_bits(uint32_t x);
template <> __device__ __forceinline__ uint32_t nan_bits<bf16>(uint32_t x) { return ((x & 0x7fff7fffu) + 0x007f007fu) & 0x80008000u; }
template <> __device__ __forceinline__ uint32_t nan_bits<float>(uint32_t x) { return ((x & 0x7fffffffu) + 0x007fffffu) & 0x80000000u; }

template <typename T>
__global__ void __launch_bounds__(kThreads) ppo_policy_fwd_bits_kernel(
    const T* __restrict__ logits, long long row_stride, const uint32_t* __restrict__ mask, long long pitch,
    const long long* __restrict__ actions, int A, float* __restrict__ new_logp,
    float* __restrict__ row_entropy, float* __restrict__ row_lse, int* __restrict__ flags) {
  __shared__ float s_red[3][kThreads / 32];
  const int row = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const T* lrow = logits + (size_t)row * row_stride;
  const uint32_t* words = mask + (size_t)row * pitch;
  const int nwords = (A + 31) >> 5;
  auto word_at = [&](int w) -> uint32_t {
    uint32_t bits = __ldg(words + w);
    if (w == nwords - 1 && (A & 31) != 0) bits &= (1u << (A & 31)) - 1u;
    return bits;
  };
  // the first two mask words of this thread are requested before the stream starts (352 words: every word of a shogi row)
  const uint32_t w0 = threadIdx.x < nwords ? word_at(threadIdx.x) : 0u;
  const uint32_t w1 = threadIdx.x + kThreads < nwords ? word_at(threadIdx.x + kThreads) : 0u;
  // streaming NaN scan of the whole raw row
  uint32_t nan_acc = 0;
  constexpr int kPer = 16 / (int)sizeof(T);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(lrow) & 15) == 0;
  const int nvec = vec_ok ? A / kPer : 0;
  const uint4* v4 = reinterpret_cast<const uint4*>(lrow);
  constexpr int kUn = 6;
  for (int j = threadIdx.x; j < nvec; j += kThreads * kUn) {
    uint4 x[kUn];
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      const int jj = j + u * kThreads;
      x[u] = jj < nvec ? __ldg(v4 + jj) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kUn; ++u) nan_acc |= nan_bits<T>(x[u].x) | nan_bits<T>(x[u].y) | nan_bits<T>(x[u].z) | nan_bits<T>(x[u].w);
  }
  for (int i = nvec * kPer + threadIdx.x; i < A; i += kThreads) {
    const float v = kb_to_float<T>(lrow[i]);
    nan_acc |= (v != v) ? 1u : 0u;
  }
  // walk 1 over the legal entries: count and max
  float m = -INFINITY;
  int legal = 0;
  for (int w = threadIdx.x, k = 0; w < nwords; w += kThreads, ++k) {
    uint32_t bits = k == 0 ? w0 : (k == 1 ? w1 : word_at(w));
    legal += __popc(bits);
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      m = fmaxf(m, kb_to_float<T>(lrow[i]));
    }
  }
  m = kb_warp_max(m);
  const float lg = kb_warp_sum((float)legal);
  const uint32_t nn = __any_sync(0xffffffffu, nan_acc != 0) ? 1u : 0u;
  if (lane == 0) { s_red[0][wid] = m; s_red[1][wid] = lg; s_red[2][wid] = (float)nn; }
  __syncthreads();
  float lgs = 0.f, nns = 0.f;
  m = -INFINITY;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) { m = fmaxf(m, s_red[0][w]); lgs += s_red[1][w]; nns += s_red[2][w]; }
  const int n_legal = (int)(lgs + 0.5f);
  if (threadIdx.x == 0) {
    if (n_legal == 0) atomicAdd(&flags[0], 1);
    if (nns != 0.f) atomicAdd(&flags[1], 1);
  }
  if (n_legal == 0) {
    if (threadIdx.x == 0) { new_logp[row] = 0.f; row_entropy[row] = 0.f; row_lse[row] = 0.f; }
    return;
  }
  // walk 2: S = sum exp(l-m), U = sum exp(l-m)*(l-m) over the legal entries (the gathers hit L1 / L2: just streamed)
  float s = 0.f, u = 0.f;
  for (int w = threadIdx.x, k = 0; w < nwords; w += kThreads, ++k) {
    uint32_t bits = k == 0 ? w0 : (k == 1 ? w1 : word_at(w));
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      const float d = kb_to_float<T>(lrow[i]) - m;
      if (d > -INFINITY) { const float e = __expf(d); s += e; u += e * d; }
    }
  }
  s = kb_warp_sum(s);
  u = kb_warp_sum(u);
  __syncthreads();   // s_red reuse
  if (lane == 0) { s_red[0][wid] = s; s_red[1][wid] = u; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float S = 0.f, U = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { S += s_red[0][w]; U += s_red[1][w]; }
    const float logS = logf(S);
    const long long a = actions[row];
    float la = -INFINITY;
    if (a >= 0 && a < A && ((word_at((int)(a >> 5)) >> (a & 31)) & 1u)) la = kb_to_float<T>(lrow[a]);
    new_logp[row] = (la - m) - logS;
    row_entropy[row] = logS - U / S;  // -sum p*logp over legal
    row_lse[row] = m + logS;
  }
}

// Same formula as ppo_policy_bwd_kernel, organised around what is sparse: the gradient row is first written as 16-byte
// zero vectors that depend on no load (the write stream — all of the kernel's real traffic — starts at once), then, after
// one barrier, each thread walks the set bits of ITS mask words (352 words over 256 threads, ~80 set bits per row) and
// stores those few gradients over the zeros. Illegal logits are never read.
// Needs 16-byte aligned gradient rows (d_row_stride a multiple of 16 / sizeof(T)); the host falls back otherwise.
template <typename T>
__global__ void __launch_bounds__(kThreads) ppo_policy_bwd_bits_kernel(
    const T* __restrict__ logits, long long row_stride, const uint32_t* __restrict__ mask, long long pitch,
    const long long* __restrict__ actions, int A, int B, const float* __restrict__ row_lse,
    const float* __restrict__ row_entropy, const float* __restrict__ dlogp,
    const float* __restrict__ g_policy, const float* __restrict__ g_entropy,
    T* __restrict__ dlogits, long long d_row_stride) {
  constexpr int kPer = 16 / (int)sizeof(T);
  const int row = blockIdx.x;
  const T* lrow = logits + (size_t)row * row_stride;
  const uint32_t* words = mask + (size_t)row * pitch;
  T* dscalar = dlogits + (size_t)row * d_row_stride;
  uint4* drow = reinterpret_cast<uint4*>(dscalar);
  const int nvec = (int)(d_row_stride / kPer);
  const int nwords = (A + 31) >> 5;
  auto word_at = [&](int w) -> uint32_t {
    uint32_t bits = __ldg(words + w);
    if (w == nwords - 1 && (A & 31) != 0) bits &= (1u << (A & 31)) - 1u;
    return bits;
  };
  const uint32_t w0 = threadIdx.x < nwords ? word_at(threadIdx.x) : 0u;
  const uint32_t w1 = threadIdx.x + kThreads < nwords ? word_at(threadIdx.x + kThreads) : 0u;
  const float lse = row_lse[row], H = row_entropy[row];
  const float gl = g_policy[0] * dlogp[row];
  const float gh = g_entropy[0] / (float)B;
  const int a = (int)actions[row];
#pragma unroll 6
  for (int j = threadIdx.x; j < nvec; j += kThreads) drow[j] = make_uint4(0, 0, 0, 0);
  // the first legal logit of both words is requested before the barrier
  const float l0 = w0 != 0 ? kb_to_float<T>(lrow[threadIdx.x * 32 + __ffs(w0) - 1]) : 0.f;
  const float l1 = w1 != 0 ? kb_to_float<T>(lrow[(threadIdx.x + kThreads) * 32 + __ffs(w1) - 1]) : 0.f;
  __syncthreads();   // orders the zero vectors of the whole CTA before the scalar stores below (same addresses)
  for (int w = threadIdx.x, k = 0; w < nwords; w += kThreads, ++k) {
    uint32_t bits = k == 0 ? w0 : (k == 1 ? w1 : word_at(w));
    bool first = k < 2;
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      const float lp = (first ? (k == 0 ? l0 : l1) : kb_to_float<T>(lrow[i])) - lse;
      first = false;
      const float p = __expf(lp);
      float dd = gl * ((i == a ? 1.f : 0.f) - p);
      if (p > 0.f) dd -= gh * p * (lp + H);
      dscalar[i] = kb_from_float<T>(dd);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// W/D/L cross-entropy (ignore_index = -1, mean over valid rows) + score MSE. Single CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) value_losses_fwd_kernel(
    const float* __restrict__ value_logits, const long long* __restrict__ cats,
    const float* __restrict__ score_pred, const float* __restrict__ score_tgt, int B,
    float* __restrict__ out3 /* [value_loss, score_loss, n_valid] */) {
  __shared__ double scratch[32];
  double ce = 0.0, nv = 0.0, se = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const long long c = cats[i];
    if (c >= 0 && c < 3) {
      const float a = value_logits[i * 3], b = value_logits[i * 3 + 1], d = value_logits[i * 3 + 2];
      const float m = fmaxf(a, fmaxf(b, d));
      const float lse = m + logf(expf(a - m) + expf(b - m) + expf(d - m));
      ce += (double)(lse - value_logits[i * 3 + c]);
      nv += 1.0;
    }
    const float e = score_pred[i] - score_tgt[i];
    se += (double)(e * e);
  }
  const double CE = kb_block_sum_d(ce, scratch);
  const double NV = kb_block_sum_d(nv, scratch);
  const double SE = kb_block_sum_d(se, scratch);
  if (threadIdx.x == 0) {
    out3[0] = NV > 0.0 ? (float)(CE / NV) : 0.f;  // all-ignored -> graph-connected zero
    out3[1] = (float)(SE / (double)B);
    out3[2] = (float)NV;
  }
}

__global__ void __launch_bounds__(256) value_losses_bwd_kernel(
    const float* __restrict__ value_logits, const long long* __restrict__ cats,
    const float* __restrict__ score_pred, const float* __restrict__ score_tgt, int B,
    const float* __restrict__ out3, const float* __restrict__ g_value, const float* __restrict__ g_score,
    float* __restrict__ dvalue_logits, float* __restrict__ dscore) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float nvalid = out3[2];
  const long long c = cats[i];
  float d0 = 0.f, d1 = 0.f, d2 = 0.f;
  if (c >= 0 && c < 3 && nvalid > 0.f) {
    const float a = value_logits[i * 3], b = value_logits[i * 3 + 1], d = value_logits[i * 3 + 2];
    const float m = fmaxf(a, fmaxf(b, d));
    const float ea = expf(a - m), eb = expf(b - m), ed = expf(d - m);
    const float inv = 1.f / (ea + eb + ed);
    const float g = g_value[0] / nvalid;
    d0 = g * (ea * inv - (c == 0 ? 1.f : 0.f));
    d1 = g * (eb * inv - (c == 1 ? 1.f : 0.f));
    d2 = g * (ed * inv - (c == 2 ? 1.f : 0.f));
  }
  dvalue_logits[i * 3] = d0; dvalue_logits[i * 3 + 1] = d1; dvalue_logits[i * 3 + 2] = d2;
  dscore[i] = g_score[0] * 2.f * (score_pred[i] - score_tgt[i]) / (float)B;
}

// (rows, A) bool -> (rows, words) bit-packed. The mask is one flat byte array; a lane takes SIXTEEN consecutive bytes of a
// row: rows of 11,259 bytes start at every alignment, so it loads the two aligned 16-byte vectors the bytes straddle (the
// second one is the next lane's first: an L1 hit) and funnel-shifts them into place. Each 4-byte word becomes a nibble
// without branches (non-zero-byte detect + one multiply), a lane makes half an output word and a lane pair one word.
// No shared memory, no barriers, every load coalesced, eight 16-byte loads in flight per lane.
constexpr int kPackIters = 4;                          // a warp covers 32 * 4 sixteen-byte pieces = 64 output words
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t v) {   // bit k = (byte k of v != 0)
  const uint32_t y = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
  return (y * 0x00204081u) >> 28;                      // bits 7, 15, 23, 31 -> 28, 29, 30, 31
}
__global__ void __launch_bounds__(256) pack_mask_bits_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits,
                                                             long long rows, int A, int words, int chunks, int base_off) {
  // `mask` is the tensor's address rounded DOWN to 16 bytes and base_off the bytes in between (a row-sliced mask starts
  // anywhere); the bytes before / after the tensor inside its first / last vector are read but never used
  const int lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long row = gw / chunks;
  if (row >= rows) return;
  const int chunk = (int)(gw - row * chunks);
  const long long total_bytes = base_off + rows * (long long)A;
  const long long begin = base_off + row * (long long)A;   // flat byte offset of the row
  const int ws = (int)(begin & 15) >> 2, bs = (int)(begin & 3) * 8;
  const uint4* V = reinterpret_cast<const uint4*>(mask);
  const long long full_vecs = total_bytes >> 4;        // aligned vectors lying wholly inside the tensor
  auto load_vec = [&](long long g) -> uint4 {
    if (g < full_vecs) return __ldg(V + g);
    uint32_t w[4] = {0, 0, 0, 0};                      // the tensor's last partial vector: byte loads
    for (int b = 0; b < 16; ++b) { const long long o = 16 * g + b; if (o < total_bytes) w[b >> 2] |= (uint32_t)mask[o] << (8 * (b & 3)); }
    return make_uint4(w[0], w[1], w[2], w[3]);
  };
  const int npieces = words * 2;
  uint4 lo[kPackIters], hi[kPackIters];
#pragma unroll
  for (int it = 0; it < kPackIters; ++it) {
    const int c = (chunk * kPackIters + it) * 32 + lane;
    const bool live = c < npieces && 16 * c < A;
    const long long g = (begin + 16ll * c) >> 4;
    lo[it] = live ? load_vec(g) : make_uint4(0, 0, 0, 0);
    hi[it] = (live && (begin & 15) != 0) ? load_vec(g + 1) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int it = 0; it < kPackIters; ++it) {
    const int c = (chunk * kPackIters + it) * 32 + lane;
    const uint32_t w8[8] = {lo[it].x, lo[it].y, lo[it].z, lo[it].w, hi[it].x, hi[it].y, hi[it].z, hi[it].w};
    uint32_t o0, o1, o2, o3;
#define KB_PACK_CASE(WS)                                                                                   \
    case WS: o0 = __funnelshift_r(w8[WS], w8[WS + 1], bs); o1 = __funnelshift_r(w8[WS + 1], w8[WS + 2], bs);  \
             o2 = __funnelshift_r(w8[WS + 2], w8[WS + 3], bs); o3 = __funnelshift_r(w8[WS + 3], w8[WS + 4], bs); break;
    switch (ws) {                                      // warp-uniform: one row per warp
      KB_PACK_CASE(0) KB_PACK_CASE(1) KB_PACK_CASE(2)
      default: KB_PACK_CASE(3)
    }
#undef KB_PACK_CASE
    uint32_t half = nonzero_nibble(o0) | (nonzero_nibble(o1) << 4) | (nonzero_nibble(o2) << 8) | (nonzero_nibble(o3) << 12);
    const int left = A - 16 * c;                       // bytes past A belong to the next row
    if (left < 16) half &= left > 0 ? (1u << left) - 1u : 0u;
    const uint32_t other = __shfl_xor_sync(0xffffffffu, half, 1);
    if ((lane & 1) == 0 && c < npieces) bits[row * words + (c >> 1)] = half | (other << 16);
  }
}

// One launch gathers a shuffled minibatch out of the device-resident rollout storage (reference katago_ppo.py:829-841:
// eight index-gathers per minibatch): observations (row = obs_floats fp32), bit-packed masks and the per-sample scalars.
struct GatherArgs {
  const float* obs; const uint32_t* bits; const long long* actions; const float* old_lp; const float* adv;
  const long long* cats; const float* score; const float* returns; const long long* idx;
  float* o_obs; uint32_t* o_bits; long long* o_actions; float* o_old_lp; float* o_adv; long long* o_cats; float* o_score;
  float* o_returns;
  int obs_floats, words; long long n_src;
};
__global__ void __launch_bounds__(256) gather_minibatch_kernel(GatherArgs a) {
  const long long r = blockIdx.x;
  long long src = a.idx[r];
  if (src < 0 || src >= a.n_src) src = 0;   // validated on the host side of the op; never out of bounds here
  const float2* so = reinterpret_cast<const float2*>(a.obs + (size_t)src * a.obs_floats);   // rows are 8-byte aligned (even float count)
  float2* dob = reinterpret_cast<float2*>(a.o_obs + (size_t)r * a.obs_floats);
  const int n2 = a.obs_floats >> 1;
  for (int i = threadIdx.x; i < n2; i += 256) dob[i] = __ldg(so + i);
  if ((a.obs_floats & 1) && threadIdx.x == 0) a.o_obs[(size_t)r * a.obs_floats + a.obs_floats - 1] = a.obs[(size_t)src * a.obs_floats + a.obs_floats - 1];
  for (int i = threadIdx.x; i < a.words; i += 256) a.o_bits[(size_t)r * a.words + i] = __ldg(a.bits + (size_t)src * a.words + i);
  if (threadIdx.x == 0) {
    a.o_actions[r] = a.actions[src]; a.o_old_lp[r] = a.old_lp[src]; a.o_adv[r] = a.adv[src]; a.o_cats[r] = a.cats[src];
    a.o_score[r] = a.score[src]; a.o_returns[r] = a.returns[src];
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { kb_set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return KB_ERR_CUDA; }
  }
  return KB_OK;
}

}  // namespace

#define KB_TRY_RC(expr) do { int r__ = (expr); if (r__ != KB_OK) return r__; } while (0)
#define KB_MASK_DISPATCH(KERNEL, T, ...)                                                     \
  do {                                                                                       \
    if (mask_kind == 0) KERNEL<T, 0> __VA_ARGS__;                                            \
    else if (mask_kind == 1) KERNEL<T, 1> __VA_ARGS__;                                       \
    else KERNEL<T, 2> __VA_ARGS__;                                                           \
  } while (0)

static int check_mask(const void* mask, int mask_kind, long long mask_pitch, int A, const char* who) {
  KB_CHECK_ARG(mask_kind >= 0 && mask_kind <= 2, "%s: mask_kind must be 0 (bytes), 1 (bits) or 2 (none)", who);
  KB_CHECK_ARG(mask_kind == 2 || mask != nullptr, "%s: null mask", who);
  KB_CHECK_ARG(mask_kind != 0 || mask_pitch >= A, "%s: byte mask pitch %lld < %d", who, mask_pitch, A);
  KB_CHECK_ARG(mask_kind != 1 || mask_pitch * 32 >= A, "%s: bit mask pitch %lld words < %d bits", who, mask_pitch, A);
  return KB_OK;
}

extern "C" int kb_policy_sample(const void* logits, int logits_dtype, long long row_stride,
                                const void* mask, const float* value_logits, const float* score_lead,
                                float alpha, int B, int A, unsigned long long seed,
                                unsigned long long offset, int logprob_mode,
                                const long long* forced_actions, long long* actions,
                                float* logp, float* values, int* legal_count, int* flags,
                                int mask_kind, long long mask_pitch, cudaStream_t stream) {
  KB_CHECK_ARG(B >= 0 && A > 0 && row_stride >= A, "kb_policy_sample: bad shape B=%d A=%d stride=%lld", B, A, row_stride);
  KB_CHECK_ARG(logits_dtype == KB_F32 || logits_dtype == KB_BF16, "kb_policy_sample: bad dtype %d", logits_dtype);
  if (B == 0) return KB_OK;
  KB_TRY_RC(check_mask(mask, mask_kind, mask_pitch, A, "kb_policy_sample"));
  KB_CHECK_ARG(logits && actions && logp && legal_count && flags, "kb_policy_sample: null pointer");
  KB_CHECK_ARG(values == nullptr || value_logits != nullptr, "kb_policy_sample: values requested without value_logits");
  if (mask_kind == 1) {   // bit-packed masks: warp-per-row kernel over the legal entries only
    const int rows_per_cta = kThreads / 32;
    const unsigned grid = (unsigned)((B + rows_per_cta - 1) / rows_per_cta);
    if (logits_dtype == KB_F32)
      policy_sample_bits_kernel<float><<<grid, kThreads, 0, stream>>>((const float*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          value_logits, score_lead, alpha, B, A, seed, offset, logprob_mode, forced_actions, actions, logp, values, legal_count, flags);
    else
      policy_sample_bits_kernel<bf16><<<grid, kThreads, 0, stream>>>((const bf16*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          value_logits, score_lead, alpha, B, A, seed, offset, logprob_mode, forced_actions, actions, logp, values, legal_count, flags);
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  const size_t smem = (size_t)A * sizeof(float);
  KB_CHECK_ARG(smem <= 200 * 1024, "kb_policy_sample: action space %d too large for one CTA", A);
  if (logits_dtype == KB_F32) {
    if (int r = set_smem(policy_sample_kernel<float, 0>, smem)) return r;
    if (int r = set_smem(policy_sample_kernel<float, 1>, smem)) return r;
    if (int r = set_smem(policy_sample_kernel<float, 2>, smem)) return r;
    KB_MASK_DISPATCH(policy_sample_kernel, float, <<<B, kThreads, smem, stream>>>((const float*)logits, row_stride, mask, mask_pitch,
        value_logits, score_lead, alpha, A, seed, offset, logprob_mode, forced_actions, actions, logp, values, legal_count, flags));
  } else {
    if (int r = set_smem(policy_sample_kernel<bf16, 0>, smem)) return r;
    if (int r = set_smem(policy_sample_kernel<bf16, 1>, smem)) return r;
    if (int r = set_smem(policy_sample_kernel<bf16, 2>, smem)) return r;
    KB_MASK_DISPATCH(policy_sample_kernel, bf16, <<<B, kThreads, smem, stream>>>((const bf16*)logits, row_stride, mask, mask_pitch,
        value_logits, score_lead, alpha, A, seed, offset, logprob_mode, forced_actions, actions, logp, values, legal_count, flags));
  }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_ppo_policy_fwd(const void* logits, int logits_dtype, long long row_stride,
                                 const void* mask, const long long* actions,
                                 const float* old_logp, const float* adv, int B, int A, float clip_eps,
                                 float* new_logp, float* row_entropy, float* row_lse, float* dlogp,
                                 float* out2, int* flags, int mask_kind, long long mask_pitch, cudaStream_t stream) {
  KB_CHECK_ARG(B > 0 && A > 0 && row_stride >= A, "kb_ppo_policy_fwd: bad shape B=%d A=%d stride=%lld", B, A, row_stride);
  KB_CHECK_ARG(logits_dtype == KB_F32 || logits_dtype == KB_BF16, "kb_ppo_policy_fwd: bad dtype %d", logits_dtype);
  KB_TRY_RC(check_mask(mask, mask_kind, mask_pitch, A, "kb_ppo_policy_fwd"));
  KB_CHECK_ARG(logits && actions && old_logp && adv && new_logp && row_entropy && row_lse && dlogp && out2 && flags,
               "kb_ppo_policy_fwd: null pointer");
  const size_t smem = (size_t)A * sizeof(float);
  KB_CHECK_ARG(smem <= 200 * 1024, "kb_ppo_policy_fwd: action space %d too large for one CTA", A);
  if (mask_kind == 1) {   // bit-packed masks: streaming NaN scan + walk over the legal entries, no shared-memory row
    if (logits_dtype == KB_F32)
      ppo_policy_fwd_bits_kernel<float><<<B, kThreads, 0, stream>>>((const float*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          actions, A, new_logp, row_entropy, row_lse, flags);
    else
      ppo_policy_fwd_bits_kernel<bf16><<<B, kThreads, 0, stream>>>((const bf16*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          actions, A, new_logp, row_entropy, row_lse, flags);
  } else if (logits_dtype == KB_F32) {
    if (int r = set_smem(ppo_policy_fwd_kernel<float, 0>, smem)) return r;
    if (int r = set_smem(ppo_policy_fwd_kernel<float, 1>, smem)) return r;
    if (int r = set_smem(ppo_policy_fwd_kernel<float, 2>, smem)) return r;
    KB_MASK_DISPATCH(ppo_policy_fwd_kernel, float, <<<B, kThreads, smem, stream>>>((const float*)logits, row_stride, mask, mask_pitch,
        actions, A, new_logp, row_entropy, row_lse, flags));
  } else {
    if (int r = set_smem(ppo_policy_fwd_kernel<bf16, 0>, smem)) return r;
    if (int r = set_smem(ppo_policy_fwd_kernel<bf16, 1>, smem)) return r;
    if (int r = set_smem(ppo_policy_fwd_kernel<bf16, 2>, smem)) return r;
    KB_MASK_DISPATCH(ppo_policy_fwd_kernel, bf16, <<<B, kThreads, smem, stream>>>((const bf16*)logits, row_stride, mask, mask_pitch,
        actions, A, new_logp, row_entropy, row_lse, flags));
  }
  KB_CUDA_LAUNCH_CHECK();
  ppo_policy_reduce_kernel<<<1, 1024, 0, stream>>>(new_logp, old_logp, adv, row_entropy, B, clip_eps, out2, dlogp);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_ppo_policy_bwd(const void* logits, int logits_dtype, long long row_stride,
                                 const void* mask, const long long* actions, int B, int A,
                                 const float* row_lse, const float* row_entropy, const float* dlogp,
                                 const float* g_policy, const float* g_entropy, void* dlogits,
                                 long long d_row_stride, int mask_kind, long long mask_pitch, cudaStream_t stream) {
  KB_CHECK_ARG(B > 0 && A > 0 && row_stride >= A && d_row_stride >= A, "kb_ppo_policy_bwd: bad shape");
  KB_CHECK_ARG(logits_dtype == KB_F32 || logits_dtype == KB_BF16, "kb_ppo_policy_bwd: bad dtype %d", logits_dtype);
  KB_TRY_RC(check_mask(mask, mask_kind, mask_pitch, A, "kb_ppo_policy_bwd"));
  KB_CHECK_ARG(logits && actions && row_lse && row_entropy && dlogp && g_policy && g_entropy && dlogits,
               "kb_ppo_policy_bwd: null pointer");
  const long long per16 = logits_dtype == KB_F32 ? 4 : 8;
  if (mask_kind == 1 && d_row_stride % per16 == 0 && ((uintptr_t)dlogits & 15) == 0) {
    if (logits_dtype == KB_F32)
      ppo_policy_bwd_bits_kernel<float><<<B, kThreads, 0, stream>>>((const float*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          actions, A, B, row_lse, row_entropy, dlogp, g_policy, g_entropy, (float*)dlogits, d_row_stride);
    else
      ppo_policy_bwd_bits_kernel<bf16><<<B, kThreads, 0, stream>>>((const bf16*)logits, row_stride, (const uint32_t*)mask, mask_pitch,
          actions, A, B, row_lse, row_entropy, dlogp, g_policy, g_entropy, (bf16*)dlogits, d_row_stride);
  } else if (logits_dtype == KB_F32)
    KB_MASK_DISPATCH(ppo_policy_bwd_kernel, float, <<<B, kThreads, 0, stream>>>((const float*)logits, row_stride, mask, mask_pitch, actions,
        A, B, row_lse, row_entropy, dlogp, g_policy, g_entropy, (float*)dlogits, d_row_stride));
  else
    KB_MASK_DISPATCH(ppo_policy_bwd_kernel, bf16, <<<B, kThreads, 0, stream>>>((const bf16*)logits, row_stride, mask, mask_pitch, actions,
        A, B, row_lse, row_entropy, dlogp, g_policy, g_entropy, (bf16*)dlogits, d_row_stride));
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_pack_mask_bits(const void* mask_bytes, void* bits, long long rows, int A, int words, cudaStream_t stream) {
  KB_CHECK_ARG(rows >= 0 && A > 0 && words * 32 >= A, "kb_pack_mask_bits: bad shape rows=%lld A=%d words=%d", rows, A, words);
  if (rows == 0) return KB_OK;
  KB_CHECK_ARG(mask_bytes && bits, "kb_pack_mask_bits: null pointer");
  const int base_off = (int)((uintptr_t)mask_bytes & 15);
  const int chunks = (words * 2 + 32 * kPackIters - 1) / (32 * kPackIters);
  const long long warps = rows * chunks;
  pack_mask_bits_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>((const uint8_t*)mask_bytes - base_off, (uint32_t*)bits, rows, A, words,
                                                                        chunks, base_off);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_gather_minibatch(const float* obs, const void* mask_bits, const long long* actions, const float* old_lp,
                                   const float* adv, const long long* cats, const float* score, const float* returns,
                                   const long long* idx, long long n_src, int n_out, int obs_floats, int words, float* o_obs,
                                   void* o_bits, long long* o_actions, float* o_old_lp, float* o_adv, long long* o_cats,
                                   float* o_score, float* o_returns, cudaStream_t stream) {
  KB_CHECK_ARG(n_out >= 0 && n_src > 0 && obs_floats > 0 && obs_floats % 2 == 0 && words > 0, "kb_gather_minibatch: bad shape");
  if (n_out == 0) return KB_OK;
  KB_CHECK_ARG(obs && mask_bits && actions && old_lp && adv && cats && score && returns && idx && o_obs && o_bits && o_actions &&
               o_old_lp && o_adv && o_cats && o_score && o_returns, "kb_gather_minibatch: null pointer");
  GatherArgs a;
  a.obs = obs; a.bits = (const uint32_t*)mask_bits; a.actions = actions; a.old_lp = old_lp; a.adv = adv; a.cats = cats; a.score = score;
  a.returns = returns; a.idx = idx; a.o_obs = o_obs; a.o_bits = (uint32_t*)o_bits; a.o_actions = o_actions; a.o_old_lp = o_old_lp;
  a.o_adv = o_adv; a.o_cats = o_cats; a.o_score = o_score; a.o_returns = o_returns; a.obs_floats = obs_floats; a.words = words; a.n_src = n_src;
  gather_minibatch_kernel<<<(unsigned)n_out, 256, 0, stream>>>(a);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_value_losses_fwd(const float* value_logits, const long long* cats, const float* score_pred,
                                   const float* score_tgt, int B, float* out3, cudaStream_t stream) {
  KB_CHECK_ARG(B > 0, "kb_value_losses_fwd: B must be > 0");
  KB_CHECK_ARG(value_logits && cats && score_pred && score_tgt && out3, "kb_value_losses_fwd: null pointer");
  value_losses_fwd_kernel<<<1, 1024, 0, stream>>>(value_logits, cats, score_pred, score_tgt, B, out3);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_value_losses_bwd(const float* value_logits, const long long* cats, const float* score_pred,
                                   const float* score_tgt, int B, const float* out3, const float* g_value,
                                   const float* g_score, float* dvalue_logits, float* dscore,
                                   cudaStream_t stream) {
  KB_CHECK_ARG(B > 0, "kb_value_losses_bwd: B must be > 0");
  KB_CHECK_ARG(value_logits && cats && score_pred && score_tgt && out3 && g_value && g_score && dvalue_logits && dscore,
               "kb_value_losses_bwd: null pointer");
  value_losses_bwd_kernel<<<kb_ceil_div(B, 256), 256, 0, stream>>>(value_logits, cats, score_pred, score_tgt, B, out3,
                                                                   g_value, g_score, dvalue_logits, dscore);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
