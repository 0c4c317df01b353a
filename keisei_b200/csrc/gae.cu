// gae.cu — reverse-time GAE scan (K17) and whole-buffer advantage normalisation (K18).
//
// Reference semantics: keisei/training/gae.py:8-73 (compute_gae), :76-148 (compute_gae_padded),
// :151-218 (compute_gae_gpu), :221-296 (compute_gae_padded_gpu); normalisation
// keisei/training/katago_ppo.py:797-798.
//
//   nv[t]    = override[t] if !isnan(override[t]) else (next_value if t is the column's last step
//              else values[t+1])
//   nd[t]    = 1 - terminated[t]
//   delta[t] = (rewards[t] + (gamma*nv[t])*nd[t]) - values[t]
//   A[t]     = delta[t] + ((gamma*lam)*nd[t]) * A[t+1]
//
// Every product/sum is rounded separately (no FMA contraction) so the result is bit-identical to
// the op-by-op PyTorch evaluation of the same expression.
//
// Layout: (T, N) row-major, N fastest. One CTA owns 32 adjacent env columns; the CTA streams the
// column strip in chunks of TC timesteps from the end: all 256 threads compute delta/decay with
// coalesced 128-byte row reads into shared memory, warp 0 runs the dependent scan out of shared
// memory (one lane per column, carry kept in a register across chunks), all threads write the
// advantages back coalesced.
#include "kb_common.cuh"

namespace {

template <typename F> struct GaeOps;
template <> struct GaeOps<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
};
template <> struct GaeOps<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};

template <typename F, int TC>
__global__ void __launch_bounds__(256) gae_scan_kernel(
    const F* __restrict__ rewards, const F* __restrict__ values, const void* __restrict__ terminated,
    int term_kind,  // 0: uint8/bool, 1: same float type as F
    const F* __restrict__ next_value, const F* __restrict__ override_nv,
    const int* __restrict__ lengths, F* __restrict__ adv, int T, int N, F gamma, F gamma_lam) {
  __shared__ F s_delta[TC][32];
  __shared__ F s_decay[TC][32];
  using O = GaeOps<F>;
  const int n0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const int c = tid & 31;
  const int n = n0 + c;
  const bool col_ok = n < N;
  int last_step = T - 1;
  F nvb = F(0);
  if (col_ok) {
    nvb = next_value[n];
    if (lengths != nullptr) {
      int l = lengths[n] - 1;
      last_step = l < 0 ? 0 : l;
    }
  }
  F carry = F(0);  // only meaningful in warp 0
  for (int t1 = T; t1 > 0; t1 -= TC) {
    const int t0 = t1 - TC > 0 ? t1 - TC : 0;
    const int rows = t1 - t0;
    // phase 1: elementwise delta / decay (each warp takes whole rows -> 128 B coalesced reads)
    for (int tl = tid >> 5; tl < rows; tl += 8) {
      const int t = t0 + tl;
      F d = F(0), k = F(0);
      if (col_ok) {
        const size_t i = (size_t)t * N + n;
        const F r = rewards[i];
        const F v = values[i];
        F nv;
        if (t == T - 1) nv = nvb; else nv = values[i + N];
        if (lengths != nullptr && t == last_step) nv = nvb;
        if (override_nv != nullptr) {
          const F o = override_nv[i];
          if (!(o != o)) nv = o;
        }
        F term;
        if (term_kind == 0) term = (F)(((const uint8_t*)terminated)[i] != 0 ? 1 : 0);
        else term = ((const F*)terminated)[i];
        const F nd = O::sub(F(1), term);
        d = O::sub(O::add(r, O::mul(O::mul(gamma, nv), nd)), v);
        k = O::mul(gamma_lam, nd);
      }
      s_delta[tl][c] = d;
      s_decay[tl][c] = k;
    }
    __syncthreads();
    // phase 2: the dependent scan, one lane per column
    if (tid < 32) {
      F last = carry;
#pragma unroll 8
      for (int tl = rows - 1; tl >= 0; --tl) {
        last = O::add(s_delta[tl][c], O::mul(s_decay[tl][c], last));
        s_delta[tl][c] = last;
      }
      carry = last;
    }
    __syncthreads();
    // phase 3: coalesced write-back
    if (col_ok) {
      for (int tl = tid >> 5; tl < rows; tl += 8) adv[(size_t)(t0 + tl) * N + n] = s_delta[tl][c];
    }
    __syncthreads();
  }
}

// (x - mean) / (std_unbiased + eps) over n elements, in place. Single CTA, double accumulators.
__global__ void __launch_bounds__(1024) adv_normalize_kernel(float* __restrict__ x, long long n, float eps) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  const double mean = kb_block_sum_d(s, scratch) / (double)n;
  double q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  const double var = kb_block_sum_d(q, scratch) / (double)(n - 1);
  const float meanf = (float)mean;
  const float denom = (float)sqrt(var) + eps;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) x[i] = (x[i] - meanf) / denom;
}

}  // namespace

extern "C" int kb_gae_scan(const void* rewards, const void* values, const void* terminated,
                           int term_kind, const void* next_value, const void* override_nv,
                           const int* lengths, void* adv, int T, int N, double gamma, double lam,
                           int dtype_is_f64, cudaStream_t stream) {
  KB_CHECK_ARG(T >= 0 && N >= 0, "kb_gae_scan: negative shape T=%d N=%d", T, N);
  KB_CHECK_ARG(term_kind == 0 || term_kind == 1, "kb_gae_scan: term_kind must be 0 or 1");
  if (T == 0 || N == 0) return KB_OK;
  KB_CHECK_ARG(rewards && values && terminated && next_value && adv, "kb_gae_scan: null pointer");
  const int grid = kb_ceil_div(N, 32);
  if (dtype_is_f64) {
    gae_scan_kernel<double, 64><<<grid, 256, 0, stream>>>(
        (const double*)rewards, (const double*)values, terminated, term_kind, (const double*)next_value,
        (const double*)override_nv, lengths, (double*)adv, T, N, gamma, gamma * lam);
  } else {
    // python float (double) scalars are rounded to the tensor dtype before the multiply
    gae_scan_kernel<float, 128><<<grid, 256, 0, stream>>>(
        (const float*)rewards, (const float*)values, terminated, term_kind, (const float*)next_value,
        (const float*)override_nv, lengths, (float*)adv, T, N, (float)gamma, (float)(gamma * lam));
  }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_advantage_normalize(float* adv, long long n, float eps, cudaStream_t stream) {
  KB_CHECK_ARG(n >= 0, "kb_advantage_normalize: negative n");
  if (n <= 1) return KB_OK;  // reference skips when numel <= 1 (katago_ppo.py:797)
  KB_CHECK_ARG(adv != nullptr, "kb_advantage_normalize: null pointer");
  adv_normalize_kernel<<<1, 1024, 0, stream>>>(adv, n, eps);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
