// optim.cu — the optimiser tail of a KataGo-PPO step as two launches over the flat gradient buffer
// (reference keisei/training/katago_ppo.py:926-933: GradScaler.unscale_ -> clip_grad_norm_(max_norm) -> Adam.step ->
// GradScaler.update). The stock sequence is a foreach inf-check/unscale, a foreach norm, a foreach
// multiply and a multi-tensor Adam: 40-60 launches and four passes over 53.4 M gradients.
//
//   kb_flat_grad_stats   one pass over the flat (still loss-scaled) gradient: sum of squares (double) and a non-finite flag
//   kb_adam_step_flat    one pass over (p, g, m, v): unscale, clip to the global norm, Adam (torch.optim.Adam semantics:
//                        no weight decay, no amsgrad), skipped entirely when the gradient was non-finite — parameters and
//                        both moments live in PyTorch's own tensors (pointer tables), so `optimizer.state_dict()` is
//                        unchanged and checkpoints interchange (checkpoint.py:123)
#include "kb_common.cuh"
#include "../../include/keisei_b200.h"

namespace {

__global__ void __launch_bounds__(256) flat_grad_stats_kernel(const float* __restrict__ g, long long n, double* __restrict__ out2) {
  __shared__ double scratch[32];
  double s = 0.0;
  int bad = 0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    const float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;   // non-finite inputs make q non-finite
    bad |= !(fabsf(q) <= 3.0e38f);
    s += (double)q;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    bad |= !(fabsf(v) <= 3.0e38f);
    s += (double)v * (double)v;
  }
  const double total = kb_block_sum_d(s, scratch);
  const int any_bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    if (isfinite(total)) atomicAdd(&out2[0], total);
    if (any_bad || !isfinite(total)) out2[1] = 1.0;
  }
}

struct AdamArgs {
  const float* flat_grad;
  float* const* p; float* const* m; float* const* v;   // device arrays [n_tensors]
  float* steps;                                        // device [n_tensors]: torch's per-parameter `step` scalars
  const long long* g_off; const long long* sizes;      // device [n_tensors]
  const int* chunk_tensor; const long long* chunk_start;  // device [n_chunks]
  const double* stats;                                 // [sum of squares of the scaled gradient, non-finite flag]
  const float* inv_scale;                              // device scalar (1 / loss scale) or null
  float grad_div, max_norm, lr, beta1, beta2, eps;
  float* grad_norm_out; float* found_inf_out;
  int chunk;
};

__global__ void __launch_bounds__(256) adam_step_flat_kernel(AdamArgs a) {
  const int t = a.chunk_tensor[blockIdx.x];
  const long long start = a.chunk_start[blockIdx.x];
  const long long size = a.sizes[t];
  const float inv = (a.inv_scale ? *a.inv_scale : 1.f) / a.grad_div;
  const bool found_inf = a.stats[1] != 0.0;
  const float norm = (float)sqrt(a.stats[0]) * inv;          // 2-norm of the unscaled, rank-averaged gradient
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.grad_norm_out = norm; *a.found_inf_out = found_inf ? 1.f : 0.f; }
  if (found_inf) return;                                      // GradScaler semantics: the whole step is skipped
  const float clip = fminf(a.max_norm / (norm + 1e-6f), 1.f); // clip_grad_norm_: coefficient clamped to 1
  const float step = a.steps[t] + 1.f;
  const float bc1 = 1.f - powf(a.beta1, step), bc2 = 1.f - powf(a.beta2, step);
  const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float gs = inv * clip;
  float* __restrict__ p = a.p[t] + start;
  float* __restrict__ m = a.m[t] + start;
  float* __restrict__ v = a.v[t] + start;
  const float* __restrict__ g = a.flat_grad + a.g_off[t] + start;
  const long long n = min((long long)a.chunk, size - start);
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float gi = g[i] * gs;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + a.eps);
  }
}

// the step counters are bumped by a trailing tiny kernel so that no chunk can observe the incremented value
__global__ void bump_steps_kernel(float* steps, int n, const double* stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && stats[1] == 0.0) steps[i] += 1.f;
}

}  // namespace

extern "C" int kb_flat_grad_stats(const float* flat_grad, long long n, double* out2, int num_sms, cudaStream_t stream) {
  KB_CHECK_ARG(flat_grad && out2 && n >= 0, "kb_flat_grad_stats: bad arguments");
  KB_CHECK_ARG((reinterpret_cast<uintptr_t>(flat_grad) & 15) == 0, "kb_flat_grad_stats: the flat gradient must be 16-byte aligned");
  KB_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), stream));
  if (n == 0) return KB_OK;
  if (num_sms <= 0) num_sms = 148;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 8LL * num_sms) blocks = 8LL * num_sms;
  if (blocks < 1) blocks = 1;
  flat_grad_stats_kernel<<<(unsigned)blocks, 256, 0, stream>>>(flat_grad, n, out2);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_adam_step_flat(const float* flat_grad, void* const* p_ptrs, void* const* m_ptrs, void* const* v_ptrs,
                                 float* steps, const long long* g_off, const long long* sizes, const int* chunk_tensor,
                                 const long long* chunk_start, int n_tensors, int n_chunks, int chunk, const double* stats,
                                 const float* inv_scale, float grad_div, float max_norm, float lr, float beta1, float beta2,
                                 float eps, float* grad_norm_out, float* found_inf_out, cudaStream_t stream) {
  KB_CHECK_ARG(flat_grad && p_ptrs && m_ptrs && v_ptrs && steps && g_off && sizes && chunk_tensor && chunk_start && stats &&
               grad_norm_out && found_inf_out, "kb_adam_step_flat: null pointer");
  KB_CHECK_ARG(n_tensors > 0 && n_chunks > 0 && chunk > 0 && grad_div > 0.f && max_norm > 0.f && lr > 0.f, "kb_adam_step_flat: bad arguments");
  AdamArgs a;
  a.flat_grad = flat_grad; a.p = (float* const*)p_ptrs; a.m = (float* const*)m_ptrs; a.v = (float* const*)v_ptrs; a.steps = steps;
  a.g_off = g_off; a.sizes = sizes; a.chunk_tensor = chunk_tensor; a.chunk_start = chunk_start; a.stats = stats;
  a.inv_scale = inv_scale; a.grad_div = grad_div; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.grad_norm_out = grad_norm_out; a.found_inf_out = found_inf_out; a.chunk = chunk;
  adam_step_flat_kernel<<<(unsigned)n_chunks, 256, 0, stream>>>(a);
  KB_CUDA_LAUNCH_CHECK();
  bump_steps_kernel<<<(unsigned)((n_tensors + 255) / 256), 256, 0, stream>>>(steps, n_tensors, stats);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
