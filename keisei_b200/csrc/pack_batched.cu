// pack_batched.cu — every weight re-pack of a network in a handful of launches.
//
// After each optimiser step the fp32 PyTorch parameters are re-packed into the kernels' layouts (conv weights ->
// [Cout][9][Cin] forward and flipped [Cin][9][Cout] data-gradient packs, Linear weights -> zero-padded bf16, eval
// BatchNorm -> folded affine): 330 small launches for the 40-block model, ~1.7 ms of pure launch latency that does not
// shrink with the batch (4 % of a 1024-sample-per-GPU step). Here the jobs travel BY VALUE as one large kernel
// parameter (CUDA >= 12.1: up to 32,764 bytes), so no device-side table and no host->device copy is needed:
// blockIdx.y = job, blockIdx.x = 2048-element chunk of that job (blocks past a job's size exit immediately).
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

constexpr int kChunk = 2048;          // elements per block (256 threads x 8)
constexpr int kJobsPerLaunch = 256;   // 256 x 72 B = 18 KB of kernel parameters

struct PackBatch { PackJob jobs[kJobsPerLaunch]; };

template <typename T>
__device__ __forceinline__ void conv_elem(const PackJob& j, long long i) {
  const int Cout = j.n0, Cin = j.n1, Cinp = j.n2;
  const int ci = (int)(i % Cinp);
  const int tap = (int)((i / Cinp) % 9);
  const int co = (int)(i / ((long long)Cinp * 9));
  const float v = ci < Cin ? j.s0[((size_t)co * Cin + ci) * 9 + tap] : 0.f;
  ((T*)j.d0)[i] = kb_from_float<T>(v);
  if (j.d1 != nullptr) ((T*)j.d1)[((size_t)ci * 9 + (8 - tap)) * Cout + co] = kb_from_float<T>(v);
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const __grid_constant__ PackBatch batch) {
  const PackJob& j = batch.jobs[blockIdx.y];
  const long long n = j.count;
  const long long base = (long long)blockIdx.x * kChunk;
  if (base >= n) return;
  const long long end = base + kChunk < n ? base + kChunk : n;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    switch (j.kind) {
      case KB_PACK_CONV:
        if (j.dtype == KB_F32) conv_elem<float>(j, i); else conv_elem<bf16>(j, i);
        break;
      case KB_PACK_LINEAR: {   // w [N][K] fp32 -> out [Np][Kp] bf16, zero padded
        const int N = j.n0, K = j.n1, Kp = j.n3;
        const int k = (int)(i % Kp), r = (int)(i / Kp);
        ((bf16*)j.d0)[i] = __float2bfloat16_rn((r < N && k < K) ? j.s0[(size_t)r * K + k] : 0.f);
        break;
      }
      default: {               // KB_PACK_BN: a = w / sqrt(rv + eps), b = bias - rm * a
        const float inv = 1.f / sqrtf(j.s3[i] + j.eps);
        const float aa = j.s0[i] * inv;
        ((float*)j.d0)[i] = aa;
        ((float*)j.d1)[i] = j.s1[i] - j.s2[i] * aa;
      }
    }
  }
}

}  // namespace

int kbk_pack_batched(const PackJob* jobs, int n_jobs, cudaStream_t st) {
  for (int first = 0; first < n_jobs; first += kJobsPerLaunch) {
    const int nj = n_jobs - first < kJobsPerLaunch ? n_jobs - first : kJobsPerLaunch;
    PackBatch b;
    memset(&b, 0, sizeof(b));
    long long max_count = 0;
    for (int k = 0; k < nj; ++k) {
      b.jobs[k] = jobs[first + k];
      if (b.jobs[k].count > max_count) max_count = b.jobs[k].count;
    }
    if (max_count == 0) continue;
    const dim3 grid((unsigned)((max_count + kChunk - 1) / kChunk), (unsigned)nj);
    pack_batched_kernel<<<grid, 256, 0, st>>>(b);
    KB_CUDA_LAUNCH_CHECK();
  }
  return KB_OK;
}
