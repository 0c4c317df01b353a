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

// Conv pack of one 32 (co) x 32 (ci) x 9 (tap) brick through shared memory: the fp32 source (co, ci, tap) is read as 32
// contiguous runs of 288 floats, the forward pack [co][tap][ci] and the flipped data-gradient pack [ci][8 - tap][co]
// are both written as 32-element contiguous runs. (The element-wise form wrote the flipped pack as 2-byte stores
// 9 * Cout elements apart — one 32-byte sector per element: the two launches of a 40 x 256 re-pack took 0.51 ms for
// 326 MB, a tenth of the copy bandwidth, after EVERY optimiser step.)
constexpr int kTileC = 32;
template <typename T>
__device__ __forceinline__ void conv_tile(const PackJob& j, int tile_idx, float (*tile)[kTileC * 9 + 1]) {
  const int Cout = j.n0, Cin = j.n1, Cinp = j.n2;
  const int ci_tiles = Cinp / kTileC;
  const int co0 = (tile_idx / ci_tiles) * kTileC, ci0 = (tile_idx % ci_tiles) * kTileC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int run = ci0 < Cin ? (Cin - ci0 < kTileC ? Cin - ci0 : kTileC) * 9 : 0;   // valid floats of a source row
  for (int co_l = wid; co_l < kTileC; co_l += 8) {
    const float* src = j.s0 + ((size_t)(co0 + co_l) * Cin + ci0) * 9;
    for (int r = lane; r < kTileC * 9; r += 32) tile[co_l][r] = r < run ? src[r] : 0.f;
  }
  __syncthreads();
  T* d0 = (T*)j.d0;
  for (int idx = threadIdx.x; idx < kTileC * 9 * kTileC; idx += 256) {
    const int ci_l = idx % kTileC, tap = (idx / kTileC) % 9, co_l = idx / (kTileC * 9);
    d0[((size_t)(co0 + co_l) * 9 + tap) * Cinp + ci0 + ci_l] = kb_from_float<T>(tile[co_l][ci_l * 9 + tap]);
  }
  if (j.d1 != nullptr) {
    T* d1 = (T*)j.d1;
    for (int idx = threadIdx.x; idx < kTileC * 9 * kTileC; idx += 256) {
      const int co_l = idx % kTileC, tap = (idx / kTileC) % 9, ci_l = idx / (kTileC * 9);
      d1[((size_t)(ci0 + ci_l) * 9 + (8 - tap)) * Cout + co0 + co_l] = kb_from_float<T>(tile[co_l][ci_l * 9 + tap]);
    }
  }
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const __grid_constant__ PackBatch batch, int tiled) {
  __shared__ float tile[kTileC][kTileC * 9 + 1];
  const PackJob& j = batch.jobs[blockIdx.y];
  if (tiled && j.kind == KB_PACK_CONV && j.n0 % kTileC == 0 && j.n2 % kTileC == 0) {   // block-uniform
    const int tiles = (j.n0 / kTileC) * (j.n2 / kTileC);
    if ((int)blockIdx.x >= tiles) return;
    if (j.dtype == KB_F32) conv_tile<float>(j, blockIdx.x, tile); else conv_tile<bf16>(j, blockIdx.x, tile);
    return;
  }
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
    static const int tiled = [] { const char* e = getenv("KB_PACK_TILED"); return (e && e[0] == '0') ? 0 : 1; }();
    pack_batched_kernel<<<grid, 256, 0, st>>>(b, tiled);
    KB_CUDA_LAUNCH_CHECK();
  }
  return KB_OK;
}
