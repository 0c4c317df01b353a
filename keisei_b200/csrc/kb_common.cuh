// kb_common.cuh — shared device/host helpers for the keisei_b200 kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#define KB_OK 0
#define KB_ERR_INVALID (-1)
#define KB_ERR_CUDA (-2)
#define KB_ERR_WORKSPACE (-3)
#define KB_ERR_UNSUPPORTED (-4)

// dtype codes used across the C-ABI
#define KB_F32 0
#define KB_BF16 1

void kb_set_error(const char* fmt, ...);
void kb_count_launch(void);  // bumps the library-wide kernel-launch counter (kb_launch_count)

#define KB_CHECK_ARG(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      kb_set_error(__VA_ARGS__);               \
      return KB_ERR_INVALID;                   \
    }                                          \
  } while (0)

#define KB_CUDA_LAUNCH_CHECK()                                                         \
  do {                                                                                 \
    kb_count_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      kb_set_error("%s:%d CUDA launch error: %s", __FILE__, __LINE__,                  \
                   cudaGetErrorString(e__));                                           \
      return KB_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define KB_CUDA_CHECK(expr)                                                            \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      kb_set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #expr,                   \
                   cudaGetErrorString(e__));                                           \
      return KB_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

static inline int kb_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Kernels meant to run NEXT TO a tcgen05 convolution (side-stream weight gradients, two-branch rollout) ask for the
// same shared-memory carve-out as the convolution (maximum shared memory): an SM only hosts CTAs of kernels whose
// L1 / shared split agrees, so a streaming kernel with the default (maximum L1) preference would wait for the SM to drain.
// Costs 10-15 % when the kernel runs ALONE (measured on the backward column kernels: less L1 for the write path), so it is
// set only where the co-residency is measured to pay: the evaluation tail kernels of the two-branch rollout.
template <typename K>
static inline void kb_prefer_max_smem_carveout(K kernel) {
  cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float kb_to_float(T v);
template <> __device__ __forceinline__ float kb_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float kb_to_float<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T kb_from_float(float v);
template <> __device__ __forceinline__ float kb_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 kb_from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float kb_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double kb_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float kb_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions. `scratch` must hold >= 32 floats. All threads get the result.
__device__ __forceinline__ float kb_block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = kb_warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? scratch[lane] : 0.f;
  r = kb_warp_sum(r);
  return r;
}
__device__ __forceinline__ float kb_block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = kb_warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? scratch[lane] : -INFINITY;
  r = kb_warp_max(r);
  return r;
}
__device__ __forceinline__ double kb_block_sum_d(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = kb_warp_sum_d(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = (lane < nwarps) ? scratch[lane] : 0.0;
  r = kb_warp_sum_d(r);
  return r;
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011). Stateless: (key, counter) -> 4x u32.
struct kb_philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ kb_philox4 kb_philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
  uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  kb_philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}
// uniform in the OPEN interval (0,1): 23 random bits + 0.5 — the largest value, (2^23 - 0.5) / 2^23 = 0.99999994, is
// representable in fp32 (with 24 bits the top value rounds to exactly 1.0 and -log(-log(u)) becomes +inf)
__device__ __forceinline__ float kb_u32_to_unit(uint32_t x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }


// Programmatic dependent launch for the kernels that call tcptx::pdl_wait(): OFF by default, KB_PDL=1 switches it on.
// Measured on the graph-replayed rollout (same box, alternating runs): 27.2-27.4 ms with, 26.8-27.1 ms without at 4096
// boards; 5.9-6.6 vs 5.6-5.7 ms at 512 — the persistent one-CTA-per-SM kernels leave no room for a dependent grid's CTAs
// until their own exit, and the early CTAs then sit in griddepcontrol.wait on SMs the tail of the previous grid shares.
inline bool kb_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("KB_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v != 0;
}
// <<<grid, block, smem, stream>>> with the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t kb_launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = kb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#endif  // __CUDACC__
