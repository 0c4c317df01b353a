// gemm_simt.cu — small dense layers of the model on CUDA cores, fp32 accumulate.
//
// Covers the Linear / 1x1-conv layers around the trunk (reference se_resnet.py:57-61 global_fc,
// :63-66 SE, :119-130 heads) forward and backward: tiny GEMMs (M = boards or pixels) whose FLOPs
// are ~0.2 % of the model. One generic 64x64x16 register-tiled kernel with optional operand
// transposes, a per-k affine(+ReLU) prologue on A (BatchNorm+ReLU feeding the policy conv),
// bias/ReLU/mask epilogue, board-pitched rows (the padded (B, 11264) policy buffer) and split-K
// with fp32 atomics (weight gradients reduce over the batch).
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float ld_any(const void* p, int dtype, long long i) {
  return dtype == KB_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}

__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g, int k_per_slice) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_slice;
  const int k_end = min(g.K, k_begin + k_per_slice);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- A tile (BM x BK) ----
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + it * 256;
      int m, k;
      if (g.transA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < k_end) {
        const long long row = g.transA ? gk : gm, col = g.transA ? gm : gk;
        long long off;
        if (g.a_group_rows > 0) off = (row / g.a_group_rows) * g.a_group_pitch + (row % g.a_group_rows) * g.lda + col;
        else off = row * g.lda + col;
        v = ld_any(g.A, g.a_dtype, off);
        if (g.a_pa) v = fmaf(v, g.a_pa[gk], g.a_pb[gk]);
        if (g.a_relu) v = fmaxf(v, 0.f);
      }
      As[k][m] = v;
    }
    // ---- B tile (BK x BN) ----
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + it * 256;
      int n, k;
      if (g.transB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < k_end) v = ld_any(g.B, g.b_dtype, g.transB ? (long long)gn * g.ldb + gk : (long long)gk * g.ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
    long long crow;
    if (g.c_group_rows > 0) crow = ((long long)gm / g.c_group_rows) * g.c_group_pitch + ((long long)gm % g.c_group_rows) * g.ldc;
    else crow = (long long)gm * g.ldc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += g.bias[gn];
      if (g.splitk > 1) {
        atomicAdd(((float*)g.C) + crow + gn, v);
      } else {
        if (g.relu) v = fmaxf(v, 0.f);
        if (g.mask_src && !(g.mask_src[(long long)gm * g.ld_mask + gn] > 0.f)) v = 0.f;
        if (g.c_dtype == KB_F32) ((float*)g.C)[crow + gn] = v;
        else ((bf16*)g.C)[crow + gn] = __float2bfloat16_rn(v);
      }
    }
  }
}

// out[n] += sum_m X[m][n]; block (32, 8)
__global__ void colsum_kernel(const void* __restrict__ X, int dtype, long long ldx, int group_rows, long long group_pitch,
                              int M, int N, int rows_per_block, float* out) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int n = blockIdx.y * 32 + cx;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + ry; r < r1; r += 8) {
      long long off;
      if (group_rows > 0) off = ((long long)r / group_rows) * group_pitch + ((long long)r % group_rows) * ldx + n;
      else off = (long long)r * ldx + n;
      s += ld_any(X, dtype, off);
    }
  sh[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += sh[i][cx];
    atomicAdd(&out[n], a);
  }
}

}  // namespace

int kbk_gemm(const GemmArgs& g, cudaStream_t st) {
  KB_CHECK_ARG(g.M >= 0 && g.N > 0 && g.K > 0, "gemm: bad shape M=%d N=%d K=%d", g.M, g.N, g.K);
  if (g.M == 0) return KB_OK;
  int splitk = g.splitk < 1 ? 1 : g.splitk;
  KB_CHECK_ARG(splitk == 1 || (g.c_dtype == KB_F32 && !g.relu && !g.mask_src), "gemm: split-K needs a plain fp32 output");
  int kps = kb_ceil_div(g.K, splitk);
  kps = kb_ceil_div(kps, BK) * BK;
  splitk = kb_ceil_div(g.K, kps);
  GemmArgs a = g;
  a.splitk = g.splitk > 1 ? 2 : 1;  // >1 only selects the atomic epilogue
  gemm_kernel<<<dim3(kb_ceil_div(g.N, BN), kb_ceil_div(g.M, BM), splitk), 256, 0, st>>>(a, kps);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_colsum(const void* X, int dtype, long long ldx, int group_rows, long long group_pitch, int M, int N, float* out,
               cudaStream_t st) {
  if (M == 0) return KB_OK;
  const int rpb = 1024;
  colsum_kernel<<<dim3(kb_ceil_div(M, rpb), kb_ceil_div(N, 32)), 256, 0, st>>>(X, dtype, ldx, group_rows, group_pitch, M, N, rpb, out);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
