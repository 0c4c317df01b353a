// resnet_heads.cu — the policy / value head front ends of the plain ResNet (reference
// keisei/training/models/resnet.py:49-59, 76-84): two 1x1 convolutions with 2 + 1 output channels,
// BatchNorm, ReLU and the NCHW flatten that feeds policy_fc / value_fc1, plus their backward.
//
// All three output channels are produced in ONE pass over the trunk output (HBM-bound: the
// [B][81][C] activation is read once forward and once backward, its gradient written once):
// warp = pixel, lane = 4-channel vectors, three dot products reduced with warp shuffles.
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
};

constexpr int kMaxVec = 8;  // 4-channel vectors per lane (template NV <= 8): C <= 32 * 4 * 8 = 1024

// raw[(b*81+p)*3 + j] = sum_c x[b][p][c] * w_j[c]   (j = 0,1: policy_conv rows, j = 2: value_conv)
// sums (optional): double [6] = {sum p0, sum p1, sumsq p0, sumsq p1, sum v, sumsq v}
template <typename T, int NV>
__global__ void __launch_bounds__(256) resnet_head_conv_kernel(const T* __restrict__ x, const float* __restrict__ wp,
                                                                 const float* __restrict__ wv, float* __restrict__ raw,
                                                                 int B, int C, double* sums) {
  __shared__ float red[8][6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w[3][NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = c < C;
      w[0][k][i] = ok ? wp[c + i] : 0.f; w[1][k][i] = ok ? wp[C + c + i] : 0.f; w[2][k][i] = ok ? wv[c + i] : 0.f;
    }
  }
  float s[3] = {0.f, 0.f, 0.f}, q[3] = {0.f, 0.f, 0.f};
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int p = warp; p < 81; p += 8) {
      const T* row = x + ((size_t)b * 81 + p) * C;
      float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * 4;
        if (c < C) {
          float v[4];
          Vec4<T>::load(row + c, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) { d[0] = fmaf(v[i], w[0][k][i], d[0]); d[1] = fmaf(v[i], w[1][k][i], d[1]); d[2] = fmaf(v[i], w[2][k][i], d[2]); }
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) d[j] = kb_warp_sum(d[j]);
      if (lane == 0) {
        float* o = raw + ((size_t)b * 81 + p) * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) { o[j] = d[j]; s[j] += d[j]; q[j] = fmaf(d[j], d[j], q[j]); }
      }
    }
  }
  if (sums == nullptr) return;
  if (lane == 0) { for (int j = 0; j < 3; ++j) { red[warp][j] = s[j]; red[warp][3 + j] = q[j]; } }
  __syncthreads();
  if (threadIdx.x < 6) {
    double a = 0.0;
    for (int wi = 0; wi < 8; ++wi) a += (double)red[wi][threadIdx.x];
    // threadIdx.x: 0,1,2 = sums of p0,p1,v; 3,4,5 = sums of squares
    const int slot[6] = {0, 1, 4, 2, 3, 5};
    atomicAdd(&sums[slot[threadIdx.x]], a);
  }
}

// BatchNorm affine + ReLU + NCHW flatten: p_flat[b][c*81+p] (fp32, optional bf16 copy with row pitch
// pitch_bf whose columns 162.. are zeroed), v_flat[b][p]. One thread per (board, pixel).
__global__ void resnet_head_act_kernel(const float* __restrict__ raw, const float* __restrict__ ap, const float* __restrict__ bp,
                                       const float* __restrict__ av, const float* __restrict__ bv, float* __restrict__ p_flat,
                                       bf16* __restrict__ p_flat_bf, int pitch_bf, float* __restrict__ v_flat, long long M) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const long long b = i / 81; const int p = (int)(i - b * 81);
  const float r0 = raw[i * 3], r1 = raw[i * 3 + 1], r2 = raw[i * 3 + 2];
  const float a0 = fmaxf(fmaf(r0, ap[0], bp[0]), 0.f), a1 = fmaxf(fmaf(r1, ap[1], bp[1]), 0.f), a2 = fmaxf(fmaf(r2, av[0], bv[0]), 0.f);
  p_flat[b * 162 + p] = a0;
  p_flat[b * 162 + 81 + p] = a1;
  v_flat[i] = a2;
  if (p_flat_bf) {
    bf16* o = p_flat_bf + b * pitch_bf;
    o[p] = __float2bfloat16_rn(a0);
    o[81 + p] = __float2bfloat16_rn(a1);
    if (162 + p < pitch_bf) o[162 + p] = __float2bfloat16_rn(0.f);
  }
}

// ReLU mask + un-flatten + BatchNorm-backward statistics:
// d3[(b*81+p)*3+j] = dflat_j * [raw_j*a_j+b_j > 0]; sums += {sum d p0, sum d p1, sum d*raw p0, sum d*raw p1, sum d v, sum d*raw v}
__global__ void __launch_bounds__(256) resnet_head_bwd_act_kernel(const float* __restrict__ dp_flat, const float* __restrict__ dv_flat,
                                                                    const float* __restrict__ raw, const float* __restrict__ ap,
                                                                    const float* __restrict__ bp, const float* __restrict__ av,
                                                                    const float* __restrict__ bv, float* __restrict__ d3,
                                                                    long long M, double* sums) {
  __shared__ float red[8][6];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < M) {
    const long long b = i / 81; const int p = (int)(i - b * 81);
    const float r0 = raw[i * 3], r1 = raw[i * 3 + 1], r2 = raw[i * 3 + 2];
    const float g0 = fmaf(r0, ap[0], bp[0]) > 0.f ? dp_flat[b * 162 + p] : 0.f;
    const float g1 = fmaf(r1, ap[1], bp[1]) > 0.f ? dp_flat[b * 162 + 81 + p] : 0.f;
    const float g2 = fmaf(r2, av[0], bv[0]) > 0.f ? dv_flat[i] : 0.f;
    d3[i * 3] = g0; d3[i * 3 + 1] = g1; d3[i * 3 + 2] = g2;
    v[0] = g0; v[1] = g1; v[2] = g0 * r0; v[3] = g1 * r1; v[4] = g2; v[5] = g2 * r2;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = kb_warp_sum(v[j]);
  if (lane == 0) { for (int j = 0; j < 6; ++j) red[warp][j] = v[j]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    double a = 0.0;
    for (int wi = 0; wi < 8; ++wi) a += (double)red[wi][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], a);
  }
}

// dz_j = k1_j*d3_j - k2_j*raw_j - k3_j (BatchNorm backward); dx[b][p][c] = sum_j dz_j * w_j[c];
// dwp[j][c] += sum_rows dz_j * x[c], dwv[c] likewise. k = {k1p[2], k2p[2], k3p[2], k1v, k2v, k3v} as separate arrays.
template <typename T, int NV>
__global__ void __launch_bounds__(256) resnet_head_bwd_x_kernel(const T* __restrict__ x, const float* __restrict__ d3,
                                                                  const float* __restrict__ raw, const float* __restrict__ kp,
                                                                  const float* __restrict__ kv, int kstride,
                                                                  const float* __restrict__ wp, const float* __restrict__ wv,
                                                                  T* __restrict__ dx, float* __restrict__ dwp,
                                                                  float* __restrict__ dwv, int B, int C) {
  extern __shared__ float acc_sh[];  // [3][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) acc_sh[i] = 0.f;
  float w[3][NV][4], acc[3][NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = c < C;
      w[0][k][i] = ok ? wp[c + i] : 0.f; w[1][k][i] = ok ? wp[C + c + i] : 0.f; w[2][k][i] = ok ? wv[c + i] : 0.f;
      acc[0][k][i] = 0.f; acc[1][k][i] = 0.f; acc[2][k][i] = 0.f;
    }
  }
  // kp: [3][kstride] = k1, k2, k3 of the 2 policy channels; kv likewise for the value channel
  const float k1[3] = {kp[0], kp[1], kv[0]}, k2[3] = {kp[kstride], kp[kstride + 1], kv[kstride]},
              k3[3] = {kp[2 * kstride], kp[2 * kstride + 1], kv[2 * kstride]};
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int p = warp; p < 81; p += 8) {
      const size_t r = (size_t)b * 81 + p;
      float dz[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) dz[j] = k1[j] * d3[r * 3 + j] - k2[j] * raw[r * 3 + j] - k3[j];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * 4;
        if (c < C) {
          float v[4], o[4];
          Vec4<T>::load(x + r * C + c, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i] = fmaf(dz[0], w[0][k][i], fmaf(dz[1], w[1][k][i], dz[2] * w[2][k][i]));
            acc[0][k][i] = fmaf(dz[0], v[i], acc[0][k][i]);
            acc[1][k][i] = fmaf(dz[1], v[i], acc[1][k][i]);
            acc[2][k][i] = fmaf(dz[2], v[i], acc[2][k][i]);
          }
          Vec4<T>::store(dx + r * C + c, o);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(&acc_sh[c + i], acc[0][k][i]); atomicAdd(&acc_sh[C + c + i], acc[1][k][i]); atomicAdd(&acc_sh[2 * C + c + i], acc[2][k][i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&dwp[i], acc_sh[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dwv[i], acc_sh[2 * C + i]);
}

__global__ void tanh_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out2, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float y = tanhf(in[i]);
  out[i] = y;
  if (out2) out2[i] = y;
}
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  dx[i] = dy[i] * (1.f - y[i] * y[i]);
}

inline int head_grid(int B) { const int g = 148 * 4; return B < g ? B : g; }

}  // namespace

int kbk_resnet_head_conv(const void* x, int dtype, const float* wp, const float* wv, float* raw, int B, int C, double* sums,
                         cudaStream_t st) {
  KB_CHECK_ARG(C % 4 == 0 && C <= 32 * 4 * kMaxVec, "resnet head: C=%d unsupported", C);
  if (B == 0) return KB_OK;
#define KB_HC(NV_)                                                                                                      \
  do {                                                                                                                  \
    if (dtype == KB_F32) resnet_head_conv_kernel<float, NV_><<<head_grid(B), 256, 0, st>>>((const float*)x, wp, wv, raw, B, C, sums); \
    else resnet_head_conv_kernel<bf16, NV_><<<head_grid(B), 256, 0, st>>>((const bf16*)x, wp, wv, raw, B, C, sums);      \
  } while (0)
  const int nv = (C / 4 + 31) / 32;
  if (nv <= 1) KB_HC(1); else if (nv <= 2) KB_HC(2); else if (nv <= 4) KB_HC(4); else KB_HC(8);
#undef KB_HC
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_resnet_head_act(const float* raw, const float* ap, const float* bp, const float* av, const float* bv, float* p_flat,
                        void* p_flat_bf, int pitch_bf, float* v_flat, int B, cudaStream_t st) {
  const long long M = (long long)B * 81;
  if (M == 0) return KB_OK;
  KB_CHECK_ARG(p_flat_bf == nullptr || (pitch_bf >= 162 && pitch_bf <= 162 + 81), "resnet head: bad bf16 pitch %d", pitch_bf);
  resnet_head_act_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(raw, ap, bp, av, bv, p_flat, (bf16*)p_flat_bf, pitch_bf, v_flat, M);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_resnet_head_bwd_act(const float* dp_flat, const float* dv_flat, const float* raw, const float* ap, const float* bp,
                            const float* av, const float* bv, float* d3, int B, double* sums, cudaStream_t st) {
  const long long M = (long long)B * 81;
  if (M == 0) return KB_OK;
  resnet_head_bwd_act_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(dp_flat, dv_flat, raw, ap, bp, av, bv, d3, M, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_resnet_head_bwd_x(const void* x, int dtype, const float* d3, const float* raw, const float* kp, const float* kv,
                          int kstride, const float* wp, const float* wv, void* dx, float* dwp, float* dwv, int B, int C,
                          cudaStream_t st) {
  KB_CHECK_ARG(C % 4 == 0 && C <= 32 * 4 * kMaxVec, "resnet head: C=%d unsupported", C);
  if (B == 0) return KB_OK;
  const size_t smem = 3 * (size_t)C * sizeof(float);
#define KB_HB(NV_)                                                                                                      \
  do {                                                                                                                  \
    if (dtype == KB_F32)                                                                                                \
      resnet_head_bwd_x_kernel<float, NV_><<<head_grid(B), 256, smem, st>>>((const float*)x, d3, raw, kp, kv, kstride, wp, wv, (float*)dx, dwp, dwv, B, C); \
    else                                                                                                                \
      resnet_head_bwd_x_kernel<bf16, NV_><<<head_grid(B), 256, smem, st>>>((const bf16*)x, d3, raw, kp, kv, kstride, wp, wv, (bf16*)dx, dwp, dwv, B, C); \
  } while (0)
  const int nv = (C / 4 + 31) / 32;
  if (nv <= 1) KB_HB(1); else if (nv <= 2) KB_HB(2); else if (nv <= 4) KB_HB(4); else KB_HB(8);
#undef KB_HB
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_tanh_fwd(const float* in, float* out, float* out2, long long n, cudaStream_t st) {
  if (n == 0) return KB_OK;
  tanh_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, out2, n);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
int kbk_tanh_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st) {
  if (n == 0) return KB_OK;
  tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dy, y, dx, n);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
