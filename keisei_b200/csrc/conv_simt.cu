// conv_simt.cu — fp32-accurate SIMT 3x3 convolution kernels on NHWC 9x9 boards, plus the
// weight / observation packing kernels shared with the tcgen05 path.
//
// These are the *accurate* path (fp32 FMA accumulation, any channel count that is a multiple of 4)
// used for the fp32 parity bar (1e-4 relative through 81 stacked convolutions cannot be met by a
// single bf16/TF32 tensor-core pass) and for small research models; the bf16 throughput path is
// conv_tc.cu. Both share the thread-owns-a-channel epilogue in conv_epilogue.cuh.
//
// Reference ops replaced: F.conv2d(padding=1, bias=False) at se_resnet.py:50,52,110 (forward) and
// its autograd (dgrad = the same kernel on flipped/transposed weights; wgrad below).
#include "kb_common.cuh"
#include "conv_epilogue.cuh"
#include "kb_kernels.h"

namespace {

// --------------------------------------------------------------------------------------------
// packing
// --------------------------------------------------------------------------------------------
// w  [Cout][Cin][3][3] fp32 (PyTorch)  ->  wf [Cout][9][Cinp]  (forward, k = tap*Cinp + ci)
//                                          wd [Cinp][9][Cout]  (dgrad: taps flipped, channels swapped)
template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd,
                                        int Cout, int Cin, int Cinp) {
  const long long n = (long long)Cout * 9 * Cinp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cinp);
    const int tap = (int)((i / Cinp) % 9);
    const int co = (int)(i / ((long long)Cinp * 9));
    const float v = ci < Cin ? w[((size_t)co * Cin + ci) * 9 + tap] : 0.f;
    wf[i] = kb_from_float<T>(v);
    if (wd != nullptr) wd[((size_t)ci * 9 + (8 - tap)) * Cout + co] = kb_from_float<T>(v);
  }
}

// obs [B][Cin][81] fp32 (NCHW) -> [B][81][Cinp] T (NHWC, zero padded channels)
template <typename T>
__global__ void __launch_bounds__(256) pack_obs_kernel(const float* __restrict__ obs, T* __restrict__ out,
                                                        int Cin, int Cinp) {
  extern __shared__ float s_obs[];  // [Cin][81]
  const int b = blockIdx.x;
  const float* src = obs + (size_t)b * Cin * 81;
  for (int i = threadIdx.x; i < Cin * 81; i += blockDim.x) s_obs[i] = src[i];
  __syncthreads();
  T* dst = out + (size_t)b * 81 * Cinp;
  for (int i = threadIdx.x; i < 81 * Cinp; i += blockDim.x) {
    const int ci = i % Cinp, p = i / Cinp;
    dst[i] = kb_from_float<T>(ci < Cin ? s_obs[ci * 81 + p] : 0.f);
  }
}

// --------------------------------------------------------------------------------------------
// forward / dgrad: one CTA = one board x CT output channels, thread = output channel
// --------------------------------------------------------------------------------------------
template <typename T, int CT>
__global__ void __launch_bounds__(CT) conv3x3_simt_kernel(const T* __restrict__ in, const T* __restrict__ w,
                                                          T* __restrict__ out, int B, int Cin, int Cout,
                                                          ConvEpi epi) {
  __shared__ __align__(16) float s_in[121 * 32];  // padded 11x11 board, 32-channel chunk
  const int b = blockIdx.x;
  const int c = blockIdx.y * CT + threadIdx.x;
  const bool c_ok = c < Cout;
  float acc[81];
#pragma unroll
  for (int p = 0; p < 81; ++p) acc[p] = 0.f;
  for (int i = threadIdx.x; i < 121 * 32; i += CT) s_in[i] = 0.f;  // halo stays zero for the whole kernel
  for (int cin0 = 0; cin0 < Cin; cin0 += 32) {
    const int nci = min(32, Cin - cin0);
    __syncthreads();
    for (int i = threadIdx.x; i < 81 * 32; i += CT) {
      const int ci = i & 31, p = i >> 5;
      const int pos = (p / 9 + 1) * 11 + (p % 9 + 1);
      s_in[pos * 32 + ci] = ci < nci ? kb_to_float<T>(in[((size_t)b * 81 + p) * Cin + cin0 + ci]) : 0.f;
    }
    __syncthreads();
    if (c_ok) {
      for (int tap = 0; tap < 9; ++tap) {
        const int toff = (tap / 3) * 11 + (tap % 3);
        const T* wrow = w + ((size_t)c * 9 + tap) * Cin + cin0;
        for (int ci4 = 0; ci4 < nci; ci4 += 4) {
          const float w0 = kb_to_float<T>(wrow[ci4]), w1 = kb_to_float<T>(wrow[ci4 + 1]);
          const float w2 = kb_to_float<T>(wrow[ci4 + 2]), w3 = kb_to_float<T>(wrow[ci4 + 3]);
          const float4* sp = reinterpret_cast<const float4*>(s_in + toff * 32 + ci4);
#pragma unroll
          for (int p = 0; p < 81; ++p) {
            const float4 v = sp[((p / 9) * 11 + (p % 9)) * 8];
            acc[p] = fmaf(w0, v.x, fmaf(w1, v.y, fmaf(w2, v.z, fmaf(w3, v.w, acc[p]))));
          }
        }
      }
    }
  }
  if (c_ok) {
    ConvEpiThread<T, 1, kEpiDynamic> et(epi, c, Cout, B);
    et.begin_board(b, out);
#pragma unroll
    for (int p = 0; p < 81; ++p) {
      const float ms = epi.mask_src ? kb_to_float<T>(((const T*)epi.mask_src)[et.index(b, p)]) : 0.f;
      et.value(0, p, acc[p], ms);
    }
    et.board_done(0, b);
    et.finish(1);
  }
}

// --------------------------------------------------------------------------------------------
// wgrad: dW[co][ci][tap] += sum_b sum_p dY[b][p][co] * X[b][p + shift(tap)][ci]
// CTA tile = 32 co x 32 ci x 9 taps over a slice of the boards; fp32 atomics into the PyTorch
// gradient layout [Cout][Cin_true][3][3].
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv3x3_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  float* __restrict__ dw, int B, int Cin, int Cout,
                                                                  int Cin_true, int boards_per_slice) {
  __shared__ float s_x[121 * 32];
  __shared__ __align__(16) float s_dy[81 * 32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int b_begin = blockIdx.z * boards_per_slice;
  const int b_end = min(B, b_begin + boards_per_slice);
  float acc[4][9];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  for (int i = threadIdx.x; i < 121 * 32; i += 256) s_x[i] = 0.f;
  for (int b = b_begin; b < b_end; ++b) {
    __syncthreads();
    for (int i = threadIdx.x; i < 81 * 32; i += 256) {
      const int cc = i & 31, p = i >> 5;
      const int pos = (p / 9 + 1) * 11 + (p % 9 + 1);
      s_x[pos * 32 + cc] = (ci0 + cc < Cin) ? kb_to_float<T>(x[((size_t)b * 81 + p) * Cin + ci0 + cc]) : 0.f;
      s_dy[p * 32 + cc] = (co0 + cc < Cout) ? kb_to_float<T>(dy[((size_t)b * 81 + p) * Cout + co0 + cc]) : 0.f;
    }
    __syncthreads();
#pragma unroll 3
    for (int p = 0; p < 81; ++p) {
      const float4 d = *reinterpret_cast<const float4*>(s_dy + p * 32 + ty * 4);
      const int base = ((p / 9) * 11 + (p % 9)) * 32 + tx;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = s_x[base + ((t / 3) * 11 + (t % 3)) * 32];
        acc[0][t] = fmaf(d.x, xv, acc[0][t]);
        acc[1][t] = fmaf(d.y, xv, acc[1][t]);
        acc[2][t] = fmaf(d.z, xv, acc[2][t]);
        acc[3][t] = fmaf(d.w, xv, acc[3][t]);
      }
    }
  }
  const int ci = ci0 + tx;
  if (ci < Cin_true) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty * 4 + j;
      if (co < Cout) {
        float* dst = dw + ((size_t)co * Cin_true + ci) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[j][t]);
      }
    }
  }
}

template <typename T>
int launch_conv_simt(const T* in, const T* w, T* out, int B, int Cin, int Cout, const ConvEpi& epi, cudaStream_t st) {
  if (Cout >= 128) {
    conv3x3_simt_kernel<T, 128><<<dim3(B, kb_ceil_div(Cout, 128)), 128, 0, st>>>(in, w, out, B, Cin, Cout, epi);
  } else if (Cout >= 64) {
    conv3x3_simt_kernel<T, 64><<<dim3(B, kb_ceil_div(Cout, 64)), 64, 0, st>>>(in, w, out, B, Cin, Cout, epi);
  } else {
    conv3x3_simt_kernel<T, 32><<<dim3(B, kb_ceil_div(Cout, 32)), 32, 0, st>>>(in, w, out, B, Cin, Cout, epi);
  }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

}  // namespace

int kbk_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cinp, int dtype, cudaStream_t st) {
  const long long n = (long long)Cout * 9 * Cinp;
  const int grid = (int)min((long long)1184, (n + 255) / 256);
  if (dtype == KB_F32) pack_conv_weight_kernel<float><<<grid, 256, 0, st>>>(w, (float*)wf, (float*)wd, Cout, Cin, Cinp);
  else pack_conv_weight_kernel<bf16><<<grid, 256, 0, st>>>(w, (bf16*)wf, (bf16*)wd, Cout, Cin, Cinp);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_pack_obs(const float* obs, void* out, int B, int Cin, int Cinp, int dtype, cudaStream_t st) {
  const size_t smem = (size_t)Cin * 81 * sizeof(float);
  KB_CHECK_ARG(smem <= 48 * 1024, "pack_obs: obs_channels %d too large", Cin);
  if (dtype == KB_F32) pack_obs_kernel<float><<<B, 256, smem, st>>>(obs, (float*)out, Cin, Cinp);
  else pack_obs_kernel<bf16><<<B, 256, smem, st>>>(obs, (bf16*)out, Cin, Cinp);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_conv3x3_simt(const void* in, const void* w, void* out, int B, int Cin, int Cout, int dtype,
                     const ConvEpi& epi, cudaStream_t st) {
  KB_CHECK_ARG(Cin % 4 == 0, "conv3x3_simt: Cin=%d must be a multiple of 4", Cin);
  if (B == 0) return KB_OK;
  if (dtype == KB_F32) return launch_conv_simt<float>((const float*)in, (const float*)w, (float*)out, B, Cin, Cout, epi, st);
  return launch_conv_simt<bf16>((const bf16*)in, (const bf16*)w, (bf16*)out, B, Cin, Cout, epi, st);
}

int kbk_conv3x3_wgrad_simt(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true,
                           int dtype, cudaStream_t st) {
  if (B == 0) return KB_OK;
  const int tiles = kb_ceil_div(Cin, 32) * kb_ceil_div(Cout, 32);
  int slices = kb_ceil_div(148 * 4, tiles);
  if (slices > B) slices = B;
  if (slices < 1) slices = 1;
  const int bps = kb_ceil_div(B, slices);
  slices = kb_ceil_div(B, bps);
  const dim3 grid(kb_ceil_div(Cin, 32), kb_ceil_div(Cout, 32), slices);
  if (dtype == KB_F32)
    conv3x3_wgrad_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, dw, B, Cin, Cout, Cin_true, bps);
  else
    conv3x3_wgrad_simt_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, B, Cin, Cout, Cin_true, bps);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
