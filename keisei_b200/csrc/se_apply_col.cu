// se_apply_col.cu — the tail of a GlobalPoolBiasBlock (bf16 activations) as a tiny MLP kernel plus ONE streaming pass:
//
//   se_mlp_fwd_kernel    se_in = mean_p(bn2(z2));  se = W2 relu(W1 se_in + b1) + b2            (se_resnet.py:83-86)
//   se_apply_col_kernel  x' = relu(bn2(z2) * sigmoid(scale) + shift + x);  pool' = (mean, max, std)(x')   (:87-98)
//
// The streaming pass uses the column layout of the backward kernels (blocks.cu): a thread owns 4 channels of one board
// for all 81 pixels, so the pool statistics (and the tie counts the amax backward needs) are thread-local — no shared
// memory, no barriers — with 3 pixels x 2 streams of independent 8-byte loads in flight per thread. It supersedes the
// TMA-bulk-staged kernel in se_apply.cu (kept, KB_SE_APPLY=tma): that one reached 67-69 % of the measured HBM peak
// because a stage cannot be refilled while its warp group still reads it; this layout is bounded by DRAM only.
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

template <int S>
__device__ __forceinline__ float warp_multi_sum(float (&v)[S], int lane) {  // see se_bwd.cu
#pragma unroll
  for (int h = S / 2; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = S; o < 32; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// thread = channel; W1 [S][C] and W2^T [S][2C] in shared memory; CTAs walk boards with the next board's mean in flight
// WSM = true: both weight matrices staged in shared memory (96 KB for C = 256, S = 32). WSM = false: weights read
// through L1/L2 every board (393 MB of L2 traffic per 4096 boards — noise next to a convolution's 4.8 GB) and only
// 1.3 KB of shared memory, so the kernel can sit on an SM NEXT TO a convolution CTA (two-branch rollout).
template <int S, bool WSM>
__global__ void __launch_bounds__(256) se_mlp_fwd_kernel(SeApplyArgs g) {
  extern __shared__ float sm[];
  const int C = g.C, c = threadIdx.x, lane = c & 31, warp = c >> 5, nwarps = C >> 5;
  float* hs = sm;                  // [S]
  float* red = hs + S;             // [8][S]
  float* W2t = red + 8 * S;        // [S][2C]   (WSM only)
  float* W1s = W2t + S * 2 * C;    // [S][C]    (WSM only)
  if (WSM) {
    for (int i = c; i < 2 * C * S; i += C) { const int j = i / S, s = i - j * S; W2t[s * 2 * C + j] = g.w2[i]; }
    for (int i = c; i < S * C; i += C) W1s[i] = g.w1[i];
  }
  const float* w2a = g.w2 + (size_t)c * S;        // rows of se_fc2 for this channel's scale / shift outputs
  const float* w2b = g.w2 + (size_t)(C + c) * S;
  const float a2 = g.a ? g.a[c] : 1.f, b2 = g.a ? g.b[c] : 0.f;
  const float bias_sc = g.b2[c], bias_sh = g.b2[C + c];
  const float b1 = c < S ? g.b1[c] : 0.f;
  __syncthreads();
  int b = blockIdx.x;
  float m = b < g.B ? g.bmean[(size_t)b * C + c] : 0.f;
  for (; b < g.B; b += gridDim.x) {
    const float v = fmaf(m, a2, b2);
    if (b + (int)gridDim.x < g.B) m = g.bmean[(size_t)(b + gridDim.x) * C + c];
    if (g.se_in_out) g.se_in_out[(size_t)b * C + c] = v;
    float part[S];
#pragma unroll
    for (int s = 0; s < S; ++s) part[s] = (WSM ? W1s[s * C + c] : __ldg(g.w1 + s * C + c)) * v;
    const float ps = warp_multi_sum<S>(part, lane);
    if (lane < S) red[warp * S + lane] = ps;
    __syncthreads();
    if (c < S) {
      float d = b1;
      for (int w = 0; w < nwarps; ++w) d += red[w * S + c];
      d = fmaxf(d, 0.f);
      hs[c] = d;
      if (g.seh_out) g.seh_out[(size_t)b * S + c] = d;
    }
    __syncthreads();
    float sc = bias_sc, sh = bias_sh;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float h = hs[s];
      sc = fmaf(WSM ? W2t[s * 2 * C + c] : __ldg(w2a + s), h, sc);
      sh = fmaf(WSM ? W2t[s * 2 * C + C + c] : __ldg(w2b + s), h, sh);
    }
    // eval (se_raw == 0): the scale half is stored with the sigmoid already applied
    g.se_out[(size_t)b * 2 * C + c] = g.se_raw ? sc : sigmoid_f(sc);
    g.se_out[(size_t)b * 2 * C + C + c] = sh;
  }
}

__device__ __forceinline__ void ld4_bf16(const bf16* p, float (&v)[4]) {
  uint2 u;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p) : "memory");
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int kPix = 3;  // 81 = 27 x 3

template <bool TIES, int kPix>
__global__ void __launch_bounds__(256, kPix == 3 ? 3 : 1) se_apply_col_kernel(SeApplyArgs g) {
  const int C = g.C, TPB = C / 4, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * 4;
  float a_[4] = {1.f, 1.f, 1.f, 1.f}, b_[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.a) {
    const float4 a4 = ldg4(g.a + c0), b4 = ldg4(g.b + c0);
    a_[0] = a4.x; a_[1] = a4.y; a_[2] = a4.z; a_[3] = a4.w; b_[0] = b4.x; b_[1] = b4.y; b_[2] = b4.z; b_[3] = b4.w;
  }
  for (int b = blockIdx.x * BPC + slot; b < g.B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
    float sg[4], sf[4];
    {
      const float4 s4 = ldg4(g.se_out + (size_t)b * 2 * C + c0), h4 = ldg4(g.se_out + (size_t)b * 2 * C + C + c0);
      float sig[4] = {s4.x, s4.y, s4.z, s4.w};
      const float sh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (g.se_raw) sig[k] = sigmoid_f(sig[k]);
        sg[k] = a_[k] * sig[k];                 // (z*a+b)*sig + shift = z*(a*sig) + (b*sig + shift)
        sf[k] = fmaf(b_[k], sig[k], sh[k]);
      }
    }
    // statistics: sums of the fp32 outputs shifted by the pixel-0 value (a constant board gives an exact zero
    // variance); max and tie counts on the STORED (bf16-rounded) values — the backward compares x == max
    float k0[4], ds[4], dss[4];
    __nv_bfloat162 mx2[2], tie2[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { k0[k] = 0.f; ds[k] = 0.f; dss[k] = 0.f; }
    mx2[0] = mx2[1] = __float2bfloat162_rn(0.f);   // outputs are >= 0
    tie2[0] = tie2[1] = __float2bfloat162_rn(0.f);
#pragma unroll 1
    for (int p = 0; p < 81; p += kPix) {
      float zv[kPix][4], rv[kPix][4];
#pragma unroll
      for (int j = 0; j < kPix; ++j) {
        ld4_bf16(g.z + base + (size_t)(p + j) * C, zv[j]);
        ld4_bf16(g.res + base + (size_t)(p + j) * C, rv[j]);
      }
#pragma unroll
      for (int j = 0; j < kPix; ++j) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = fmaxf(fmaf(zv[j][k], sg[k], sf[k]) + rv[j][k], 0.f);
        __nv_bfloat162 pk[2] = {__floats2bfloat162_rn(o[0], o[1]), __floats2bfloat162_rn(o[2], o[3])};
        uint2 st;
        st.x = *reinterpret_cast<const uint32_t*>(&pk[0]); st.y = *reinterpret_cast<const uint32_t*>(&pk[1]);
        *reinterpret_cast<uint2*>(g.out + base + (size_t)(p + j) * C) = st;
        if (j == 0 && p == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) k0[k] = o[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = o[k] - k0[k];
          ds[k] += d; dss[k] = fmaf(d, d, dss[k]);
        }
        // running max / tie count on the stored pairs with packed bf16x2 instructions (counts <= 81: exact in bf16)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const __nv_bfloat162 nm = __hmax2(mx2[h], pk[h]);
          if (TIES) tie2[h] = __hfma2(tie2[h], __heq2(mx2[h], nm), __heq2(pk[h], nm));  // count*[max unchanged] + [value == max]
          mx2[h] = nm;
        }
      }
    }
    const float mx[4] = {__low2float(mx2[0]), __high2float(mx2[0]), __low2float(mx2[1]), __high2float(mx2[1])};
    const float tie[4] = {__low2float(tie2[0]), __high2float(tie2[0]), __low2float(tie2[1]), __high2float(tie2[1])};
    float mean[4], sd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dm = ds[k] * (1.f / 81.f);
      mean[k] = k0[k] + dm;
      sd[k] = sqrtf(fmaxf(dss[k] * (1.f / 81.f) - dm * dm, 0.f));
    }
    float* pr = g.pool + (size_t)b * 3 * C + c0;
    *reinterpret_cast<float4*>(pr) = make_float4(mean[0], mean[1], mean[2], mean[3]);
    *reinterpret_cast<float4*>(pr + C) = make_float4(mx[0], mx[1], mx[2], mx[3]);
    *reinterpret_cast<float4*>(pr + 2 * C) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    if (g.pool_bf) {
      bf16* pb = g.pool_bf + (size_t)b * 3 * C + c0;
      auto st4 = [](bf16* q, const float (&v)[4]) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(q) = u;
      };
      st4(pb, mean); st4(pb + C, mx); st4(pb + 2 * C, sd);
    }
    if (TIES) *reinterpret_cast<float4*>(g.ties + (size_t)b * C + c0) = make_float4(tie[0], tie[1], tie[2], tie[3]);
  }
}

size_t mlp_smem(int C, int S) { return (size_t)(S * 2 * C + S * C + S + 8 * S) * sizeof(float); }

template <int S>
int launch_mlp(const SeApplyArgs& a, int num_sms, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(se_mlp_fwd_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  if (a.ties == nullptr && a.B >= 1000) {
    // evaluation half batches of the two-branch rollout: small enough to share an SM with a convolution CTA of the other
    // branch (bit-identical to the shared-memory variant: same arithmetic, same order, only the weight source differs)
    const int grid = a.B < 4 * num_sms ? a.B : 4 * num_sms;
    kb_prefer_max_smem_carveout(se_mlp_fwd_kernel<S, false>);
    se_mlp_fwd_kernel<S, false><<<grid, a.C, (size_t)(9 * S) * sizeof(float), st>>>(a);
  } else {
    const int grid = a.B < 2 * num_sms ? a.B : 2 * num_sms;
    se_mlp_fwd_kernel<S, true><<<grid, a.C, mlp_smem(a.C, S), st>>>(a);
  }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

}  // namespace

int kbk_se_apply_col_supported(int C, int S) {
  return C % 32 == 0 && C >= 32 && C <= 256 && 256 % (C / 4) == 0 && (S == 4 || S == 8 || S == 16 || S == 32) &&
         mlp_smem(C, S) <= 100 * 1024;
}

int kbk_se_apply_col(const SeApplyArgs& a, int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(kbk_se_apply_col_supported(a.C, a.S), "se_apply_col: unsupported shape C=%d S=%d", a.C, a.S);
  KB_CHECK_ARG(a.z && a.res && a.out && a.bmean && a.w1 && a.b1 && a.w2 && a.b2 && a.pool && a.se_out, "se_apply: null pointer");
  if (a.B == 0) return KB_OK;
  if (num_sms <= 0) num_sms = 148;
  int r;
  switch (a.S) {
    case 4: r = launch_mlp<4>(a, num_sms, st); break;
    case 8: r = launch_mlp<8>(a, num_sms, st); break;
    case 16: r = launch_mlp<16>(a, num_sms, st); break;
    default: r = launch_mlp<32>(a, num_sms, st); break;
  }
  if (r != KB_OK) return r;
  const int bpc = 256 / (a.C / 4);
  int grid = kb_ceil_div(a.B, bpc);
  if (grid > num_sms * 8) grid = num_sms * 8;
  // small launches are latency-bound on the 27 dependent 3-pixel round trips of a thread: 9 pixels per trip (blocks.cu: col_small)
  const bool small = kb_ceil_div(a.B, bpc) <= 2 * num_sms;
  if (a.ties) { if (small) se_apply_col_kernel<true, 9><<<grid, 256, 0, st>>>(a); else se_apply_col_kernel<true, 3><<<grid, 256, 0, st>>>(a); }
  else { kb_prefer_max_smem_carveout(se_apply_col_kernel<false, 3>); se_apply_col_kernel<false, 3><<<grid, 256, 0, st>>>(a); }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
