// blocks.cu — HBM-bound elementwise + reduction kernels of the SE-ResNet block, forward and
// backward: BatchNorm finalize/apply, SE scale+shift, residual, ReLU, global-pool statistics and
// their gradients (reference se_resnet.py:68-98 and the autograd of those lines).
//
// Layout: NHWC activations [B][81][C]; one CTA per board, one thread per channel, so every
// per-(board, channel) reduction over the 81 pixels is a register loop with 2*C-byte coalesced rows.
#include <stdlib.h>
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <typename T>
__global__ void apply_kernel(ApplyArgs g) {
  const int b = blockIdx.x, c = threadIdx.x, C = g.C;
  if (c >= C) return;
  const float a_ = g.a ? g.a[c] : 1.f, b_ = g.a ? g.b[c] : 0.f;
  const float sg = g.se ? sigmoidf_(g.se[(size_t)b * 2 * C + c]) : 1.f;
  const float sf = g.se ? g.se[(size_t)b * 2 * C + C + c] : 0.f;
  const float gb = g.gbias ? g.gbias[(size_t)b * C + c] : 0.f;
  const T* z = (const T*)g.z + (size_t)b * 81 * C + c;
  const T* res = g.res ? (const T*)g.res + (size_t)b * 81 * C + c : nullptr;
  T* out = (T*)g.out + (size_t)b * 81 * C + c;
  float s = 0.f, mx = -INFINITY, k0 = 0.f, ds = 0.f, dss = 0.f, tie = 0.f;
#pragma unroll 9
  for (int p = 0; p < 81; ++p) {
    float v = fmaf(kb_to_float<T>(z[(size_t)p * C]), a_, b_);
    v = fmaf(v, sg, sf);
    if (res) v += kb_to_float<T>(res[(size_t)p * C]);
    v = fmaxf(v, 0.f) + gb;
    const T st = kb_from_float<T>(v);
    out[(size_t)p * C] = st;
    const float r = kb_to_float<T>(st);
    if (p == 0) k0 = r;
    const float d = r - k0;
    tie = r > mx ? 1.f : (r == mx ? tie + 1.f : tie);
    s += r; mx = fmaxf(mx, r); ds += d; dss = fmaf(d, d, dss);
  }
  if (g.pool && g.ties) g.ties[(size_t)b * C + c] = tie;
  if (g.pool) {
    const float dm = ds * (1.f / 81.f);
    float* pr = g.pool + (size_t)b * 3 * C;
    const float mean = s * (1.f / 81.f), sd = sqrtf(fmaxf(dss * (1.f / 81.f) - dm * dm, 0.f));
    pr[c] = mean;
    pr[C + c] = mx;
    pr[2 * C + c] = sd;
    if (g.pool_bf) {
      bf16* pb = (bf16*)g.pool_bf + (size_t)b * 3 * C;
      pb[c] = __float2bfloat16_rn(mean); pb[C + c] = __float2bfloat16_rn(mx); pb[2 * C + c] = __float2bfloat16_rn(sd);
    }
  }
}

__global__ void bn_eval_affine_kernel(const float* w, const float* bias, const float* rm, const float* rv, float eps,
                                      int C, float* a, float* b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv = 1.f / sqrtf(rv[c] + eps);
  const float aa = w[c] * inv;
  a[c] = aa;
  b[c] = bias[c] - rm[c] * aa;
}

__global__ void affine_rows_kernel(const float* __restrict__ in, const float* __restrict__ a, const float* __restrict__ b,
                                   float* __restrict__ out, bf16* __restrict__ out_bf, long long n, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const float v = fmaf(in[i], a[c], b[c]);
  out[i] = v;
  if (out_bf) out_bf[i] = __float2bfloat16_rn(v);
}

__global__ void bn_finalize_kernel(double* sums, double count, const float* w, const float* bias, const float* rm,
                                   const float* rv, float* rm_out, float* rv_out, long long* nbt, float momentum, float eps,
                                   int C, float* a, float* b, float* mean_o, float* invstd_o) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  sums[c] = 0.0; sums[C + c] = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float aa = w[c] * invstd;
  a[c] = aa;
  b[c] = bias[c] - (float)mean * aa;
  mean_o[c] = (float)mean;
  invstd_o[c] = invstd;
  if (rm != nullptr) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    // a failed SyncBatchNorm exchange (lost peer: the sums come back NaN) must not be committed to the running statistics
    const bool ok = isfinite(mean) && isfinite(unb);
    rm_out[c] = ok ? (1.f - momentum) * rm[c] + momentum * (float)mean : rm[c];
    rv_out[c] = ok ? (1.f - momentum) * rv[c] + momentum * (float)unb : rv[c];
  }
}

// column sums / sums of squares of a [M][C] fp32 matrix -> double atomics. block (32, 8)
__global__ void rows_stats_kernel(const float* __restrict__ x, long long M, int C, long long rows_per_block, double* sums) {
  __shared__ float sh[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float s = 0.f, q = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) { const float v = x[r * C + c]; s += v; q = fmaf(v, v, q); }
  sh[0][ry][cx] = s; sh[1][ry][cx] = q;
  __syncthreads();
  if (ry == 0 && c < C) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i][cx]; b += sh[1][i][cx]; }
    atomicAdd(&sums[c], a);
    atomicAdd(&sums[C + c], b);
  }
}

// ---------------- backward ----------------
template <typename T>
__global__ void block_bwd_reduce_kernel(BlockBwdArgs g) {
  const int b = blockIdx.x, c = threadIdx.x, C = g.C;
  if (c >= C) return;
  const size_t base = (size_t)b * 81 * C + c;
  const T* dxp = (const T*)g.dxp + base; const T* xp = (const T*)g.xp + base; const T* z2 = (const T*)g.z2 + base;
  float s = 0.f, sz = 0.f;
#pragma unroll 9
  for (int p = 0; p < 81; ++p) {
    const float du = kb_to_float<T>(xp[(size_t)p * C]) > 0.f ? kb_to_float<T>(dxp[(size_t)p * C]) : 0.f;
    s += du;
    sz = fmaf(du, kb_to_float<T>(z2[(size_t)p * C]), sz);
  }
  g.s_du[(size_t)b * C + c] = s;
  g.s_duz[(size_t)b * C + c] = sz;
}

__global__ void se_bwd_prep_kernel(const float* s_du, const float* s_duz, const float* a2, const float* b2,
                                   const float* se, float* dse, int B, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int c = (int)(i % C); const long long b = i / C;
  const float sg = sigmoidf_(se[b * 2 * C + c]);
  const float duzh = fmaf(a2[c], s_duz[i], b2[c] * s_du[i]);  // sum_p du * zhat2
  dse[b * 2 * C + c] = duzh * sg * (1.f - sg);
  dse[b * 2 * C + C + c] = s_du[i];
}

// block (32, 8): channel tile x board slice
__global__ void bn2_bwd_sums_kernel(const float* s_du, const float* s_duz, const float* se, const float* dmean,
                                    const float* bsum2, int B, int C, int boards_per_block, double* sums) {
  __shared__ float sh[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const int b0 = blockIdx.x * boards_per_block, b1 = min(B, b0 + boards_per_block);
  float s1 = 0.f, s2 = 0.f;
  if (c < C)
    for (int b = b0 + ry; b < b1; b += 8) {
      const size_t i = (size_t)b * C + c;
      const float sg = sigmoidf_(se[(size_t)b * 2 * C + c]);
      const float dm = dmean[i];
      s1 += fmaf(sg, s_du[i], dm);
      s2 += fmaf(sg, s_duz[i], dm * bsum2[i]);  // bsum2 holds the board MEAN of z2
    }
  sh[0][ry][cx] = s1; sh[1][ry][cx] = s2;
  __syncthreads();
  if (ry == 0 && c < C) {
    double a = 0.0, d = 0.0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i][cx]; d += sh[1][i][cx]; }
    atomicAdd(&sums[c], a);
    atomicAdd(&sums[C + c], d);
  }
}

// sums: [2][C] (sum dzh, sum dzh*z) over the batch that defines the BatchNorm statistics (all ranks under
// SyncBatchNorm); sums_local (optional): this rank's share, used for dgamma / dbeta (data-parallel gradient
// averaging sums the ranks' shares, exactly like torch.nn.SyncBatchNorm's backward).
__global__ void bn_bwd_finalize_kernel(double* sums, double* sums_local, double count, const float* w, const float* mean,
                                       const float* invstd, float* k1, float* k2, float* k3, float* dgamma, float* dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double S1 = sums[c], S2 = sums[C + c];
  sums[c] = 0.0; sums[C + c] = 0.0;
  double L1 = S1, L2 = S2;
  if (sums_local) { L1 = sums_local[c]; L2 = sums_local[C + c]; }
  const double mu = mean[c], is = invstd[c];
  const double dg = is * (S2 - mu * S1);
  const double kk1 = (double)w[c] * is;
  if (dgamma) dgamma[c] = (float)(is * (L2 - mu * L1));
  if (dbeta) dbeta[c] = (float)L1;
  k1[c] = (float)kk1;
  k2[c] = (float)(kk1 * is * dg / count);
  k3[c] = (float)(kk1 * (S1 / count - mu * is * dg / count));
}

template <typename T>
__global__ void block_bwd_dz2_kernel(PassBArgs g) {
  const int b = blockIdx.x, c = threadIdx.x, C = g.C;
  if (c >= C) return;
  const size_t base = (size_t)b * 81 * C + c;
  const T* dxp = (const T*)g.dxp + base; const T* xp = (const T*)g.xp + base; const T* z2 = (const T*)g.z2 + base;
  T* dz2 = (T*)g.dz2 + base;
  const float sg = sigmoidf_(g.se[(size_t)b * 2 * C + c]);
  const float dm = g.dse_in[(size_t)b * C + c] * (1.f / 81.f);
  const float k1 = g.k1[c], k2 = g.k2[c], k3 = g.k3[c];
#pragma unroll 9
  for (int p = 0; p < 81; ++p) {
    const float du = (g.xp == nullptr || kb_to_float<T>(xp[(size_t)p * C]) > 0.f) ? kb_to_float<T>(dxp[(size_t)p * C]) : 0.f;
    const float dzh = fmaf(du, sg, dm);
    dz2[(size_t)p * C] = kb_from_float<T>(k1 * dzh - k2 * kb_to_float<T>(z2[(size_t)p * C]) - k3);
  }
}

template <typename T>
__global__ void bn_bwd_apply_kernel(T* __restrict__ d, const T* __restrict__ z, const float* __restrict__ k1,
                                    const float* __restrict__ k2, const float* __restrict__ k3, long long n, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    d[i] = kb_from_float<T>(k1[c] * kb_to_float<T>(d[i]) - k2[c] * kb_to_float<T>(z[i]) - k3[c]);
  }
}

template <typename T>
__global__ void block_bwd_dx_kernel(PassDArgs g) {
  const int b = blockIdx.x, c = threadIdx.x, C = g.C;
  if (c >= C) return;
  const size_t base = (size_t)b * 81 * C + c;
  const T* x = (const T*)g.x + base;
  const T* dxc = g.dxc ? (const T*)g.dxc + base : nullptr;
  const T* dxp = g.dxp ? (const T*)g.dxp + base : nullptr;
  const T* xp = g.xp ? (const T*)g.xp + base : nullptr;
  T* dx = (T*)g.dx + base;
  float gmean = 0.f, gmax = 0.f, gstd = 0.f, mean = 0.f, mx = 0.f, hs = 0.f, hsz = 0.f;
  if (g.dpool) {
    const float* pr = g.pool + (size_t)b * 3 * C;
    const float* dp = g.dpool + (size_t)b * 3 * C;
    mean = pr[c]; mx = pr[C + c];
    const float sd = pr[2 * C + c];
    gmean = dp[c] * (1.f / 81.f);
    gstd = sd > 0.f ? dp[2 * C + c] / (81.f * sd) : 0.f;  // torch: d std/dx = 0 where std == 0
    int ties = 0;
#pragma unroll 9
    for (int p = 0; p < 81; ++p) ties += (kb_to_float<T>(x[(size_t)p * C]) == mx);
    gmax = ties > 0 ? dp[C + c] / (float)ties : 0.f;  // amax backward splits evenly across ties
  }
#pragma unroll 9
  for (int p = 0; p < 81; ++p) {
    float v = dxc ? kb_to_float<T>(dxc[(size_t)p * C]) : 0.f;
    if (dxp) {
      const float up = kb_to_float<T>(dxp[(size_t)p * C]);
      v += (xp == nullptr || kb_to_float<T>(xp[(size_t)p * C]) > 0.f) ? up : 0.f;
    }
    if (g.dpool) {
      const float xv = kb_to_float<T>(x[(size_t)p * C]);
      v += gmean + (xv == mx ? gmax : 0.f) + gstd * (xv - mean);
    }
    if (g.mask_out && !(kb_to_float<T>(x[(size_t)p * C]) > 0.f)) v = 0.f;
    const T st = kb_from_float<T>(v);
    dx[(size_t)p * C] = st;
    if (g.z_next) {
      const float r = kb_to_float<T>(st);
      hs += r;
      hsz = fmaf(r, kb_to_float<T>(((const T*)g.z_next)[base + (size_t)p * C]), hsz);
    }
  }
  if (g.z_next) { g.s_du[(size_t)b * C + c] = hs; g.s_duz[(size_t)b * C + c] = hsz; }
}

template <typename T>
__global__ void relu_bwd_stats_kernel(const T* dy, const T* __restrict__ y, const T* __restrict__ z,
                                      T* dzh, int C, double* sums) {  // dzh may alias dy
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const size_t base = (size_t)b * 81 * C + c;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 9
  for (int p = 0; p < 81; ++p) {
    const size_t i = base + (size_t)p * C;
    const float d = kb_to_float<T>(y[i]) > 0.f ? kb_to_float<T>(dy[i]) : 0.f;
    const T st = kb_from_float<T>(d);
    dzh[i] = st;
    const float r = kb_to_float<T>(st);
    s1 += r;
    s2 = fmaf(r, kb_to_float<T>(z[i]), s2);
  }
  atomicAdd(&sums[c], (double)s1);
  atomicAdd(&sums[C + c], (double)s2);
}

__global__ void relu_bwd_stats_f32_kernel(float* __restrict__ d, const float* __restrict__ act, const float* __restrict__ z,
                                          long long M, int C, long long rows_per_block, double* sums) {
  __shared__ float sh[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) {
      const long long i = r * C + c;
      const float v = act[i] > 0.f ? d[i] : 0.f;
      d[i] = v;
      s1 += v;
      s2 = fmaf(v, z[i], s2);
    }
  sh[0][ry][cx] = s1; sh[1][ry][cx] = s2;
  __syncthreads();
  if (ry == 0 && c < C) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i][cx]; q += sh[1][i][cx]; }
    atomicAdd(&sums[c], a);
    atomicAdd(&sums[C + c], q);
  }
}

// =================================================================================================
// Vectorised variants: thread = 8 consecutive channels (one 16-byte bf16 / two 16-byte fp32 accesses)
// x one of NPL pixel lanes; CTA = one board, 256 threads. C/8 channel groups * NPL lanes = 256, so a
// warp reads whole 128..512-byte rows; per-(board, channel) reductions finish through shared memory.
// Used when C % 8 == 0 and 256 % (C/8) == 0 (C = 16, 32, 64, 128, 256, 512 ...).
// =================================================================================================
constexpr int kVW = 4;  // channels per thread in the vectorised kernels (4 -> half the register state of 8)
template <typename T, int W> struct VV;
template <> struct VV<float, 8> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p) : "memory");
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4) : "memory");
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void round(float (&)[8]) {}
};
template <> struct VV<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    // volatile: ptxas otherwise sinks the loads of an unrolled group next to their uses (one pair in flight instead
    // of the whole group — seen in the ncu source view as a full-latency stall per pixel)
    float4 a;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p) : "memory");
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void round(float (&)[4]) {}
};
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> struct VV<bf16, 8> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf162(v[0], v[1]), pack_bf162(v[2], v[3]), pack_bf162(v[4], v[5]), pack_bf162(v[6], v[7]));
  }
  static __device__ __forceinline__ void round(float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  }
};
template <> struct VV<bf16, 4> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    uint2 u;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p) : "memory");
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf162(v[0], v[1]), pack_bf162(v[2], v[3]));
  }
  static __device__ __forceinline__ void round(float (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  }
};
template <typename T> using V8 = VV<T, kVW>;
// small fp32 vectors that are re-read by many threads (BatchNorm coefficients, per-board SE / bias values): the cached
// read-only path (L1-allocating), unlike the streaming activation loads above
__device__ __forceinline__ void ldf8(const float* p, float (&v)[kVW]) {
  static_assert(kVW == 4, "ldf8 assumes 4 channels per thread");
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// SE = squeeze-excite scale/shift + residual present; POOL = emit global-pool statistics of the output.
// Each CTA walks kApplyBoardsPerCta boards (fewer, longer CTAs: less scheduling / tail overhead per board).
template <typename T, bool SE, bool POOL>
__global__ void __launch_bounds__(256) apply_vec_kernel(ApplyArgs g, int kApplyBoardsPerCta) {
  __shared__ float red[POOL ? 5 : 1][POOL ? 256 * kVW : 1];  // [quantity][lane * C + c]
  const int C = g.C, C8 = C / kVW, NPL = 256 / C8;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  float a_[kVW], b_[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { a_[i] = 1.f; b_[i] = 0.f; }
  if (g.a) { ldf8(g.a + c0, a_); ldf8(g.b + c0, b_); }
  const int b_end = min(g.B, (int)(blockIdx.x + 1) * kApplyBoardsPerCta);
  for (int b = blockIdx.x * kApplyBoardsPerCta; b < b_end; ++b) {
    float sg[kVW], sf[kVW], gb[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) gb[i] = 0.f;
    if (SE && g.se != nullptr) {  // SE = residual present; the squeeze-excite scale/shift itself is optional (plain ResNet)
      ldf8(g.se + (size_t)b * 2 * C + c0, sg); ldf8(g.se + (size_t)b * 2 * C + C + c0, sf);
#pragma unroll
      for (int i = 0; i < kVW; ++i) {
        sg[i] = sigmoidf_(sg[i]);
        // fold the BN affine into the SE scale/shift: (z*a+b)*sg+sf = z*(a*sg) + (b*sg+sf)
        sf[i] = fmaf(b_[i], sg[i], sf[i]); sg[i] *= a_[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < kVW; ++i) { sg[i] = a_[i]; sf[i] = b_[i]; }
    }
    if (g.gbias) ldf8(g.gbias + (size_t)b * C + c0, gb);
    const size_t base = (size_t)b * 81 * C + c0;
    float s[kVW], mx[kVW], k0[kVW], ds[kVW], dss[kVW], tie[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) { s[i] = 0.f; mx[i] = -INFINITY; k0[i] = 0.f; ds[i] = 0.f; dss[i] = 0.f; tie[i] = 0.f; }
    int cnt = 0;
    auto finish = [&](float (&v)[kVW], const float (&r)[kVW], int p) {
#pragma unroll
      for (int i = 0; i < kVW; ++i) {
        float t = fmaf(v[i], sg[i], sf[i]);
        if (SE) t += r[i];
        v[i] = fmaxf(t, 0.f) + gb[i];
      }
      V8<T>::store((T*)g.out + base + (size_t)p * C, v);
      if (POOL) {
        V8<T>::round(v);
#pragma unroll
        for (int i = 0; i < kVW; ++i) {
          if (cnt == 0) k0[i] = v[i];
          const float d = v[i] - k0[i];
          s[i] += v[i]; ds[i] += d; dss[i] = fmaf(d, d, dss[i]);
          // running (max, number of elements equal to it): amax backward splits evenly across ties
          tie[i] = v[i] > mx[i] ? 1.f : (v[i] == mx[i] ? tie[i] + 1.f : tie[i]);
          mx[i] = fmaxf(mx[i], v[i]);
        }
        ++cnt;
      }
    };
    for (int p = pl; p < 81; p += 2 * NPL) {  // two pixels per iteration: 2-4 independent vector loads in flight
      const bool two = p + NPL < 81;
      float v0[kVW], r0[kVW], v1[kVW], r1[kVW];
      V8<T>::load((const T*)g.z + base + (size_t)p * C, v0);
      if (SE) V8<T>::load((const T*)g.res + base + (size_t)p * C, r0);
      if (two) {
        V8<T>::load((const T*)g.z + base + (size_t)(p + NPL) * C, v1);
        if (SE) V8<T>::load((const T*)g.res + base + (size_t)(p + NPL) * C, r1);
      }
      finish(v0, r0, p);
      if (two) finish(v1, r1, p + NPL);
    }
    if (!POOL) continue;
    // per-lane (count, mean, M2) -> Chan's parallel merge across the NPL pixel lanes
    const float fc = (float)cnt;
    __syncthreads();  // previous board's readers are done with `red`
#pragma unroll
    for (int i = 0; i < kVW; ++i) {
      const int o = pl * C + c0 + i;
      red[0][o] = s[i];
      red[POOL ? 1 : 0][o] = mx[i];
      red[POOL ? 2 : 0][o] = cnt > 0 ? k0[i] + ds[i] / fc : 0.f;            // lane mean
      red[POOL ? 3 : 0][o] = cnt > 0 ? dss[i] - ds[i] * ds[i] / fc : 0.f;   // lane M2
      red[POOL ? 4 : 0][o] = tie[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float S = 0.f, M = -INFINITY;
      for (int l = 0; l < NPL; ++l) { S += red[0][l * C + c]; M = fmaxf(M, red[POOL ? 1 : 0][l * C + c]); }
      const float mean = S * (1.f / 81.f);
      float m2 = 0.f;
      for (int l = 0; l < NPL; ++l) {
        const int n_l = l < 81 ? (81 - l + NPL - 1) / NPL : 0;  // pixels lane l visited
        const float dm = red[POOL ? 2 : 0][l * C + c] - mean;
        m2 += red[POOL ? 3 : 0][l * C + c] + (float)n_l * dm * dm;
      }
      const float sd = sqrtf(fmaxf(m2 * (1.f / 81.f), 0.f));
      float* pr = g.pool + (size_t)b * 3 * C;
      pr[c] = mean; pr[C + c] = M; pr[2 * C + c] = sd;
      if (g.pool_bf) {
        bf16* pb = (bf16*)g.pool_bf + (size_t)b * 3 * C;
        pb[c] = __float2bfloat16_rn(mean); pb[C + c] = __float2bfloat16_rn(M); pb[2 * C + c] = __float2bfloat16_rn(sd);
      }
      if (g.ties) {
        float t = 0.f;
        for (int l = 0; l < NPL; ++l) if (red[POOL ? 1 : 0][l * C + c] == M) t += red[POOL ? 4 : 0][l * C + c];
        g.ties[(size_t)b * C + c] = t;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) block_bwd_reduce_vec_kernel(BlockBwdArgs g) {
  __shared__ float red[2][256 * kVW];
  const int C = g.C, C8 = C / kVW, NPL = 256 / C8;
  const int b = blockIdx.x, cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  const size_t base = (size_t)b * 81 * C + c0;
  float s[kVW], sz[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { s[i] = 0.f; sz[i] = 0.f; }
  for (int p = pl; p < 81; p += NPL) {
    float d[kVW], x[kVW], z[kVW];
    V8<T>::load((const T*)g.dxp + base + (size_t)p * C, d);
    V8<T>::load((const T*)g.xp + base + (size_t)p * C, x);
    V8<T>::load((const T*)g.z2 + base + (size_t)p * C, z);
#pragma unroll
    for (int i = 0; i < kVW; ++i) { const float du = x[i] > 0.f ? d[i] : 0.f; s[i] += du; sz[i] = fmaf(du, z[i], sz[i]); }
  }
#pragma unroll
  for (int i = 0; i < kVW; ++i) { red[0][pl * C + c0 + i] = s[i]; red[1][pl * C + c0 + i] = sz[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < NPL; ++l) { a += red[0][l * C + c]; q += red[1][l * C + c]; }
    g.s_du[(size_t)b * C + c] = a;
    g.s_duz[(size_t)b * C + c] = q;
  }
}

template <typename T, bool XP>  // XP: dxp is the raw gradient and xp (the block output) supplies the ReLU mask
__global__ void __launch_bounds__(256) block_bwd_dz2_vec_kernel(PassBArgs g) {
  const int C = g.C, C8 = C / kVW, NPL = 256 / C8;
  const int b = blockIdx.x, cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  const size_t base = (size_t)b * 81 * C + c0;
  float sg[kVW], dm[kVW], k1[kVW], k2[kVW], k3[kVW];
  ldf8(g.se + (size_t)b * 2 * C + c0, sg); ldf8(g.dse_in + (size_t)b * C + c0, dm);
  ldf8(g.k1 + c0, k1); ldf8(g.k2 + c0, k2); ldf8(g.k3 + c0, k3);
#pragma unroll
  for (int i = 0; i < kVW; ++i) { sg[i] = sigmoidf_(sg[i]); dm[i] *= (1.f / 81.f); }
  constexpr int PIX = 4;  // pixels in flight per thread: 8 independent vector loads (2 streams) before the first use
  for (int p = pl; p < 81; p += PIX * NPL) {
    float d[PIX][kVW], x[PIX][kVW], z[PIX][kVW];
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int pj = p + j * NPL;
#pragma unroll
      for (int i = 0; i < kVW; ++i) x[j][i] = 0.f;
      if (pj < 81) {
        V8<T>::load((const T*)g.dxp + base + (size_t)pj * C, d[j]);
        if (XP) V8<T>::load((const T*)g.xp + base + (size_t)pj * C, x[j]);  // else: dxp is already masked (du)
        V8<T>::load((const T*)g.z2 + base + (size_t)pj * C, z[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int pj = p + j * NPL;
      if (pj < 81) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) {
          const float du = (!XP || x[j][i] > 0.f) ? d[j][i] : 0.f;
          d[j][i] = k1[i] * fmaf(du, sg[i], dm[i]) - k2[i] * z[j][i] - k3[i];
        }
        V8<T>::store((T*)g.dz2 + base + (size_t)pj * C, d[j]);
      }
    }
  }
}

// HOT: the in-tower configuration (conv data gradient + pre-masked residual gradient + pool backward with forward tie
// counts + masked output + next block's board sums) with every pointer test folded at compile time.
template <typename T, bool HOT>
__global__ void __launch_bounds__(256, 3) block_bwd_dx_vec_kernel(PassDArgs g) {
  const bool has_dxc = HOT || g.dxc != nullptr, has_dxp = HOT || g.dxp != nullptr, has_xp = !HOT && g.xp != nullptr;
  const bool has_dpool = HOT || g.dpool != nullptr, has_ties = HOT || g.ties != nullptr, has_zn = HOT || g.z_next != nullptr;
  const bool mask_out = HOT || g.mask_out != 0;
  __shared__ float red[256 * kVW];
  __shared__ float red2[256 * kVW];
  const int C = g.C, C8 = C / kVW, NPL = 256 / C8;
  const int b = blockIdx.x, cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  const size_t base = (size_t)b * 81 * C + c0;
  float gmean[kVW], gmax[kVW], gstd[kVW], mean[kVW], mx[kVW], hs[kVW], hsz[kVW];
  const bool need_x = has_dpool || mask_out;
#pragma unroll
  for (int i = 0; i < kVW; ++i) { gmean[i] = 0.f; gmax[i] = 0.f; gstd[i] = 0.f; mean[i] = 0.f; mx[i] = 0.f; hs[i] = 0.f; hsz[i] = 0.f; }
  if (has_dpool) {
    const float* pr = g.pool + (size_t)b * 3 * C;
    const float* dp = g.dpool + (size_t)b * 3 * C;
    float sd[kVW], dmean[kVW], dmaxv[kVW], dstd[kVW], ties[kVW];
    ldf8(pr + c0, mean); ldf8(pr + C + c0, mx); ldf8(pr + 2 * C + c0, sd);
    ldf8(dp + c0, dmean); ldf8(dp + C + c0, dmaxv); ldf8(dp + 2 * C + c0, dstd);
    if (has_ties) {
      ldf8(g.ties + (size_t)b * C + c0, ties);  // counted by the forward apply kernel
    } else {
      float tl[kVW];
#pragma unroll
      for (int i = 0; i < kVW; ++i) tl[i] = 0.f;
      for (int p = pl; p < 81; p += NPL) {
        float x[kVW];
        V8<T>::load((const T*)g.x + base + (size_t)p * C, x);
#pragma unroll
        for (int i = 0; i < kVW; ++i) tl[i] += (x[i] == mx[i]) ? 1.f : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kVW; ++i) red[pl * C + c0 + i] = tl[i];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kVW; ++i) {
        float t = 0.f;
        for (int l = 0; l < NPL; ++l) t += red[l * C + c0 + i];
        ties[i] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < kVW; ++i) {
      gmean[i] = dmean[i] * (1.f / 81.f);
      gstd[i] = sd[i] > 0.f ? dstd[i] / (81.f * sd[i]) : 0.f;   // torch: d std/dx = 0 where std == 0
      gmax[i] = ties[i] > 0.f ? dmaxv[i] / ties[i] : 0.f;       // amax backward splits evenly across ties
    }
  }
  auto one = [&](int p, float (&v)[kVW], const float (&t)[kVW], const float (&y)[kVW], const float (&xx)[kVW],
                 const float (&zn)[kVW]) {
    if (has_dxp) {
#pragma unroll
      for (int i = 0; i < kVW; ++i) v[i] += (!has_xp || y[i] > 0.f) ? t[i] : 0.f;
    }
    if (has_dpool) {
#pragma unroll
      for (int i = 0; i < kVW; ++i) v[i] += gmean[i] + (xx[i] == mx[i] ? gmax[i] : 0.f) + gstd[i] * (xx[i] - mean[i]);
    }
    if (mask_out) {
#pragma unroll
      for (int i = 0; i < kVW; ++i) v[i] = xx[i] > 0.f ? v[i] : 0.f;
    }
    V8<T>::store((T*)g.dx + base + (size_t)p * C, v);
    if (has_zn) {
      V8<T>::round(v);
#pragma unroll
      for (int i = 0; i < kVW; ++i) { hs[i] += v[i]; hsz[i] = fmaf(v[i], zn[i], hsz[i]); }
    }
  };
  for (int p = pl; p < 81; p += 2 * NPL) {  // two pixels per iteration: up to 8 independent vector loads in flight
    const bool two = p + NPL < 81;
    const int p1 = p + NPL;
    float v0[kVW], t0[kVW], y0[kVW], x0[kVW], n0[kVW], v1[kVW], t1[kVW], y1[kVW], x1[kVW], n1[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) {
      v0[i] = 0.f; v1[i] = 0.f; t0[i] = 0.f; t1[i] = 0.f; y0[i] = 0.f; y1[i] = 0.f; x0[i] = 0.f; x1[i] = 0.f; n0[i] = 0.f; n1[i] = 0.f;
    }
    if (has_dxc) V8<T>::load((const T*)g.dxc + base + (size_t)p * C, v0);
    if (has_dxp) V8<T>::load((const T*)g.dxp + base + (size_t)p * C, t0);
    if (has_xp) V8<T>::load((const T*)g.xp + base + (size_t)p * C, y0);
    if (need_x) V8<T>::load((const T*)g.x + base + (size_t)p * C, x0);
    if (has_zn) V8<T>::load((const T*)g.z_next + base + (size_t)p * C, n0);
    if (two) {
      if (has_dxc) V8<T>::load((const T*)g.dxc + base + (size_t)p1 * C, v1);
      if (has_dxp) V8<T>::load((const T*)g.dxp + base + (size_t)p1 * C, t1);
      if (has_xp) V8<T>::load((const T*)g.xp + base + (size_t)p1 * C, y1);
      if (need_x) V8<T>::load((const T*)g.x + base + (size_t)p1 * C, x1);
      if (has_zn) V8<T>::load((const T*)g.z_next + base + (size_t)p1 * C, n1);
    }
    one(p, v0, t0, y0, x0, n0);
    if (two) one(p1, v1, t1, y1, x1, n1);
  }
  if (has_zn) {  // per-(board, channel) sums of du and du * z_next across the pixel lanes
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kVW; ++i) { red[pl * C + c0 + i] = hs[i]; red2[pl * C + c0 + i] = hsz[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float a = 0.f, q = 0.f;
      for (int l = 0; l < NPL; ++l) { a += red[l * C + c]; q += red2[l * C + c]; }
      g.s_du[(size_t)b * C + c] = a;
      g.s_duz[(size_t)b * C + c] = q;
    }
  }
}

// dzh = dy * [mask > 0] with mask = y (ma == null) or y*ma[c] + mb[c]; per-channel sums of dzh and
// dzh*z (double atomics); optional per-(board, channel) sum of the UNMASKED dy (gpool-bias gradient).
// dzh may alias dy (in place). Two pixels per iteration keep 4-6 independent 16-byte loads in flight.
constexpr int kStatsBoardsPerCta = 4;  // boards per CTA: 4x fewer double atomics on the 2*C channel accumulators

template <typename T>
__global__ void __launch_bounds__(256, 4) relu_bwd_stats_vec_kernel(const T* dy, const T* __restrict__ y,
                                                                   const T* __restrict__ z, T* dzh, int B, int C,
                                                                   const float* __restrict__ ma, const float* __restrict__ mb,
                                                                   float* __restrict__ board_sum, double* sums) {
  __shared__ float red[2][256 * kVW];
  const int C8 = C / kVW, NPL = 256 / C8;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  const bool same = (y == z);
  float s1[kVW], s2[kVW], fa[kVW], fb[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { s1[i] = 0.f; s2[i] = 0.f; fa[i] = 1.f; fb[i] = 0.f; }
  if (ma) { ldf8(ma + c0, fa); ldf8(mb + c0, fb); }
  const int b_end = min(B, (int)(blockIdx.x + 1) * kStatsBoardsPerCta);
  for (int b = blockIdx.x * kStatsBoardsPerCta; b < b_end; ++b) {
    const size_t base = (size_t)b * 81 * C + c0;
    float s0[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) s0[i] = 0.f;
    for (int p = pl; p < 81; p += 2 * NPL) {
      const bool two = p + NPL < 81;
      float d0[kVW], a0[kVW], z0[kVW], d1[kVW], a1[kVW], z1[kVW];
      V8<T>::load(dy + base + (size_t)p * C, d0);
      V8<T>::load(y + base + (size_t)p * C, a0);
      if (!same) V8<T>::load(z + base + (size_t)p * C, z0);
      if (two) {
        V8<T>::load(dy + base + (size_t)(p + NPL) * C, d1);
        V8<T>::load(y + base + (size_t)(p + NPL) * C, a1);
        if (!same) V8<T>::load(z + base + (size_t)(p + NPL) * C, z1);
      }
#pragma unroll
      for (int i = 0; i < kVW; ++i) {
        s0[i] += d0[i];
        if (same) z0[i] = a0[i];
        d0[i] = fmaf(a0[i], fa[i], fb[i]) > 0.f ? d0[i] : 0.f;
      }
      V8<T>::store(dzh + base + (size_t)p * C, d0);
      V8<T>::round(d0);
#pragma unroll
      for (int i = 0; i < kVW; ++i) { s1[i] += d0[i]; s2[i] = fmaf(d0[i], z0[i], s2[i]); }
      if (two) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) {
          s0[i] += d1[i];
          if (same) z1[i] = a1[i];
          d1[i] = fmaf(a1[i], fa[i], fb[i]) > 0.f ? d1[i] : 0.f;
        }
        V8<T>::store(dzh + base + (size_t)(p + NPL) * C, d1);
        V8<T>::round(d1);
#pragma unroll
        for (int i = 0; i < kVW; ++i) { s1[i] += d1[i]; s2[i] = fmaf(d1[i], z1[i], s2[i]); }
      }
    }
    if (board_sum) {  // per-(board, channel) sum of the unmasked gradient
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kVW; ++i) red[0][pl * C + c0 + i] = s0[i];
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += 256) {
        float u = 0.f;
        for (int l = 0; l < NPL; ++l) u += red[0][l * C + c];
        board_sum[(size_t)b * C + c] = u;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kVW; ++i) { red[0][pl * C + c0 + i] = s1[i]; red[1][pl * C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < NPL; ++l) { a += red[0][l * C + c]; q += red[1][l * C + c]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)q);
  }
}

// The block variant of relu_bwd_stats_vec_kernel (mask = z*ma + mb > 0, dy rewritten in place): two streams only, so
// FOUR pixels per thread are kept in flight (8 independent vector loads) to cover the HBM latency.
template <typename T>
__global__ void __launch_bounds__(256, 3) mask_bwd_stats_vec_kernel(T* d, const T* __restrict__ z, int B, int C,
                                                                   const float* __restrict__ ma, const float* __restrict__ mb,
                                                                   float* __restrict__ board_sum, double* sums) {
  __shared__ float red[2][256 * kVW];
  constexpr int PIX = 4;
  const int C8 = C / kVW, NPL = 256 / C8;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, c0 = cg * kVW;
  float s1[kVW], s2[kVW], fa[kVW], fb[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  ldf8(ma + c0, fa); ldf8(mb + c0, fb);
  const int b_end = min(B, (int)(blockIdx.x + 1) * kStatsBoardsPerCta);
  for (int b = blockIdx.x * kStatsBoardsPerCta; b < b_end; ++b) {
    const size_t base = (size_t)b * 81 * C + c0;
    float s0[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) s0[i] = 0.f;
    for (int p = pl; p < 81; p += PIX * NPL) {
      float dd[PIX][kVW], zz[PIX][kVW];
#pragma unroll
      for (int j = 0; j < PIX; ++j) {
        const int pj = p + j * NPL;
        if (pj < 81) {
          V8<T>::load(d + base + (size_t)pj * C, dd[j]);
          V8<T>::load(z + base + (size_t)pj * C, zz[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < PIX; ++j) {
        const int pj = p + j * NPL;
        if (pj < 81) {
#pragma unroll
          for (int i = 0; i < kVW; ++i) {
            s0[i] += dd[j][i];
            dd[j][i] = fmaf(zz[j][i], fa[i], fb[i]) > 0.f ? dd[j][i] : 0.f;
          }
          V8<T>::store(d + base + (size_t)pj * C, dd[j]);
          V8<T>::round(dd[j]);
#pragma unroll
          for (int i = 0; i < kVW; ++i) { s1[i] += dd[j][i]; s2[i] = fmaf(dd[j][i], zz[j][i], s2[i]); }
        }
      }
    }
    if (board_sum) {  // per-(board, channel) sum of the unmasked gradient
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kVW; ++i) red[0][pl * C + c0 + i] = s0[i];
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += 256) {
        float u = 0.f;
        for (int l = 0; l < NPL; ++l) u += red[0][l * C + c];
        board_sum[(size_t)b * C + c] = u;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kVW; ++i) { red[0][pl * C + c0 + i] = s1[i]; red[1][pl * C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < NPL; ++l) { a += red[0][l * C + c]; q += red[1][l * C + c]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)q);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_vec_kernel(T* __restrict__ d, const T* __restrict__ z,
                                                                 const float* __restrict__ k1, const float* __restrict__ k2,
                                                                 const float* __restrict__ k3, long long n8, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)((i * kVW) % C);
    float v[kVW], zz[kVW], a[kVW], b[kVW], e[kVW];
    V8<T>::load(d + i * kVW, v); V8<T>::load(z + i * kVW, zz);
    ldf8(k1 + c0, a); ldf8(k2 + c0, b); ldf8(k3 + c0, e);
#pragma unroll
    for (int j = 0; j < kVW; ++j) v[j] = a[j] * v[j] - b[j] * zz[j] - e[j];
    V8<T>::store(d + i * kVW, v);
  }
}

// =================================================================================================
// Column kernels: one thread owns kVW channels of ONE board for all 81 pixels (C/kVW threads per board,
// 256/(C/kVW) boards per CTA pass). Per-(board, channel) reductions are thread-local — no shared memory, no
// barriers in the streaming loop — and the per-board preamble is amortised over 81 pixels. 81 = 27 x 3: three
// pixels (x 2-5 streams) of independent vector loads are in flight per thread, with no tail iteration.
// =================================================================================================
constexpr int kColPix = 3;
inline bool col_ok(int C) { return C % kVW == 0 && C / kVW <= 256 && 256 % (C / kVW) == 0; }
// Small launches (at most two CTAs per SM: up to ~1200 boards at 256 channels — the reference's 256-sample minibatches, the
// per-GPU share of an 8-GPU update) are latency-bound: a thread walks its 81 pixels in 27 dependent round trips of 3
// pixels (21-26 us whatever the batch). They run the 9-pixel instantiation: 9 round trips, one CTA per SM's worth of
// registers, which is all such a launch can use anyway.
inline bool col_small(int B, int C) { return kb_ceil_div(B, 256 / (C / kVW)) <= 2 * 148; }
inline int col_grid(int B, int C) {
  const int bpc = 256 / (C / kVW);
  const int want = kb_ceil_div(B, bpc);
  return want < 148 * 8 ? want : 148 * 8;
}

__device__ __forceinline__ void stf4(float* p, const float (&v)[kVW]) {
  static_assert(kVW == 4, "stf4 assumes 4 channels per thread");
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// pass B: dz2 = k1 * (du * sigmoid(scale) + dse_in / 81) - k2 * z2 - k3     (du already masked by the block's ReLU)
template <typename T, int kColPix>
__global__ void __launch_bounds__(256, kColPix == 3 ? 3 : 1) block_bwd_dz2_col_kernel(PassBArgs g) {
  const int C = g.C, TPB = C / kVW, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * kVW;
  float k1[kVW], k2[kVW], k3[kVW];
  ldf8(g.k1 + c0, k1); ldf8(g.k2 + c0, k2); ldf8(g.k3 + c0, k3);
  for (int b = blockIdx.x * BPC + slot; b < g.B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
    float sg[kVW], dm[kVW];
    ldf8(g.se + (size_t)b * 2 * C + c0, sg); ldf8(g.dse_in + (size_t)b * C + c0, dm);
#pragma unroll
    for (int i = 0; i < kVW; ++i) { sg[i] = k1[i] * sigmoidf_(sg[i]); dm[i] = k1[i] * dm[i] * (1.f / 81.f) - k3[i]; }
#pragma unroll 1
    for (int p = 0; p < 81; p += kColPix) {
      float d[kColPix][kVW], z[kColPix][kVW];
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
        V8<T>::load((const T*)g.dxp + base + (size_t)(p + j) * C, d[j]);
        V8<T>::load((const T*)g.z2 + base + (size_t)(p + j) * C, z[j]);
      }
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) d[j][i] = fmaf(d[j][i], sg[i], dm[i]) - k2[i] * z[j][i];
        V8<T>::store((T*)g.dz2 + base + (size_t)(p + j) * C, d[j]);
      }
    }
  }
}

// pass D, in-tower configuration (see PassDArgs): dx = [x > 0] * (dxc + du' + pool backward), plus the board sums
// of dx and dx * z_next for the producing block.
template <typename T, bool ZN, int kColPix>
__global__ void __launch_bounds__(256, kColPix == 3 ? 3 : 1) block_bwd_dx_col_kernel(PassDArgs g) {
  const int C = g.C, TPB = C / kVW, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * kVW;
  for (int b = blockIdx.x * BPC + slot; b < g.B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
    float gmean[kVW], gmax[kVW], gstd[kVW], mean[kVW], mx[kVW], hs[kVW], hsz[kVW];
    {
      const float* pr = g.pool + (size_t)b * 3 * C;
      const float* dp = g.dpool + (size_t)b * 3 * C;
      float sd[kVW], dmean[kVW], dmaxv[kVW], dstd[kVW], ties[kVW];
      ldf8(pr + c0, mean); ldf8(pr + C + c0, mx); ldf8(pr + 2 * C + c0, sd);
      ldf8(dp + c0, dmean); ldf8(dp + C + c0, dmaxv); ldf8(dp + 2 * C + c0, dstd);
      ldf8(g.ties + (size_t)b * C + c0, ties);
#pragma unroll
      for (int i = 0; i < kVW; ++i) {
        gstd[i] = sd[i] > 0.f ? dstd[i] / (81.f * sd[i]) : 0.f;   // torch: d std/dx = 0 where std == 0
        gmax[i] = ties[i] > 0.f ? dmaxv[i] / ties[i] : 0.f;       // amax backward splits evenly across ties
        gmean[i] = dmean[i] * (1.f / 81.f);
        hs[i] = 0.f; hsz[i] = 0.f;
      }
    }
#pragma unroll 1
    for (int p = 0; p < 81; p += kColPix) {
      float v[kColPix][kVW], t[kColPix][kVW], x[kColPix][kVW], zn[kColPix][kVW];
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
        const size_t o = base + (size_t)(p + j) * C;
        V8<T>::load((const T*)g.dxc + o, v[j]);
        V8<T>::load((const T*)g.dxp + o, t[j]);
        V8<T>::load((const T*)g.x + o, x[j]);
        if (ZN) V8<T>::load((const T*)g.z_next + o, zn[j]);
      }
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) {
          const float xv = x[j][i];
          float r = v[j][i] + t[j][i];
          r += gmean[i] + (xv == mx[i] ? gmax[i] : 0.f) + gstd[i] * (xv - mean[i]);
          v[j][i] = xv > 0.f ? r : 0.f;
        }
        V8<T>::store((T*)g.dx + base + (size_t)(p + j) * C, v[j]);
        if (ZN) {
          V8<T>::round(v[j]);
#pragma unroll
          for (int i = 0; i < kVW; ++i) { hs[i] += v[j][i]; hsz[i] = fmaf(v[j][i], zn[j][i], hsz[i]); }
        }
      }
    }
    if (ZN) { stf4(g.s_du + (size_t)b * C + c0, hs); stf4(g.s_duz + (size_t)b * C + c0, hsz); }
  }
}

// mask_bwd_stats, column form: d <- d * [z*ma + mb > 0] in place; board_sum[b][c] = sum_p d (unmasked);
// sums += per-channel sums of the masked gradient and of masked gradient * z (double atomics, once per CTA).
// kStore = false: the masked gradient is not written back (two tensor passes instead of three); the consumer,
// bn_bwd_apply_flat_kernel<T, true>, recomputes the same mask from z.
template <typename T, int kColPix, bool kStore = true>
__global__ void __launch_bounds__(256, kColPix == 3 ? 3 : 1) mask_bwd_stats_col_kernel(T* d, const T* __restrict__ z, int B, int C,
                                                                 const float* __restrict__ ma, const float* __restrict__ mb,
                                                                 float* __restrict__ board_sum, double* sums) {
  __shared__ float red[2][256 * kVW];
  const int TPB = C / kVW, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * kVW;
  float s1[kVW], s2[kVW], fa[kVW], fb[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  ldf8(ma + c0, fa); ldf8(mb + c0, fb);
  for (int b = blockIdx.x * BPC + slot; b < B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
    float s0[kVW];
#pragma unroll
    for (int i = 0; i < kVW; ++i) s0[i] = 0.f;
#pragma unroll 1
    for (int p = 0; p < 81; p += kColPix) {
      float dd[kColPix][kVW], zz[kColPix][kVW];
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
        V8<T>::load(d + base + (size_t)(p + j) * C, dd[j]);
        V8<T>::load(z + base + (size_t)(p + j) * C, zz[j]);
      }
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) {
          s0[i] += dd[j][i];
          dd[j][i] = fmaf(zz[j][i], fa[i], fb[i]) > 0.f ? dd[j][i] : 0.f;
        }
        if (kStore) V8<T>::store(d + base + (size_t)(p + j) * C, dd[j]);
        V8<T>::round(dd[j]);
#pragma unroll
        for (int i = 0; i < kVW; ++i) { s1[i] += dd[j][i]; s2[i] = fmaf(dd[j][i], zz[j][i], s2[i]); }
      }
    }
    if (board_sum) stf4(board_sum + (size_t)b * C + c0, s0);
  }
#pragma unroll
  for (int i = 0; i < kVW; ++i) { red[0][slot * C + c0 + i] = s1[i]; red[1][slot * C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < BPC; ++l) { a += red[0][l * C + c]; q += red[1][l * C + c]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)q);
  }
}

// out = relu(z*a[c] + b[c]) + gbias[board][c], flat: the whole tensor is one contiguous stream (consecutive CTAs read
// consecutive 2 KB segments, the DRAM-friendliest pattern; bn_bwd_apply_vec_kernel has the same shape). The grid stride
// is a multiple of C, so a thread keeps the same channels and its BatchNorm coefficients live in registers; the
// per-board bias is a 16-byte L1/L2 hit per vector. Four vectors per thread are in flight.
template <typename T>
__global__ void __launch_bounds__(256, 4) apply_flat_kernel(ApplyArgs g, unsigned rows) {
  constexpr int U = 4;
  const int C = g.C, cpv = C / kVW;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned rs = (gridDim.x * blockDim.x) / cpv;   // pixel rows advanced per grid stride (exact: see launcher)
  const int c0 = (int)(tid % cpv) * kVW;
  float a_[kVW], b_[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { a_[i] = 1.f; b_[i] = 0.f; }
  if (g.a) { ldf8(g.a + c0, a_); ldf8(g.b + c0, b_); }
  for (unsigned row = tid / cpv; row < rows; row += U * rs) {
    float v[U][kVW], gb[U][kVW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // clamped, branch-free loads (a branch would pin each conversion next to its load and serialise the group)
      const unsigned ru = min(row + u * rs, rows - 1);
#pragma unroll
      for (int k = 0; k < kVW; ++k) gb[u][k] = 0.f;
      V8<T>::load((const T*)g.z + (size_t)ru * C + c0, v[u]);
      if (g.gbias) ldf8(g.gbias + (size_t)(ru / 81u) * C + c0, gb[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned ru = row + u * rs;
      if (ru < rows) {
#pragma unroll
        for (int k = 0; k < kVW; ++k) v[u][k] = fmaxf(fmaf(v[u][k], a_[k], b_[k]), 0.f) + gb[u][k];
        V8<T>::store((T*)g.out + (size_t)ru * C + c0, v[u]);
      }
    }
  }
}

// dz = k1[c]*dzh - k2[c]*z - k3[c] in place, flat (same shape as apply_flat_kernel: coefficients in registers, four
// branch-free vector pairs in flight per thread, consecutive CTAs on consecutive 2 KB segments).
// kMask: d holds the UNMASKED gradient of relu(z*ma + mb); the ReLU mask is recomputed here from z (same expression as
// mask_bwd_stats_col_kernel, which then does not have to write the masked tensor back).
template <typename T, bool kMask = false>
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_flat_kernel(T* d, const T* __restrict__ z, const float* __restrict__ k1,
                                                                  const float* __restrict__ k2, const float* __restrict__ k3,
                                                                  unsigned rows, int C, const float* __restrict__ ma = nullptr,
                                                                  const float* __restrict__ mb = nullptr) {
  constexpr int U = 4;
  const int cpv = C / kVW;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned rs = (gridDim.x * blockDim.x) / cpv;
  const int c0 = (int)(tid % cpv) * kVW;
  float a[kVW], b[kVW], e[kVW], fa[kVW], fb[kVW];
  ldf8(k1 + c0, a); ldf8(k2 + c0, b); ldf8(k3 + c0, e);
  if (kMask) { ldf8(ma + c0, fa); ldf8(mb + c0, fb); }
  for (unsigned row = tid / cpv; row < rows; row += U * rs) {
    float v[U][kVW], zz[U][kVW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned ru = min(row + u * rs, rows - 1);
      V8<T>::load(d + (size_t)ru * C + c0, v[u]);
      V8<T>::load(z + (size_t)ru * C + c0, zz[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned ru = row + u * rs;
      if (ru < rows) {
#pragma unroll
        for (int j = 0; j < kVW; ++j) {
          const float dv = (!kMask || fmaf(zz[u][j], fa[j], fb[j]) > 0.f) ? v[u][j] : 0.f;
          v[u][j] = a[j] * dv - b[j] * zz[u][j] - e[j];
        }
        V8<T>::store(d + (size_t)ru * C + c0, v[u]);
      }
    }
  }
}

inline int flat_grid(long long rows, int cpv) {
  long long ctas = (rows * cpv + 256 * 4 - 1) / (256 * 4);
  if (ctas > 148 * 8) ctas = 148 * 8;
  while (((ctas * 256) % cpv) != 0) ++ctas;  // thread total must be a multiple of the vectors per pixel row
  return (int)ctas;
}

// relu_bwd_stats, column form: dzh = dy * [y > 0] (dzh may alias dy); sums += per-channel sums of dzh and dzh * z.
// SAME: y and z are the same tensor (two input streams instead of three).
template <typename T, bool SAME>
__global__ void __launch_bounds__(256, 3) relu_bwd_stats_col_kernel(const T* dy, const T* __restrict__ y, const T* __restrict__ z,
                                                                   T* dzh, int B, int C, double* sums) {
  __shared__ float red[2][256 * kVW];
  const int TPB = C / kVW, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * kVW;
  float s1[kVW], s2[kVW];
#pragma unroll
  for (int i = 0; i < kVW; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for (int b = blockIdx.x * BPC + slot; b < B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
#pragma unroll 1
    for (int p = 0; p < 81; p += kColPix) {
      float dd[kColPix][kVW], yy[kColPix][kVW], zz[kColPix][kVW];
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
        V8<T>::load(dy + base + (size_t)(p + j) * C, dd[j]);
        V8<T>::load(y + base + (size_t)(p + j) * C, yy[j]);
        if (!SAME) V8<T>::load(z + base + (size_t)(p + j) * C, zz[j]);
      }
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) dd[j][i] = yy[j][i] > 0.f ? dd[j][i] : 0.f;
        V8<T>::store(dzh + base + (size_t)(p + j) * C, dd[j]);
        V8<T>::round(dd[j]);
#pragma unroll
        for (int i = 0; i < kVW; ++i) { s1[i] += dd[j][i]; s2[i] = fmaf(dd[j][i], SAME ? yy[j][i] : zz[j][i], s2[i]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kVW; ++i) { red[0][slot * C + c0 + i] = s1[i]; red[1][slot * C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < BPC; ++l) { a += red[0][l * C + c]; q += red[1][l * C + c]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)q);
  }
}

// pass D of a plain residual block, column form: dx = dxc + dxp * [xp > 0]  (no pool branch, unmasked output)
template <typename T>
__global__ void __launch_bounds__(256, 3) resblock_bwd_dx_col_kernel(PassDArgs g) {
  const int C = g.C, TPB = C / kVW, BPC = 256 / TPB;
  const int slot = threadIdx.x / TPB, c0 = (threadIdx.x % TPB) * kVW;
  for (int b = blockIdx.x * BPC + slot; b < g.B; b += gridDim.x * BPC) {
    const size_t base = (size_t)b * 81 * C + c0;
#pragma unroll 1
    for (int p = 0; p < 81; p += kColPix) {
      float v[kColPix][kVW], t[kColPix][kVW], y[kColPix][kVW];
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
        const size_t o = base + (size_t)(p + j) * C;
        V8<T>::load((const T*)g.dxc + o, v[j]);
        V8<T>::load((const T*)g.dxp + o, t[j]);
        V8<T>::load((const T*)g.xp + o, y[j]);
      }
#pragma unroll
      for (int j = 0; j < kColPix; ++j) {
#pragma unroll
        for (int i = 0; i < kVW; ++i) v[j][i] += y[j][i] > 0.f ? t[j][i] : 0.f;
        V8<T>::store((T*)g.dx + base + (size_t)(p + j) * C, v[j]);
      }
    }
  }
}

inline bool vec_ok(int C) { return C % kVW == 0 && C / kVW <= 256 && 256 % (C / kVW) == 0; }

inline int ch_threads(int C) { return ((C + 31) / 32) * 32; }

}  // namespace

#define KB_DISPATCH_T(dtype, KERNEL, GRID, BLOCK, SMEM, ST, ...)                                  \
  do {                                                                                            \
    if ((dtype) == KB_F32) KERNEL<float><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__);                 \
    else KERNEL<bf16><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__);                                    \
    KB_CUDA_LAUNCH_CHECK();                                                                       \
  } while (0)

int kbk_apply(const ApplyArgs& a, cudaStream_t st) {
  KB_CHECK_ARG(a.C >= 1 && a.C <= 1024, "apply: C=%d out of range", a.C);
  if (a.B == 0) return KB_OK;
  if (a.C % kVW == 0 && a.se == nullptr && a.res == nullptr && a.pool == nullptr && a.pool_bf == nullptr && a.ties == nullptr) {
    // plain BatchNorm + ReLU + bias: flat streaming kernel; grid * 256 threads is a multiple of C / kVW
    const long long rows = (long long)a.B * 81;
    const int cpv = a.C / kVW;                       // vectors per pixel row
    if (rows < (1LL << 31)) {
      KB_DISPATCH_T(a.dtype, apply_flat_kernel, flat_grid(rows, cpv), 256, 0, st, a, (unsigned)rows);
      return KB_OK;
    }
  }
  if (vec_ok(a.C) && (a.se == nullptr || a.res != nullptr)) {
    const bool se = a.res != nullptr, pool = a.pool != nullptr;
    static int bpc = 0;
    if (bpc == 0) { const char* e = getenv("KB_APPLY_BPC"); bpc = e ? atoi(e) : 1; if (bpc < 1) bpc = 1; }
#define KB_APPLY_VEC(SE_, POOL_)                                                                   \
    do {                                                                                           \
      if (a.dtype == KB_F32) apply_vec_kernel<float, SE_, POOL_><<<kb_ceil_div(a.B, bpc), 256, 0, st>>>(a, bpc); \
      else apply_vec_kernel<bf16, SE_, POOL_><<<kb_ceil_div(a.B, bpc), 256, 0, st>>>(a, bpc);               \
    } while (0)
    if (se && pool) KB_APPLY_VEC(true, true);
    else if (se) KB_APPLY_VEC(true, false);
    else if (pool) KB_APPLY_VEC(false, true);
    else KB_APPLY_VEC(false, false);
#undef KB_APPLY_VEC
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  KB_DISPATCH_T(a.dtype, apply_kernel, a.B, ch_threads(a.C), 0, st, a);
  return KB_OK;
}

int kbk_bn_eval_affine(const float* w, const float* bias, const float* rm, const float* rv, float eps, int C, float* a,
                       float* b, cudaStream_t st) {
  bn_eval_affine_kernel<<<kb_ceil_div(C, 128), 128, 0, st>>>(w, bias, rm, rv, eps, C, a, b);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_bn_finalize(double* sums, double count, const float* w, const float* bias, const float* running_mean,
                    const float* running_var, float* rm_out, float* rv_out, long long* nbt, float momentum, float eps,
                    int C, float* a, float* b, float* mean, float* invstd, cudaStream_t st) {
  bn_finalize_kernel<<<kb_ceil_div(C, 128), 128, 0, st>>>(sums, count, w, bias, running_mean, running_var, rm_out, rv_out,
                                                          nbt, momentum, eps, C, a, b, mean, invstd);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_affine_rows(const float* in, const float* a, const float* b, float* out, void* out_bf16, long long rows, int C,
                    cudaStream_t st) {
  const long long n = rows * C;
  if (n == 0) return KB_OK;
  affine_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, a, b, out, (bf16*)out_bf16, n, C);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_rows_stats(const float* x, long long M, int C, double* sums, cudaStream_t st) {
  if (M == 0) return KB_OK;
  const long long rpb = 512;
  rows_stats_kernel<<<dim3((unsigned)((M + rpb - 1) / rpb), kb_ceil_div(C, 32)), 256, 0, st>>>(x, M, C, rpb, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_block_bwd_reduce(const BlockBwdArgs& a, cudaStream_t st) {
  KB_CHECK_ARG(a.C <= 1024, "block_bwd_reduce: C too large");
  if (vec_ok(a.C)) KB_DISPATCH_T(a.dtype, block_bwd_reduce_vec_kernel, a.B, 256, 0, st, a);
  else KB_DISPATCH_T(a.dtype, block_bwd_reduce_kernel, a.B, ch_threads(a.C), 0, st, a);
  return KB_OK;
}

int kbk_se_bwd_prep(const float* s_du, const float* s_duz, const float* a2, const float* b2, const float* se, float* dse,
                    int B, int C, cudaStream_t st) {
  se_bwd_prep_kernel<<<kb_ceil_div((long long)B * C, 256), 256, 0, st>>>(s_du, s_duz, a2, b2, se, dse, B, C);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_bn2_bwd_sums(const float* s_du, const float* s_duz, const float* se, const float* dmean,
                     const float* bsum2, int B, int C, double* sums, cudaStream_t st) {
  const int bpb = 64;
  bn2_bwd_sums_kernel<<<dim3(kb_ceil_div(B, bpb), kb_ceil_div(C, 32)), 256, 0, st>>>(s_du, s_duz, se, dmean, bsum2, B, C,
                                                                                     bpb, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_bn_bwd_finalize(double* sums, double count, const float* w, const float* mean, const float* invstd, float* k1,
                        float* k2, float* k3, float* dgamma, float* dbeta, int C, cudaStream_t st, double* sums_local) {
  bn_bwd_finalize_kernel<<<kb_ceil_div(C, 128), 128, 0, st>>>(sums, sums_local, count, w, mean, invstd, k1, k2, k3, dgamma, dbeta, C);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_block_bwd_dz2(const PassBArgs& a, cudaStream_t st) {
  KB_CHECK_ARG(a.C <= 1024, "block_bwd_dz2: C too large");
  static int col = -1;
  if (col < 0) { const char* e = getenv("KB_COL_KERNELS"); col = (e && e[0] == '0') ? 0 : 1; }
  if (col && col_ok(a.C) && a.xp == nullptr) {
    const bool small = col_small(a.B, a.C);
    if (a.dtype == KB_F32) { block_bwd_dz2_col_kernel<float, 3><<<col_grid(a.B, a.C), 256, 0, st>>>(a); }
    else if (small) { block_bwd_dz2_col_kernel<bf16, 9><<<col_grid(a.B, a.C), 256, 0, st>>>(a); }
    else { block_bwd_dz2_col_kernel<bf16, 3><<<col_grid(a.B, a.C), 256, 0, st>>>(a); }
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (vec_ok(a.C)) {
    if (a.xp != nullptr) {
      if (a.dtype == KB_F32) block_bwd_dz2_vec_kernel<float, true><<<a.B, 256, 0, st>>>(a);
      else block_bwd_dz2_vec_kernel<bf16, true><<<a.B, 256, 0, st>>>(a);
    } else {
      if (a.dtype == KB_F32) block_bwd_dz2_vec_kernel<float, false><<<a.B, 256, 0, st>>>(a);
      else block_bwd_dz2_vec_kernel<bf16, false><<<a.B, 256, 0, st>>>(a);
    }
    KB_CUDA_LAUNCH_CHECK();
  }
  else KB_DISPATCH_T(a.dtype, block_bwd_dz2_kernel, a.B, ch_threads(a.C), 0, st, a);
  return KB_OK;
}

int kbk_bn_bwd_apply(void* d, const void* z, const float* k1, const float* k2, const float* k3, long long rows, int C,
                     int dtype, cudaStream_t st) {
  const long long n = rows * C;
  if (n == 0) return KB_OK;
  if (C % kVW == 0 && rows < (1LL << 31)) {
    const int grid = flat_grid(rows, C / kVW);
    if (dtype == KB_F32) { bn_bwd_apply_flat_kernel<float><<<grid, 256, 0, st>>>((float*)d, (const float*)z, k1, k2, k3, (unsigned)rows, C); }
    else { bn_bwd_apply_flat_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)d, (const bf16*)z, k1, k2, k3, (unsigned)rows, C); }
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (C % kVW == 0) {
    const long long n8 = n / kVW;
    const int grid8 = (int)min((long long)148 * 16, (n8 + 255) / 256);
    if (dtype == KB_F32) bn_bwd_apply_vec_kernel<float><<<grid8, 256, 0, st>>>((float*)d, (const float*)z, k1, k2, k3, n8, C);
    else bn_bwd_apply_vec_kernel<bf16><<<grid8, 256, 0, st>>>((bf16*)d, (const bf16*)z, k1, k2, k3, n8, C);
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  const int grid = (int)min((long long)148 * 16, (n + 255) / 256);
  if (dtype == KB_F32) bn_bwd_apply_kernel<float><<<grid, 256, 0, st>>>((float*)d, (const float*)z, k1, k2, k3, n, C);
  else bn_bwd_apply_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)d, (const bf16*)z, k1, k2, k3, n, C);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

// dz = k1 * (d * [z*ma + mb > 0]) - k2*z - k3 in place: BatchNorm backward of relu(bn(z)) with the ReLU mask recomputed from z
int kbk_bn_bwd_apply_masked_supported(long long rows, int C) { return col_ok(C) && rows < (1LL << 31) ? 1 : 0; }

int kbk_bn_bwd_apply_masked(void* d, const void* z, const float* k1, const float* k2, const float* k3, const float* ma,
                            const float* mb, long long rows, int C, int dtype, cudaStream_t st) {
  KB_CHECK_ARG(kbk_bn_bwd_apply_masked_supported(rows, C), "bn_bwd_apply_masked: unsupported shape");
  KB_CHECK_ARG(ma != nullptr && mb != nullptr, "bn_bwd_apply_masked: mask coefficients missing");
  if (rows == 0) return KB_OK;
  const int grid = flat_grid(rows, C / kVW);
  if (dtype == KB_F32) { bn_bwd_apply_flat_kernel<float, true><<<grid, 256, 0, st>>>((float*)d, (const float*)z, k1, k2, k3, (unsigned)rows, C, ma, mb); }
  else { bn_bwd_apply_flat_kernel<bf16, true><<<grid, 256, 0, st>>>((bf16*)d, (const bf16*)z, k1, k2, k3, (unsigned)rows, C, ma, mb); }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_block_bwd_dx(const PassDArgs& a, cudaStream_t st) {
  KB_CHECK_ARG(a.C <= 1024, "block_bwd_dx: C too large");
  static int col = -1;
  if (col < 0) { const char* e = getenv("KB_COL_KERNELS"); col = (e && e[0] == '0') ? 0 : 1; }
  if (col && col_ok(a.C) && a.dxc && a.dxp && !a.xp && a.dpool && a.ties && a.mask_out) {
    const int grid = col_grid(a.B, a.C);
    const bool small = col_small(a.B, a.C);
    if (a.z_next) {
      if (a.dtype == KB_F32) { block_bwd_dx_col_kernel<float, true, 3><<<grid, 256, 0, st>>>(a); }
      else if (small) { block_bwd_dx_col_kernel<bf16, true, 9><<<grid, 256, 0, st>>>(a); }
      else { block_bwd_dx_col_kernel<bf16, true, 3><<<grid, 256, 0, st>>>(a); }
    } else {
      if (a.dtype == KB_F32) { block_bwd_dx_col_kernel<float, false, 3><<<grid, 256, 0, st>>>(a); }
      else if (small) { block_bwd_dx_col_kernel<bf16, false, 9><<<grid, 256, 0, st>>>(a); }
      else { block_bwd_dx_col_kernel<bf16, false, 3><<<grid, 256, 0, st>>>(a); }
    }
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (col && col_ok(a.C) && a.dxc && a.dxp && a.xp && !a.dpool && !a.z_next && !a.mask_out) {
    KB_DISPATCH_T(a.dtype, resblock_bwd_dx_col_kernel, col_grid(a.B, a.C), 256, 0, st, a);
    return KB_OK;
  }
  if (vec_ok(a.C)) {
    const bool hot = a.dxc && a.dxp && !a.xp && a.dpool && a.ties && a.z_next && a.mask_out;
    if (hot) {
      if (a.dtype == KB_F32) block_bwd_dx_vec_kernel<float, true><<<a.B, 256, 0, st>>>(a);
      else block_bwd_dx_vec_kernel<bf16, true><<<a.B, 256, 0, st>>>(a);
    } else {
      if (a.dtype == KB_F32) block_bwd_dx_vec_kernel<float, false><<<a.B, 256, 0, st>>>(a);
      else block_bwd_dx_vec_kernel<bf16, false><<<a.B, 256, 0, st>>>(a);
    }
    KB_CUDA_LAUNCH_CHECK();
  }
  else KB_DISPATCH_T(a.dtype, block_bwd_dx_kernel, a.B, ch_threads(a.C), 0, st, a);
  return KB_OK;
}

int kbk_relu_bwd_stats(const void* dy, const void* y, const void* z, void* dzh, long long rows, int C, int dtype,
                       double* sums, cudaStream_t st) {
  KB_CHECK_ARG(rows % 81 == 0 && C <= 1024, "relu_bwd_stats: bad shape");
  const int B = (int)(rows / 81);
  static int col = -1;
  if (col < 0) { const char* e = getenv("KB_COL_KERNELS"); col = (e && e[0] == '0') ? 0 : 1; }
  if (col && col_ok(C)) {
    const int grid = col_grid(B, C);
    const bool same = (y == z);
    if (dtype == KB_F32) {
      if (same) relu_bwd_stats_col_kernel<float, true><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, (const float*)z, (float*)dzh, B, C, sums);
      else relu_bwd_stats_col_kernel<float, false><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, (const float*)z, (float*)dzh, B, C, sums);
    } else {
      if (same) relu_bwd_stats_col_kernel<bf16, true><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)z, (bf16*)dzh, B, C, sums);
      else relu_bwd_stats_col_kernel<bf16, false><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)z, (bf16*)dzh, B, C, sums);
    }
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (vec_ok(C)) {
    if (dtype == KB_F32)
      relu_bwd_stats_vec_kernel<float><<<kb_ceil_div(B, kStatsBoardsPerCta), 256, 0, st>>>((const float*)dy, (const float*)y, (const float*)z, (float*)dzh, B, C, nullptr, nullptr, nullptr, sums);
    else
      relu_bwd_stats_vec_kernel<bf16><<<kb_ceil_div(B, kStatsBoardsPerCta), 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)z, (bf16*)dzh, B, C, nullptr, nullptr, nullptr, sums);
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (dtype == KB_F32)
    relu_bwd_stats_kernel<float><<<B, ch_threads(C), 0, st>>>((const float*)dy, (const float*)y, (const float*)z, (float*)dzh, C, sums);
  else
    relu_bwd_stats_kernel<bf16><<<B, ch_threads(C), 0, st>>>((const bf16*)dy, (const bf16*)y, (const bf16*)z, (bf16*)dzh, C, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_mask_bwd_stats_supported(int C) { return vec_ok(C) ? 1 : 0; }

int kbk_mask_bwd_stats(void* d_inout, const void* z, const float* ma, const float* mb, float* board_sum, int B, int C,
                       int dtype, double* sums, cudaStream_t st) {
  KB_CHECK_ARG(vec_ok(C), "mask_bwd_stats: unsupported channel count %d", C);
  if (B == 0) return KB_OK;
  KB_CHECK_ARG(ma != nullptr && mb != nullptr, "mask_bwd_stats: mask coefficients missing");
  static int col = -1;
  if (col < 0) { const char* e = getenv("KB_COL_KERNELS"); col = (e && e[0] == '0') ? 0 : 1; }
  if (col && col_ok(C)) {
    const int grid = col_grid(B, C);
    if (dtype == KB_F32) { mask_bwd_stats_col_kernel<float, 3><<<grid, 256, 0, st>>>((float*)d_inout, (const float*)z, B, C, ma, mb, board_sum, sums); }
    else if (col_small(B, C)) { mask_bwd_stats_col_kernel<bf16, 9><<<grid, 256, 0, st>>>((bf16*)d_inout, (const bf16*)z, B, C, ma, mb, board_sum, sums); }
    else { mask_bwd_stats_col_kernel<bf16, 3><<<grid, 256, 0, st>>>((bf16*)d_inout, (const bf16*)z, B, C, ma, mb, board_sum, sums); }
    KB_CUDA_LAUNCH_CHECK();
    return KB_OK;
  }
  if (dtype == KB_F32)
    mask_bwd_stats_vec_kernel<float><<<kb_ceil_div(B, kStatsBoardsPerCta), 256, 0, st>>>((float*)d_inout, (const float*)z, B, C, ma, mb, board_sum, sums);
  else
    mask_bwd_stats_vec_kernel<bf16><<<kb_ceil_div(B, kStatsBoardsPerCta), 256, 0, st>>>((bf16*)d_inout, (const bf16*)z, B, C, ma, mb, board_sum, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

// Statistics only: `d` is left unmasked (see kbk_bn_bwd_apply_masked). Column kernels only.
int kbk_mask_bwd_stats_ro_supported(int C) {
  static int col = -1;
  if (col < 0) { const char* e = getenv("KB_COL_KERNELS"); col = (e && e[0] == '0') ? 0 : 1; }
  return col && col_ok(C) ? 1 : 0;
}

int kbk_mask_bwd_stats_ro(const void* d, const void* z, const float* ma, const float* mb, float* board_sum, int B, int C,
                          int dtype, double* sums, cudaStream_t st) {
  KB_CHECK_ARG(kbk_mask_bwd_stats_ro_supported(C), "mask_bwd_stats_ro: unsupported channel count %d", C);
  if (B == 0) return KB_OK;
  KB_CHECK_ARG(ma != nullptr && mb != nullptr, "mask_bwd_stats_ro: mask coefficients missing");
  const int grid = col_grid(B, C);
  if (dtype == KB_F32) { mask_bwd_stats_col_kernel<float, 3, false><<<grid, 256, 0, st>>>((float*)d, (const float*)z, B, C, ma, mb, board_sum, sums); }
  else if (col_small(B, C)) { mask_bwd_stats_col_kernel<bf16, 9, false><<<grid, 256, 0, st>>>((bf16*)d, (const bf16*)z, B, C, ma, mb, board_sum, sums); }
  else { mask_bwd_stats_col_kernel<bf16, 3, false><<<grid, 256, 0, st>>>((bf16*)d, (const bf16*)z, B, C, ma, mb, board_sum, sums); }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_relu_bwd_stats_f32(float* d, const float* act, const float* z, long long M, int C, double* sums, cudaStream_t st) {
  if (M == 0) return KB_OK;
  const long long rpb = 512;
  relu_bwd_stats_f32_kernel<<<dim3((unsigned)((M + rpb - 1) / rpb), kb_ceil_div(C, 32)), 256, 0, st>>>(d, act, z, M, C, rpb, sums);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_fill_zero(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return KB_OK;
  KB_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return KB_OK;
}
