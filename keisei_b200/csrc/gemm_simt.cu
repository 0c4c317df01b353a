// gemm_simt.cu — small dense layers of the model on CUDA cores, fp32 accumulate.
//
// Covers the Linear / 1x1-conv layers around the trunk (reference se_resnet.py:57-61 global_fc,
// :63-66 SE, :119-130 heads) forward and backward: skinny GEMMs (M = boards or pixels) whose FLOPs
// are ~0.2 % of the model. One register-tiled kernel (BMx64x16, 256 threads, 4 values of each operand
// per thread per K step fetched as ONE 16-byte / 8-byte vector along the contiguous dimension,
// register double buffering) with optional operand transposes, a per-k affine(+ReLU) prologue on A,
// bias/ReLU/mask epilogue, board-pitched rows (the padded (B, 11264) policy buffer) and split-K
// with fp32 atomics (weight gradients reduce over the batch).
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

constexpr int BN = 64, BK = 16;

// 4 consecutive elements starting at element offset `off`
__device__ __forceinline__ float4 ld4(const void* p, int dtype, long long off, bool vec_ok, int valid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid <= 0) return v;
  if (dtype == KB_F32) {
    const float* f = (const float*)p + off;
    if (vec_ok && valid >= 4) return __ldg(reinterpret_cast<const float4*>(f));
    v.x = f[0];
    if (valid > 1) v.y = f[1];
    if (valid > 2) v.z = f[2];
    if (valid > 3) v.w = f[3];
  } else {
    const bf16* h = (const bf16*)p + off;
    if (vec_ok && valid >= 4) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(h));
      v.x = __uint_as_float(u.x << 16); v.y = __uint_as_float(u.x & 0xffff0000u);
      v.z = __uint_as_float(u.y << 16); v.w = __uint_as_float(u.y & 0xffff0000u);
      return v;
    }
    v.x = __bfloat162float(h[0]);
    if (valid > 1) v.y = __bfloat162float(h[1]);
    if (valid > 2) v.z = __bfloat162float(h[2]);
    if (valid > 3) v.w = __bfloat162float(h[3]);
  }
  return v;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// TF32 = true (BM = 64 only): the same tiles and operand staging, but the inner product runs on the tensor cores as
// warp-level m16n8k8 TF32 MMAs with fp32 accumulation (operands rounded to TF32 when they are staged). Used for the
// bf16 / AMP path only, where the reference itself computes these Linear layers in bf16 under autocast (TF32 keeps 3
// more mantissa bits); the fp32 path keeps exact fp32 FMAs. These GEMMs are far too small for a tcgen05 pipeline to
// pay off (K or N of 32..768) but large enough (M = 8192 boards) to be CUDA-core bound as SIMT.
// (bx, by, bz) = (N tile, M tile, K slice) of this CTA: blockIdx for a plain launch, decoded from a linear tile index in
// a grouped launch
template <int BM, bool TF32>
__device__ __forceinline__ void gemm_body(const GemmArgs& g, int k_per_slice, int a_vec, int b_vec, int bx, int by, int bz) {
  constexpr int TM = BM / 16;  // rows per thread (4 or 2)
  constexpr int PAD = TF32 ? 8 : 4;  // +8: the (k = lane%4, m = lane/4) fragment reads hit 32 distinct banks
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = by * BM, n0 = bx * BN;
  const int k_begin = bz * k_per_slice;
  const int k_end = min(g.K, k_begin + k_per_slice);
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // per-thread operand coordinates: 4 contiguous elements per K step
  // A: !transA -> (m = tid/4, k4 = (tid%4)*4);  transA -> (k = tid/(BM/4), m4 = (tid%(BM/4))*4)
  const bool a_active = g.transA ? (tid < BK * (BM / 4)) : (tid < BM * 4);
  const int a_r = g.transA ? tid / (BM / 4) : tid / 4;          // transA: k row ; else m row
  const int a_c = g.transA ? (tid % (BM / 4)) * 4 : (tid % 4) * 4;  // transA: m offset ; else k offset
  // B: transB (B[n][k]) -> (n = tid/4, k4 = (tid%4)*4);  !transB (B[k][n]) -> (k = tid/16, n4 = (tid%16)*4)
  const int b_r = g.transB ? tid / 4 : tid / 16;
  const int b_c = g.transB ? (tid % 4) * 4 : (tid % 16) * 4;

  auto a_row_off = [&](long long row) -> long long {
    if (g.a_group_rows > 0) return (row / g.a_group_rows) * g.a_group_pitch + (row % g.a_group_rows) * g.lda;
    return row * g.lda;
  };
  auto load_a = [&](int k0) -> float4 {
    if (!a_active) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.transA) {
      const int gk = k0 + a_r, gm = m0 + a_c;
      if (gk >= k_end) return make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v = ld4(g.A, g.a_dtype, a_row_off(gk) + gm, a_vec != 0, g.M - gm);
      if (g.a_pa) { const float pa = g.a_pa[gk], pb = g.a_pb[gk]; v.x = fmaf(v.x, pa, pb); v.y = fmaf(v.y, pa, pb); v.z = fmaf(v.z, pa, pb); v.w = fmaf(v.w, pa, pb); }
      if (g.a_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      return v;
    }
    const int gm = m0 + a_r, gk = k0 + a_c;
    if (gm >= g.M) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = ld4(g.A, g.a_dtype, a_row_off(gm) + gk, a_vec != 0, k_end - gk);
    if (g.a_pa) {
      const int kk = min(gk, g.K - 4 > 0 ? g.K - 1 : 0);
      (void)kk;
      if (gk < k_end) v.x = fmaf(v.x, g.a_pa[gk], g.a_pb[gk]);
      if (gk + 1 < k_end) v.y = fmaf(v.y, g.a_pa[gk + 1], g.a_pb[gk + 1]);
      if (gk + 2 < k_end) v.z = fmaf(v.z, g.a_pa[gk + 2], g.a_pb[gk + 2]);
      if (gk + 3 < k_end) v.w = fmaf(v.w, g.a_pa[gk + 3], g.a_pb[gk + 3]);
    }
    if (g.a_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
  };
  auto load_b = [&](int k0) -> float4 {
    if (g.transB) {
      const int gn = n0 + b_r, gk = k0 + b_c;
      if (gn >= g.N) return make_float4(0.f, 0.f, 0.f, 0.f);
      return ld4(g.B, g.b_dtype, (long long)gn * g.ldb + gk, b_vec != 0, k_end - gk);
    }
    const int gk = k0 + b_r, gn = n0 + b_c;
    if (gk >= k_end) return make_float4(0.f, 0.f, 0.f, 0.f);
    return ld4(g.B, g.b_dtype, (long long)gk * g.ldb + gn, b_vec != 0, g.N - gn);
  };
  auto store_a = [&](float4 v) {
    if (!a_active) return;
    if (TF32) { v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w); }
    if (g.transA) *reinterpret_cast<float4*>(&As[a_r][a_c]) = v;
    else { As[a_c][a_r] = v.x; As[a_c + 1][a_r] = v.y; As[a_c + 2][a_r] = v.z; As[a_c + 3][a_r] = v.w; }
  };
  auto store_b = [&](float4 v) {
    if (TF32) { v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w); }
    if (g.transB) { Bs[b_c][b_r] = v.x; Bs[b_c + 1][b_r] = v.y; Bs[b_c + 2][b_r] = v.z; Bs[b_c + 3][b_r] = v.w; }
    else *reinterpret_cast<float4*>(&Bs[b_r][b_c]) = v;
  };

  // one K step of the inner product on the staged tiles
  auto compute_step = [&]() {
    if constexpr (TF32) {
      // warp (wm, wn) of a 4 x 2 grid owns rows 16*wm..+15 and columns 32*wn..+31 (four m16n8 tiles)
      const int lane = tid & 31, wid = tid >> 5, wm = wid & 3, wn = wid >> 2, gq = lane >> 2, tq = lane & 3;
#pragma unroll
      for (int ks = 0; ks < BK; ks += 8) {
        uint32_t af[4];
        af[0] = __float_as_uint(As[ks + tq][wm * 16 + gq]);
        af[1] = __float_as_uint(As[ks + tq][wm * 16 + gq + 8]);
        af[2] = __float_as_uint(As[ks + tq + 4][wm * 16 + gq]);
        af[3] = __float_as_uint(As[ks + tq + 4][wm * 16 + gq + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t bfr[2];
          bfr[0] = __float_as_uint(Bs[ks + tq][wn * 32 + nt * 8 + gq]);
          bfr[1] = __float_as_uint(Bs[ks + tq + 4][wn * 32 + nt * 8 + gq]);
          mma_tf32_16x8x8(acc[nt], af, bfr);
        }
      }
    } else {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM];
      if (TM == 4) { const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]); av[0] = a.x; av[1] = a.y; av[TM - 2] = a.z; av[TM - 1] = a.w; }
      else { const float2 a = *reinterpret_cast<const float2*>(&As[k][ty * 2]); av[0] = a.x; av[1] = a.y; }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    }
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // Lean main loop for the common case — fp32 operands with aligned 4-element groups, no A prologue, no board-pitched
  // rows, an interior tile and a whole number of K steps: running pointers and unconditional 16-byte loads. (ncu on the
  // global-pool-MLP backward group at 8192 samples, generic loop: 37 warp instructions per MMA — IMAD 23 %, ISETP 13 %,
  // LDC 10 %, BRA 9 %, HMMA 2.7 % — issue-bound at 111 us for 3.2 GFLOP; the bounds / layout logic per load was the cost.)
  const bool lean = (a_vec & 2) != 0 && b_vec != 0 && g.a_dtype == KB_F32 && g.b_dtype == KB_F32 && g.a_group_rows == 0 &&
                    g.a_pa == nullptr && !g.a_relu && m0 + BM <= g.M && n0 + BN <= g.N && k_end > k_begin &&
                    ((k_end - k_begin) % BK) == 0;
  if (lean) {
    const long long lda = g.lda, ldb = g.ldb;
    const float* pa = (const float*)g.A + (g.transA ? (long long)(k_begin + a_r) * lda + m0 + a_c : (long long)(m0 + a_r) * lda + k_begin + a_c);
    const float* pb = (const float*)g.B + (g.transB ? (long long)(n0 + b_r) * ldb + k_begin + b_c : (long long)(k_begin + b_r) * ldb + n0 + b_c);
    const long long a_step = g.transA ? (long long)BK * lda : BK, b_step = g.transB ? BK : (long long)BK * ldb;
    const int steps = (k_end - k_begin) / BK;
    auto lda4 = [&]() { const float4 v = a_active ? __ldg(reinterpret_cast<const float4*>(pa)) : zero4; pa += a_step; return v; };
    auto ldb4 = [&]() { const float4 v = __ldg(reinterpret_cast<const float4*>(pb)); pb += b_step; return v; };
    float4 ra = lda4(), rb = ldb4();
    float4 ra1 = zero4, rb1 = zero4;
    if (steps > 1) { ra1 = lda4(); rb1 = ldb4(); }
    for (int sidx = 0; sidx < steps; ++sidx) {
      store_a(ra); store_b(rb);
      __syncthreads();
      ra = ra1; rb = rb1;
      if (sidx + 2 < steps) { ra1 = lda4(); rb1 = ldb4(); }
      compute_step();
      __syncthreads();
    }
  } else {
  // register prefetch TWO K steps ahead: these problems are small (a few tiles, 8-64 K steps), so a CTA's time is the
  // chain of its K steps and each step exposes one global-load latency divided by the prefetch depth
  float4 ra = load_a(k_begin), rb = load_b(k_begin);
  float4 ra1 = k_begin + BK < k_end ? load_a(k_begin + BK) : zero4, rb1 = k_begin + BK < k_end ? load_b(k_begin + BK) : zero4;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    store_a(ra); store_b(rb);
    __syncthreads();
    ra = ra1; rb = rb1;
    if (k0 + 2 * BK < k_end) { ra1 = load_a(k0 + 2 * BK); rb1 = load_b(k0 + 2 * BK); }
    compute_step();
    __syncthreads();
  }
  }
  // ---- epilogue ----
  // interior tile with plain fp32 rows: no bounds / layout decisions per element
  const bool lean_out = m0 + BM <= g.M && n0 + BN <= g.N && g.c_group_rows == 0 && g.c_dtype == KB_F32;
  float* const Cf = (float*)g.C;
  const long long ldc = g.ldc;
  const float* const bias = (g.bias && bz == 0) ? g.bias : nullptr;
  const bool atomic_out = g.splitk > 1;
  const bool relu_out = g.relu != 0;
  const float* const mask_src = g.mask_src;
  const long long ld_mask = g.ld_mask;
  auto emit_lean = [&](int gm, int gn, float v) {
    if (bias) v += bias[gn];
    float* dst = Cf + (long long)gm * ldc + gn;
    if (atomic_out) { atomicAdd(dst, v); return; }
    if (relu_out) v = fmaxf(v, 0.f);
    if (mask_src && !(mask_src[(long long)gm * ld_mask + gn] > 0.f)) v = 0.f;
    *dst = v;
  };
  auto emit = [&](int gm, int gn, float v) {
    if (gm >= g.M || gn >= g.N) return;
    long long crow;
    if (g.c_group_rows > 0) crow = ((long long)gm / g.c_group_rows) * g.c_group_pitch + ((long long)gm % g.c_group_rows) * g.ldc;
    else crow = (long long)gm * g.ldc;
    if (g.bias && bz == 0) v += g.bias[gn];
    if (g.splitk > 1) {
      atomicAdd(((float*)g.C) + crow + gn, v);
    } else {
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask_src && !(g.mask_src[(long long)gm * g.ld_mask + gn] > 0.f)) v = 0.f;
      if (g.c_dtype == KB_F32) ((float*)g.C)[crow + gn] = v;
      else ((bf16*)g.C)[crow + gn] = __float2bfloat16_rn(v);
    }
  };
  if constexpr (TF32) {
    // m16n8 accumulator: c0/c1 -> row lane/4, cols 2*(lane%4) + {0,1}; c2/c3 -> row lane/4 + 8
    const int lane = tid & 31, wid = tid >> 5, wm = wid & 3, wn = wid >> 2, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int gn = n0 + wn * 32 + nt * 8 + 2 * tq, gm = m0 + wm * 16 + gq;
      if (lean_out) {
        emit_lean(gm, gn, acc[nt][0]); emit_lean(gm, gn + 1, acc[nt][1]);
        emit_lean(gm + 8, gn, acc[nt][2]); emit_lean(gm + 8, gn + 1, acc[nt][3]);
      } else {
        emit(gm, gn, acc[nt][0]); emit(gm, gn + 1, acc[nt][1]);
        emit(gm + 8, gn, acc[nt][2]); emit(gm + 8, gn + 1, acc[nt][3]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (lean_out) emit_lean(m0 + ty * TM + i, n0 + tx * 4 + j, acc[i][j]);
        else emit(m0 + ty * TM + i, n0 + tx * 4 + j, acc[i][j]);
      }
  }
}

template <int BM, bool TF32>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g, int k_per_slice, int a_vec, int b_vec) {
  gemm_body<BM, TF32>(g, k_per_slice, a_vec, b_vec, blockIdx.x, blockIdx.y, blockIdx.z);
}

// out[n] += sum_m X[m][n]; block (32, 8)
struct ColsumArgs { const void* X; int dtype; long long ldx; int group_rows; long long group_pitch; int M, N, rows_per_block; float* out; };

__device__ __forceinline__ void colsum_body(const void* __restrict__ X, int dtype, long long ldx, int group_rows, long long group_pitch,
                                            int M, int N, int rows_per_block, float* out, int bx, int by) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int n = by * 32 + cx;
  const int r0 = bx * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + ry; r < r1; r += 8) {
      long long off;
      if (group_rows > 0) off = ((long long)r / group_rows) * group_pitch + ((long long)r % group_rows) * ldx + n;
      else off = (long long)r * ldx + n;
      s += dtype == KB_F32 ? ((const float*)X)[off] : __bfloat162float(((const bf16*)X)[off]);
    }
  sh[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += sh[i][cx];
    atomicAdd(&out[n], a);
  }
}

__global__ void colsum_kernel(const void* __restrict__ X, int dtype, long long ldx, int group_rows, long long group_pitch,
                              int M, int N, int rows_per_block, float* out) {
  colsum_body(X, dtype, ldx, group_rows, group_pitch, M, N, rows_per_block, out, blockIdx.x, blockIdx.y);
}

// Several independent small GEMMs and column sums as ONE launch (grouped): the Linear-layer backward of a block is four
// latency-bound GEMMs of a few dozen tiles each plus two bias-gradient column sums — issued one after the other they cost
// a launch latency apiece and never fill the GPU; grouped, the independent ones run side by side. Tiles are numbered
// linearly over the problems; every CTA decodes its problem and its (N tile, M tile, K slice).
constexpr int kMaxGroupGemms = 4, kMaxGroupColsums = 4;
struct GemmGroup {
  GemmArgs g[kMaxGroupGemms];
  int kps[kMaxGroupGemms], a_vec[kMaxGroupGemms], b_vec[kMaxGroupGemms], nx[kMaxGroupGemms], ny[kMaxGroupGemms];
  int first[kMaxGroupGemms + 1];                 // first linear tile of each GEMM; first[n] = start of the column sums
  int n;
  ColsumArgs cs[kMaxGroupColsums];
  int cs_first[kMaxGroupColsums + 1], cs_nx[kMaxGroupColsums];
  int ncs;
};

template <bool TF32>
__global__ void __launch_bounds__(256) gemm_group_kernel(const __grid_constant__ GemmGroup G) {
  const int t = blockIdx.x;
  if (t < G.first[G.n]) {
    int p = 0;
    while (p + 1 < G.n && t >= G.first[p + 1]) ++p;
    const int local = t - G.first[p];
    const int bx = local % G.nx[p], by = (local / G.nx[p]) % G.ny[p], bz = local / (G.nx[p] * G.ny[p]);
    gemm_body<64, TF32>(G.g[p], G.kps[p], G.a_vec[p], G.b_vec[p], bx, by, bz);
    return;
  }
  const int u = t - G.first[G.n];
  int q = 0;
  while (q + 1 < G.ncs && u >= G.cs_first[q + 1]) ++q;
  const int local = u - G.cs_first[q];
  const ColsumArgs& c = G.cs[q];
  colsum_body(c.X, c.dtype, c.ldx, c.group_rows, c.group_pitch, c.M, c.N, c.rows_per_block, c.out, local % G.cs_nx[q], local / G.cs_nx[q]);
}

// deferral: while a group is open on this host thread, kbk_gemm / kbk_colsum append to it instead of launching
struct PendingGroup { GemmGroup grp; bool open; int tf32; };
thread_local PendingGroup g_pending = {{}, false, 0};

// can every 4-element group of this operand be fetched with one aligned vector load?
bool vec4_ok(const void* p, int dtype, long long ld, int group_rows, long long group_pitch) {
  const int esz = dtype == KB_F32 ? 4 : 2;
  if (((uintptr_t)p) % (4 * esz) != 0) return false;
  if (ld % 4 != 0) return false;
  if (group_rows > 0 && group_pitch % 4 != 0) return false;
  return true;
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
  // vector loads need the 4-element groups to be aligned AND in-bounds handling via `valid`; the
  // K offset of every slice is a multiple of 16, so alignment reduces to base/ld divisibility
  // bit 1 of a_vec: the lean main loop may be used (KB_GEMM_LEAN=0 switches it off for A/B measurements)
  static const int lean_bit = [] { const char* e = getenv("KB_GEMM_LEAN"); return (e && e[0] == '0') ? 0 : 2; }();
  const int a_vec = vec4_ok(g.A, g.a_dtype, g.lda, g.a_group_rows, g.a_group_pitch) ? (1 | lean_bit) : 0;
  const int b_vec = vec4_ok(g.B, g.b_dtype, g.ldb, 0, 0) ? 1 : 0;
  if (g_pending.open && g_pending.grp.n < kMaxGroupGemms && (g_pending.grp.n + g_pending.grp.ncs == 0 || g_pending.tf32 == (g.tf32 ? 1 : 0))) {
    GemmGroup& G = g_pending.grp;
    const int p = G.n++;
    g_pending.tf32 = g.tf32 ? 1 : 0;
    G.g[p] = a; G.kps[p] = kps; G.a_vec[p] = a_vec; G.b_vec[p] = b_vec;
    G.nx[p] = kb_ceil_div(g.N, BN); G.ny[p] = kb_ceil_div(g.M, 64);
    G.first[p + 1] = G.first[p] + G.nx[p] * G.ny[p] * splitk;
    return KB_OK;
  }
  const long long tiles64 = (long long)kb_ceil_div(g.N, BN) * kb_ceil_div(g.M, 64) * splitk;
  if (g.tf32) {
    gemm_kernel<64, true><<<dim3(kb_ceil_div(g.N, BN), kb_ceil_div(g.M, 64), splitk), 256, 0, st>>>(a, kps, a_vec, b_vec);
  } else if (tiles64 >= 2 * 148) {
    gemm_kernel<64, false><<<dim3(kb_ceil_div(g.N, BN), kb_ceil_div(g.M, 64), splitk), 256, 0, st>>>(a, kps, a_vec, b_vec);
  } else {
    gemm_kernel<32, false><<<dim3(kb_ceil_div(g.N, BN), kb_ceil_div(g.M, 32), splitk), 256, 0, st>>>(a, kps, a_vec, b_vec);
  }
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_colsum(const void* X, int dtype, long long ldx, int group_rows, long long group_pitch, int M, int N, float* out,
               cudaStream_t st) {
  if (M == 0) return KB_OK;
  const int rpb = 1024;
  if (g_pending.open && g_pending.grp.ncs < kMaxGroupColsums) {
    GemmGroup& G = g_pending.grp;
    const int q = G.ncs++;
    G.cs[q] = ColsumArgs{X, dtype, ldx, group_rows, group_pitch, M, N, rpb, out};
    G.cs_nx[q] = kb_ceil_div(M, rpb);
    G.cs_first[q + 1] = G.cs_first[q] + G.cs_nx[q] * kb_ceil_div(N, 32);
    return KB_OK;
  }
  colsum_kernel<<<dim3(kb_ceil_div(M, rpb), kb_ceil_div(N, 32)), 256, 0, st>>>(X, dtype, ldx, group_rows, group_pitch, M, N, rpb, out);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

// ---- grouped launches (see GemmGroup) ----
int kbk_gemm_group_begin() {
  KB_CHECK_ARG(!g_pending.open, "gemm group: already open on this thread");
  memset(&g_pending.grp, 0, sizeof(g_pending.grp));
  g_pending.open = true;
  g_pending.tf32 = 0;
  return KB_OK;
}

// launches what has been collected (one kernel) and keeps the group open for the next batch of independent problems
int kbk_gemm_group_flush(cudaStream_t st) {
  KB_CHECK_ARG(g_pending.open, "gemm group: not open");
  GemmGroup& G = g_pending.grp;
  const int tiles = G.first[G.n] + G.cs_first[G.ncs];
  if (tiles > 0) {
    if (g_pending.tf32) gemm_group_kernel<true><<<(unsigned)tiles, 256, 0, st>>>(G);
    else gemm_group_kernel<false><<<(unsigned)tiles, 256, 0, st>>>(G);
    KB_CUDA_LAUNCH_CHECK();
  }
  memset(&G, 0, sizeof(G));
  return KB_OK;
}

int kbk_gemm_group_end(cudaStream_t st) {
  if (!g_pending.open) return KB_OK;
  const int r = kbk_gemm_group_flush(st);
  g_pending.open = false;
  return r;
}
