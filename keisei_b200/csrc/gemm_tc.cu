// gemm_tc.cu — Linear / 1x1-conv layers on the tcgen05 tensor path (bf16 in, fp32 accumulate).
//
//   Y[m, n] = act( (sum_k X[m, k] * W[n, k]) * scale[n] + bias[n] )
//
// Same machine as conv_tc.cu with the im2col removed: the OUTPUT FEATURE n sits on the 128 TMEM lanes
// (A = packed weights [Np][Kp], 2-D TMA box 64 x 128), 256 rows m of the input matrix are the
// accumulator columns (B = X [M][Kp], 2-D TMA box 64 x 256, rows past M zero-filled), K walks in
// 64-wide blocks through a 4-stage mbarrier ring, accumulators double-buffered in TMEM. The epilogue
// thread owns one output feature: per-feature scale/bias (folded BatchNorm of the policy head),
// ReLU, and coalesced stores (32 consecutive features per warp) to an fp32 tensor and/or a bf16
// tensor (the next layer's TMA operand; features >= N are written as zeros up to nb_store so the
// K padding of the consumer is valid). Rows may be board-pitched (the (B, 11264) policy buffer).
//
// Replaces nn.Linear / 1x1 nn.Conv2d at reference se_resnet.py:57-61 (global_fc), :63-66,:85-86 (SE),
// :119-130,:144-157 (policy / value / score heads) in the bf16 path.
#include <cuda.h>
#include <stdlib.h>
#include "kb_common.cuh"
#include "tc_ptx.cuh"
#include "kb_kernels.h"

namespace {

using namespace tcptx;

constexpr int kStages = 4;
constexpr int kTileM = 128;   // output features per tile (TMEM lanes)
constexpr int kTileN = 256;   // input rows per tile (TMEM columns)
constexpr int kBlockK = 64;
constexpr int kABytes = kTileM * kBlockK * 2;
constexpr int kBBytes = kTileN * kBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

struct LinArgs {
  const float* scale;   // [N] or null
  const float* bias;    // [N] or null
  int relu;
  float* out_f32; long long ld_f;     // [M][ld_f], features < N
  bf16* out_bf; long long ld_b;       // [M][ld_b], features < nb_store (zeros for n >= N)
  int nb_store;
  int group_rows; long long group_pitch;  // bf16 output only: row m at (m/gr)*pitch + (m%gr)*ld_b
  int M, N, Kp, Np;
};

__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, LinArgs g,
                 int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 2 + b); };
  const uint32_t holder = bar_base + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ct = g.Np / kTileM;
  const int num_kb = g.Kp / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
    fence_barrier_init();
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  pdl_wait();
  pdl_launch_dependents();

  // single-issuer loops run warp-uniformly, one elected lane issues (see tc_ptx.cuh: elect_one_sync)
  if (warp == 0) {
    const bool issuer = elect_one_sync();
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int ct = t % n_ct, mt = t / n_ct;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * kStageBytes;
        if (issuer) {
          mbar_arrive_expect_tx(full_bar(stage), kABytes + kBBytes);
          tma_load_2d(a_dst, &map_w, full_bar(stage), kb * kBlockK, ct * kTileM);
          tma_load_2d(a_dst + kABytes, &map_x, full_bar(stage), kb * kBlockK, mt * kTileN);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool issuer = elect_one_sync();
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(buf), tphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kTileN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * kStageBytes;
        const uint64_t adesc = smem_desc_k128(a_addr);
        const uint64_t bdesc = smem_desc_k128(a_addr + kABytes);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int lane_grp = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int ct = t % n_ct, mt = t / n_ct;
      const int buf = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int n = ct * kTileM + lane_grp * 32 + lane;
      const int m0 = mt * kTileN;
      const bool n_ok = n < g.N;
      const float sc = (g.scale && n_ok) ? g.scale[n] : 1.f;
      const float bi = (g.bias && n_ok) ? g.bias[n] : 0.f;
      const bool st_f = g.out_f32 != nullptr && n_ok;
      const bool st_b = g.out_bf != nullptr && n < g.nb_store;
      mbar_wait(tfull_bar(buf), tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * kTileN);
      const int rows = min(kTileN, g.M - m0);
      // Epilogue address arithmetic is hoisted out of the per-column code: both output pointers advance
      // by one row per column (board-pitched rows add the pitch adjustment every `group_rows` rows).
      float* pf = g.out_f32 + ((long long)m0 * g.ld_f + n);
      long long boff = (long long)m0 * g.ld_b;
      int rem = 0;
      long long adj = 0;
      if (g.group_rows > 0) {
        rem = m0 % g.group_rows;
        boff = (long long)(m0 / g.group_rows) * g.group_pitch + (long long)rem * g.ld_b;
        adj = g.group_pitch - (long long)g.group_rows * g.ld_b;
      }
      bf16* pb = g.out_bf + (boff + n);
      const long long ldf = g.ld_f, ldb = g.ld_b;
      const int gr = g.group_rows;
      const float scm = n_ok ? sc : 0.f, bim = n_ok ? bi : 0.f;  // features past N come out as exact zeros
      if (rows == kTileN && gr == 0) {
        // full tile, plain rows: no per-column predicates at all
        for (int bt = 0; bt < kTileN / 32; ++bt) {
          uint32_t r[2][16];
          tmem_ld16(taddr + bt * 32, r[0]);
          tmem_ld16(taddr + bt * 32 + 16, r[1]);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float v = fmaf(__uint_as_float(r[q][i]), scm, bim);
              if (g.relu) v = fmaxf(v, 0.f);
              if (st_f) pf[(long long)(q * 16 + i) * ldf] = v;
              if (st_b) pb[(long long)(q * 16 + i) * ldb] = __float2bfloat16_rn(v);
            }
          }
          pf += 32 * ldf;
          pb += 32 * ldb;
        }
      } else {
        for (int bt = 0; bt < kTileN / 32; ++bt) {
          if (bt * 32 >= rows) break;
          uint32_t r[2][16];
          tmem_ld16(taddr + bt * 32, r[0]);
          tmem_ld16(taddr + bt * 32 + 16, r[1]);
          tmem_ld_wait();
          const int nvalid = min(32, rows - bt * 32);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (q * 16 + i < nvalid) {
                float v = fmaf(__uint_as_float(r[q][i]), scm, bim);
                if (g.relu) v = fmaxf(v, 0.f);
                if (st_f) *pf = v;
                if (st_b) *pb = __float2bfloat16_rn(v);
              }
              pf += ldf;
              pb += ldb;
              if (gr > 0 && ++rem == gr) { rem = 0; pb += adj; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// w [N][K] fp32 -> out [Np][Kp] bf16, zero padded
__global__ void pack_linear_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int N, int K, int Np, int Kp) {
  const long long n = (long long)Np * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp), r = (int)(i / Kp);
    out[i] = __float2bfloat16_rn((r < N && k < K) ? w[(size_t)r * K + k] : 0.f);
  }
}

// in [rows][cols] fp32 -> out [rows][ld] bf16 (columns >= cols zero)
__global__ void cast_rows_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long rows, int cols, int ld) {
  const long long n = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld); const long long r = i / ld;
    out[i] = __float2bfloat16_rn(c < cols ? in[r * cols + c] : 0.f);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}
int make_map_2d(CUtensorMap* m, const void* base, long long rows, int K, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  KB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%lld K=%d) failed: %d", rows, K, (int)r);
  return KB_OK;
}

}  // namespace

int kbk_pack_linear_weight(const float* w, void* out, int N, int K, int Np, int Kp, cudaStream_t st) {
  const long long n = (long long)Np * Kp;
  pack_linear_weight_kernel<<<(int)min((long long)592, (n + 255) / 256), 256, 0, st>>>(w, (bf16*)out, N, K, Np, Kp);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_cast_rows_bf16(const float* in, void* out, long long rows, int cols, int ld, cudaStream_t st) {
  const long long n = rows * ld;
  if (n == 0) return KB_OK;
  cast_rows_bf16_kernel<<<(int)min((long long)1184, (n + 255) / 256), 256, 0, st>>>(in, (bf16*)out, rows, cols, ld);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

int kbk_linear_tc(const void* x, long long M, int Kp, const void* w, int N, int Np, const float* scale, const float* bias,
                  int relu, float* out_f32, long long ld_f, void* out_bf, long long ld_b, int nb_store, int group_rows,
                  long long group_pitch, int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(Kp % kBlockK == 0 && Np % kTileM == 0 && N <= Np && M >= 0, "linear_tc: bad shape M=%lld Kp=%d N=%d Np=%d", M, Kp, N, Np);
  KB_CHECK_ARG(M < (1ll << 31), "linear_tc: M too large");
  if (M == 0) return KB_OK;
  CUtensorMap mw, mx;
  if (int r = make_map_2d(&mw, w, Np, Kp, kTileM)) return r;
  if (int r = make_map_2d(&mx, x, M, Kp, kTileN)) return r;
  LinArgs g;
  g.scale = scale; g.bias = bias; g.relu = relu; g.out_f32 = out_f32; g.ld_f = ld_f; g.out_bf = (bf16*)out_bf; g.ld_b = ld_b;
  g.nb_store = nb_store; g.group_rows = group_rows; g.group_pitch = group_pitch; g.M = (int)M; g.N = N; g.Kp = Kp; g.Np = Np;
  const int num_tiles = kb_ceil_div(M, kTileN) * (Np / kTileM);
  if (num_sms <= 0) num_sms = 148;
  const int grid_tiles = num_tiles;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  KB_CUDA_CHECK(kb_launch_pdl(linear_tc_kernel, grid_tiles < num_sms ? grid_tiles : num_sms, kThreads, kSmemBytes, st, mw, mx, g, num_tiles));
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
