// gpool_mlp_tc.cu — the global-pool-bias MLP of a GlobalPoolBiasBlock as ONE tcgen05 kernel (bf16 in, fp32 accumulate):
//
//   g[b, :] = W2 · relu(W1 · pool[b, :] + b1) + b2          reference se_resnet.py:57-61 (global_fc), :73-78
//
// Round 1 ran the two Linear layers as two `linear_tc_kernel` launches of 8-18 CTAs each (256-row tiles), 16 us apiece and
// 5.5 % of the rollout step / a quarter of a 512-board league step. Here a CTA takes 128 boards through BOTH layers:
//   layer 1  D1[hidden 128 (TMEM lanes)][128 boards] = W1[128 x 768] · X[boards x 768]^T, K in 64-wide blocks through a
//            3-stage TMA ring (A = packed W1, B = the bf16 pool statistics);
//   between  the 4 epilogue warps (thread = hidden unit) read D1, add b1, ReLU, optionally store the fp32 hidden layer
//            (saved for the backward in training) and write it as bf16 straight into shared memory IN THE UMMA OPERAND
//            LAYOUT (K-major rows of 128 bytes, 128-byte swizzle: 16-byte chunk c of row r lands at chunk c ^ (r & 7)),
//            fence.proxy.async, barrier;
//   layer 2  D2[256 channels = 2 x 128 lanes][128 boards] = W2[256 x 128] · hidden^T with W2 resident in shared memory
//            (64 KB, one TMA per CTA lifetime), accumulators next to D1 in TMEM (128 + 256 of the 512 columns);
//   epilogue + b2, fp32 store of g (B, 256): the per-(board, channel) bias conv1's epilogue adds after its ReLU.
// Shapes: hidden = 128 (Gp = Gk = 128), channels = 256, K = 768 — the 40 x 256 network; anything else keeps the two
// generic launches.
#include <cuda.h>
#include "kb_common.cuh"
#include "tc_ptx.cuh"
#include "kb_kernels.h"

namespace {

using namespace tcptx;

constexpr int kStages = 3;
constexpr int kRows = 128;                      // boards per tile (TMEM columns of D1 / of each D2 half)
constexpr int kHid = 128;                       // hidden units = TMEM lanes of layer 1 = K of layer 2
constexpr int kCh = 256;                        // output channels (two 128-lane halves)
constexpr int kBlockK = 64;
constexpr int kABytes = kHid * kBlockK * 2;     // 16 KB: one K block of W1 (or one (half, K block) tile of W2)
constexpr int kBBytes = kRows * kBlockK * 2;    // 16 KB: one K block of the pool statistics / of the hidden layer
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kW2Bytes = 2 * 2 * kABytes;       // [channel half][K block]
constexpr int kHBytes = 2 * kBBytes;            // [K block]
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + kW2Bytes + kHBytes + 1024 + 256;
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct MlpArgs {
  const float* b1; const float* b2;
  float* gh_out;      // [B][128] fp32 hidden layer (post-ReLU) or null
  float* g_out;       // [B][256] fp32
  int B, K;           // K = 768 (multiple of 64)
};

__global__ void __launch_bounds__(kThreads, 1)
gpool_mlp_tc_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_w2, MlpArgs g, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w2_base = smem_base + kStages * kStageBytes;
  const uint32_t h_base = w2_base + kW2Bytes;
  uint8_t* h_gen = smem_gen + kStages * kStageBytes + kW2Bytes;
  const uint32_t bar_base = h_base + kHBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t w2_full = bar_base + 8u * (2 * kStages), d1_full = w2_full + 8, d1_empty = w2_full + 16, h_full = w2_full + 24,
                 h_empty = w2_full + 32, d2_full = w2_full + 40, d2_empty = w2_full + 48, holder = w2_full + 56;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + kW2Bytes + kHBytes + 8 * (2 * kStages) + 56);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = g.K / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(w2_full, 1); mbar_init(d1_full, 1); mbar_init(d1_empty, 4); mbar_init(h_full, 128); mbar_init(h_empty, 1);
    mbar_init(d2_full, 1); mbar_init(d2_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w2);
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
  const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + kRows;   // D1: columns 0..127, D2: 128..383

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    const bool issuer = elect_one_sync();
    if (issuer) {   // W2 stays resident: [half][K block] tiles of 128 rows x 64 k
      mbar_arrive_expect_tx(w2_full, kW2Bytes);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) tma_load_2d(w2_base + (half * 2 + kb) * kABytes, &map_w2, w2_full, kb * kBlockK, half * 128);
    }
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * kStageBytes;
        if (issuer) {
          mbar_arrive_expect_tx(full_bar(stage), kABytes + kBBytes);
          tma_load_2d(a_dst, &map_w1, full_bar(stage), kb * kBlockK, 0);
          tma_load_2d(a_dst + kABytes, &map_x, full_bar(stage), kb * kBlockK, t * kRows);   // rows past B: zero fill
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool issuer = elect_one_sync();
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t tp = (uint32_t)it & 1u;
      mbar_wait(d1_empty, tp ^ 1u);           // the epilogue has read the previous tile's D1
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * kStageBytes;
        const uint64_t adesc = smem_desc_k128(a_addr), bdesc = smem_desc_k128(a_addr + kABytes);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(tmem_d1, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(d1_full);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      // layer 2: hidden (shared memory, written by the epilogue warps) x resident W2
      if (it == 0) mbar_wait(w2_full, 0);
      mbar_wait(h_full, tp);
      mbar_wait(d2_empty, tp ^ 1u);
      tc_fence_after();
      if (issuer) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t adesc = smem_desc_k128(w2_base + (half * 2 + kb) * kABytes), bdesc = smem_desc_k128(h_base + kb * kBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16(tmem_d2 + (uint32_t)(half * kRows), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
        umma_commit(h_empty);    // the hidden-layer operand may be overwritten once these MMAs have retired
        umma_commit(d2_full);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: thread = hidden unit (layer 1) / channel within a half (layer 2) =====
    const int lane_grp = warp & 3;
    const int f = lane_grp * 32 + lane;
    const float b1 = g.b1[f];
    const float b2a = g.b2[f], b2b = g.b2[128 + f];
    const uint32_t lane_off = (uint32_t)(lane_grp * 32) << 16;
    // byte offset of (row r, this thread's k = f % 64) inside a K block of the swizzled hidden operand
    const int kblk = f >> 6, kk = f & 63, chunk = kk >> 3, inner = (kk & 7) * 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t tp = (uint32_t)it & 1u;
      const int r0 = t * kRows;
      const int rows = min(kRows, g.B - r0);
      mbar_wait(d1_full, tp);
      mbar_wait(h_empty, tp ^ 1u);             // layer 2 of the previous tile no longer reads the hidden operand
      tc_fence_after();
      uint8_t* hrow = h_gen + kblk * kBBytes + inner;
#pragma unroll 1
      for (int bt = 0; bt < kRows / 16; ++bt) {
        uint32_t r[16];
        tmem_ld16(tmem_d1 + lane_off + bt * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int row = bt * 16 + i;
          const float v = fmaxf(__uint_as_float(r[i]) + b1, 0.f);
          if (g.gh_out != nullptr && row < rows) g.gh_out[(size_t)(r0 + row) * kHid + f] = v;
          *reinterpret_cast<bf16*>(hrow + row * 128 + ((chunk ^ (row & 7)) << 4)) = __float2bfloat16_rn(v);
        }
      }
      tc_fence_before();
      fence_proxy_async();                     // generic-proxy stores -> visible to the UMMA (async proxy) reads
      mbar_arrive(h_full);
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty);
      // layer 2 epilogue
      mbar_wait(d2_full, tp);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float bias = half ? b2b : b2a;
        float* out = g.g_out + (size_t)r0 * kCh + half * 128 + f;
#pragma unroll 1
        for (int bt = 0; bt < kRows / 16; ++bt) {
          uint32_t r[16];
          tmem_ld16(tmem_d2 + lane_off + (uint32_t)(half * kRows) + bt * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int row = bt * 16 + i;
            if (row < rows) out[(size_t)row * kCh] = __uint_as_float(r[i]) + bias;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
int make_map(CUtensorMap* m, const void* base, long long rows, int K, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  KB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(gpool mlp, rows=%lld K=%d) failed: %d", rows, K, (int)r);
  return KB_OK;
}

}  // namespace

int kbk_gpool_mlp_tc_supported(int C, int G) { return C == kCh && G == kHid; }

// x_bf16 [B][3C] (bf16, pitch 3C); w1 packed [128][3C] bf16; w2 packed [256][128] bf16 (kbk_pack_linear_weight layouts)
int kbk_gpool_mlp_tc(const void* x_bf16, int B, int K, const void* w1, const float* b1, const void* w2, const float* b2,
                     float* gh_out, float* g_out, int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(x_bf16 && w1 && b1 && w2 && b2 && g_out, "gpool_mlp_tc: null pointer");
  KB_CHECK_ARG(B >= 0 && K % kBlockK == 0 && K >= kBlockK, "gpool_mlp_tc: bad shape B=%d K=%d", B, K);
  if (B == 0) return KB_OK;
  CUtensorMap mw1, mx, mw2;
  if (int r = make_map(&mw1, w1, kHid, K, kHid)) return r;
  if (int r = make_map(&mx, x_bf16, B, K, kRows)) return r;
  if (int r = make_map(&mw2, w2, kCh, kHid, 128)) return r;
  MlpArgs g;
  g.b1 = b1; g.b2 = b2; g.gh_out = gh_out; g.g_out = g_out; g.B = B; g.K = K;
  const int num_tiles = kb_ceil_div(B, kRows);
  if (num_sms <= 0) num_sms = 148;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(gpool_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  KB_CUDA_CHECK(kb_launch_pdl(gpool_mlp_tc_kernel, num_tiles < num_sms ? num_tiles : num_sms, kThreads, kSmemBytes, st, mw1, mx, mw2, g, num_tiles));
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
