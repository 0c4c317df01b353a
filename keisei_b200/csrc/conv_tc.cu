// conv_tc.cu — 3x3 convolution on 9x9 boards as a tcgen05 / TMEM implicit GEMM fed by TMA (bf16).
//
//   D[cout, pixel] = sum_{tap, cin} W[cout, tap*Cin + cin] * X[board, y+dy-1, x+dx-1, cin]
//
// GEMM view per CTA tile:  M = 128 output channels (TMEM lanes), N = 256 pixel columns = 3 whole
// boards (243 valid + 13 ignored), K = 9*Cin walked in 64-wide blocks (one tap x 64 input channels).
//   * A (weights)      : 2-D TMA box [128 rows x 64 k] from the packed [Cout][9*Cin] matrix.
//   * B (activations)  : 4-D TMA box [3 boards x 9 x 9 x 64 ch] from the NHWC tensor with the box
//                        origin shifted by the tap (dx-1, dy-1): out-of-bounds rows/cols (the conv
//                        padding) and boards past the batch are zero-filled by the TMA unit, so the
//                        im2col never exists in memory and needs no halo in HBM.
//   Both land in shared memory as K-major 128-byte rows with the 128B swizzle; one elected thread
//   issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=256, K=16) four times per K block into a
//   256-column fp32 accumulator in TMEM; the accumulator is double buffered (2 x 256 = all 512
//   columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * Epilogue: 4 warps, thread = output channel (TMEM lane), columns = pixels, so BatchNorm
//     statistics, SE squeeze, global-pool statistics, ReLU masks and bias are register-local
//     (conv_epilogue.cuh). 16 tcgen05.ld.32x32b.x16 per tile.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
// Persistent: grid = min(#SM, tiles); tile t -> (board group t / n_ct, channel tile t % n_ct) so
// CTAs running together share the same boards in L2.
//
// Replaces F.conv2d at reference se_resnet.py:50,52,110 (forward) and its data gradient (same
// kernel on the flipped/transposed weight pack). The weight gradient is conv3x3_wgrad below.
#include <cuda.h>
#include <stdlib.h>
#include "kb_common.cuh"
#include "tc_ptx.cuh"
#include "conv_epilogue.cuh"
#include "kb_kernels.h"

namespace {

constexpr int kStages = 4;
constexpr int kTileM = 128;                    // output channels per tile (TMEM lanes)
constexpr int kTileN = 256;                    // pixel columns per tile
constexpr int kBoards = 3;                     // whole boards per tile
constexpr int kBlockK = 64;                    // bf16 elements per K block = one 128-byte swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB
constexpr int kBBytes = kTileN * kBlockK * 2;  // 32 KB (31,104 written by TMA)
constexpr int kBTxBytes = kBoards * 81 * kBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

using namespace tcptx;

// instruction descriptor: fp32 accumulate, bf16 A/B, both K-major, N (>>3) at bit 17, M (>>4) at bit 24
constexpr uint32_t kIdescF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// ---------------------------------------------------------------- forward / dgrad kernel
// kCl = true: launched as clusters of 2 CTAs that own the two 128-channel halves of the SAME board group.
// The activation (B) tile is identical for both, so each CTA fetches only part of it (rank 0: boards 0-1,
// rank 1: board 2) and TMA-multicasts it into both CTAs' shared memory: L2 -> SM operand traffic per tile
// drops from 1.77 MB to 1.14 MB (ncu: the single-CTA kernel's MMA issuer stalls on the full barriers).
// Stage release then needs both consumers: tcgen05.commit multicasts its arrive to both CTAs' empty barrier.
template <int F, bool kCl>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_x1,
                  bf16* __restrict__ out, int B, int Cin, int Cout, int num_tiles, ConvEpi epi) {
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
  const int n_ct = Cout / kTileM;
  const int kb_per_tap = Cin / kBlockK;
  const int num_kb = 9 * kb_per_tap;
  // tile walk: single CTA -> tiles blockIdx.x, +gridDim.x, ...; cluster -> this CTA always takes channel half
  // `crank` of board groups cluster_id, +num_clusters, ... (both CTAs of a cluster run the same group sequence)
  const uint32_t crank = kCl ? cluster_ctarank() : 0u;
  const int t_first = kCl ? (int)(blockIdx.x >> 1) * 2 + (int)crank : (int)blockIdx.x;
  const int t_step = (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kCl ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
    fence_barrier_init();
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(kCl ? (crank == 0 ? &map_x2 : &map_x1) : &map_x);
  }
  if (warp == 1) {  // TMEM owner: all 512 columns (two 256-column accumulators)
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (kCl) cluster_sync_all();  // peer barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = t_first; t < num_tiles; t += t_step) {
        const int ct = t % n_ct, grp = t / n_ct;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kb_per_tap, cc = kb - tap * kb_per_tap;
          mbar_wait(empty_bar(stage), phase ^ 1u);  // cluster: both CTAs' MMAs have retired this stage
          mbar_arrive_expect_tx(full_bar(stage), kABytes + kBTxBytes);
          const uint32_t a_dst = smem_base + stage * kStageBytes;
          tma_load_2d(a_dst, &map_w, full_bar(stage), kb * kBlockK, ct * kTileM);
          if (kCl) {
            if (crank == 0)
              tma_load_4d_mc(a_dst + kABytes, &map_x2, full_bar(stage), cc * kBlockK, tap % 3 - 1, tap / 3 - 1, grp * kBoards, (uint16_t)3);
            else
              tma_load_4d_mc(a_dst + kABytes + 2 * 81 * 128, &map_x1, full_bar(stage), cc * kBlockK, tap % 3 - 1, tap / 3 - 1,
                             grp * kBoards + 2, (uint16_t)3);
          } else {
            tma_load_4d(a_dst + kABytes, &map_x, full_bar(stage), cc * kBlockK, tap % 3 - 1, tap / 3 - 1, grp * kBoards);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = t_first; t < num_tiles; t += t_step, ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), tphase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kTileN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * kStageBytes;
          const uint64_t adesc = smem_desc_k128(a_addr);
          const uint64_t bdesc = smem_desc_k128(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdescF16, (kb | k) != 0 ? 1u : 0u);
          }
          if (kCl) umma_commit_mc(empty_bar(stage), (uint16_t)3);  // both producers write into this stage
          else umma_commit(empty_bar(stage));                      // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(tfull_bar(buf));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue: thread = output channel =====
    const int lane_grp = warp & 3;  // TMEM lanes 32*lane_grp .. +31 are the ones this warp may read
    int it = 0;
    for (int t = t_first; t < num_tiles; t += t_step, ++it) {
      const int ct = t % n_ct, grp = t / n_ct;
      const int buf = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int c = ct * kTileM + lane_grp * 32 + lane;
      const int b0 = grp * kBoards;
      const int nb_valid = min(kBoards, B - b0);
      mbar_wait(tfull_bar(buf), tphase);
      tc_fence_after();
      ConvEpiThread<bf16, kBoards, F> et(epi, c, Cout, B);
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * kTileN);
      // Mask-source prefetch (data-gradient epilogue): the ReLU mask / BN-backward statistics need one
      // activation element per accumulator. Loads are issued kPf chunks (of 16 columns) ahead of their
      // use so ~64 independent loads per thread are in flight instead of one dependent load per column.
      constexpr bool kMask = (F != kEpiDynamic) && (F & kEpiMask) != 0;
      constexpr int kPf = 4, kChunks = (kBoards * 81 + 15) / 16;
      bf16 pf[kPf + 1][16];
      const bf16* mbase = kMask ? (const bf16*)epi.mask_src + et.index(b0, 0) : nullptr;
      auto prefetch = [&](int ch) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = ch * 16 + i;  // boards are contiguous in memory: element (col) sits col*Cout further
          pf[ch % (kPf + 1)][i] = (col < kBoards * 81 && col / 81 < nb_valid) ? mbase[(unsigned)col * (unsigned)Cout] : bf16(0.f);
        }
      };
      if (kMask) {
#pragma unroll
        for (int ch = 0; ch < kPf; ++ch) prefetch(ch);
      }
#pragma unroll
      for (int ch = 0; ch < kTileN / 16; ++ch) {
        if (ch * 16 < kBoards * 81) {
          if (kMask && ch + kPf < kChunks) prefetch(ch + kPf);
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = ch * 16 + i;
            if (col < kBoards * 81) {
              const int j = col / 81, p = col % 81;
              if (j < nb_valid) {
                if (p == 0) et.begin_board(b0 + j, out);
                float ms = 0.f;
                if (kMask) ms = __bfloat162float(pf[ch % (kPf + 1)][i]);
                else if (F == kEpiDynamic && epi.mask_src != nullptr) ms = __bfloat162float(((const bf16*)epi.mask_src)[et.index(b0 + j, p)]);
                et.value(j, p, __uint_as_float(r[i]), ms);
                if (p == 80) et.board_done(j, b0 + j);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      et.finish(nb_valid);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCl) cluster_sync_all();  // the peer may still multicast into / signal this CTA until it is done too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight-gradient kernel
// dW[co][ci][tap] = sum_{board, pixel} dY[board][pixel][co] * X[board][pixel + shift(tap)][ci]
//
// GEMM view per CTA: one (tap, 128-channel half of Cout) unit and one slice of the boards.
//   D[co (128 TMEM lanes), ci (N = Cin columns)] += A[co, k] * B[ci, k],  k = pixels of one board.
// Both operands are "MN-major": in NHWC memory the channel index is contiguous for a fixed pixel k.
// TMA lands each board as 64-channel groups of [81 pixel rows x 128 B] (swizzle 128B); rows 81..95
// of every group are zeroed once and never written, so K is padded 81 -> 96 = 6 x UMMA_K for free.
// The tap shift is the TMA box origin on X (zero fill at the board edge), exactly as in the forward.
// Partial tiles go to a workspace with coalesced stores; wgrad_reduce_kernel sums the board slices
// and scatters into the PyTorch (Cout, Cin, 3, 3) fp32 gradient.
constexpr int kWgStages = 3;
constexpr int kWgRows = 96;                          // 81 pixels padded to a multiple of UMMA_K
constexpr int kWgGroupBytes = kWgRows * 128;         // one 64-channel group of one board: 12 KB
constexpr int kWgBoxBytes = 81 * 128;                // bytes TMA writes per group
constexpr int kWgAGroups = 2;                        // 128 output channels
constexpr int kWgMaxBGroups = 4;                     // up to 256 input channels
constexpr int kWgStageBytes = (kWgAGroups + kWgMaxBGroups) * kWgGroupBytes;  // 72 KB
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 1024 + 256;

// MN-major operand, SWIZZLE_128B: 64-element groups LBO apart, 8-row (k) groups SBO = 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_mn128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(kWgGroupBytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                        float* __restrict__ ws, int B, int Cin, int Cout, int units, int boards_per_slice) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kWgStages * kWgStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgStages);
  const uint32_t holder = bar_base + 8u * (2 * kWgStages + 1);
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kWgStages * kWgStageBytes + 8 * (2 * kWgStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x % units, slice = blockIdx.x / units;
  const int n_half = Cout / kTileM;
  const int tap = unit / n_half, half = unit % n_half;
  const int b_groups = Cin / 64;
  const int b_begin = slice * boards_per_slice;
  const int b_end = min(B, b_begin + boards_per_slice);
  const int nboards = b_end - b_begin;

  // zero the whole operand area once: the K-padding rows (81..95 of every group) must read as 0
  {
    uint4* p = reinterpret_cast<uint4*>(smem_gen);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kWgStages * kWgStageBytes / 16; i += kThreads) p[i] = z;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1) {
    tmem_alloc(holder, 256);
    tmem_relinquish();
  }
  fence_proxy_async();  // generic-proxy zero fill -> visible to the async proxy (TMA / UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int b = b_begin; b < b_end; ++b) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)((kWgAGroups + b_groups) * kWgBoxBytes));
        const uint32_t a_dst = smem_base + stage * kWgStageBytes;
#pragma unroll
        for (int g = 0; g < kWgAGroups; ++g)
          tma_load_4d(a_dst + g * kWgGroupBytes, &map_dy, full_bar(stage), half * kTileM + g * 64, 0, 0, b);
        for (int g = 0; g < b_groups; ++g)
          tma_load_4d(a_dst + (kWgAGroups + g) * kWgGroupBytes, &map_x, full_bar(stage), g * 64, tap % 3 - 1, tap / 3 - 1, b);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 128, N = Cin, A and B MN-major (bits 15, 16), fp32 accumulate, bf16 inputs
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Cin >> 3) << 17) |
                             ((uint32_t)(kTileM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nboards; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * kWgStageBytes;
        const uint64_t adesc = smem_desc_mn128(a_addr);
        const uint64_t bdesc = smem_desc_mn128(a_addr + kWgAGroups * kWgGroupBytes);
#pragma unroll
        for (int k = 0; k < kWgRows / 16; ++k) {
          // 16 pixel rows = 2048 bytes further down every group: +128 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done_bar);
    }
  } else {
    const int lane_grp = warp & 3;
    float* dst = ws + ((size_t)blockIdx.x * Cin) * kTileM + lane_grp * 32 + lane;  // ws[cta][ci][co_local]
    if (nboards > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
      for (int ch = 0; ch < Cin / 16; ++ch) {
        uint32_t r[16];
        tmem_ld16(taddr + ch * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[(size_t)(ch * 16 + i) * kTileM] = __uint_as_float(r[i]);
      }
    } else {
      for (int ci = 0; ci < Cin; ++ci) dst[(size_t)ci * kTileM] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// dw[(co*Cin_true + ci)*9 + tap] += sum_slices ws[slice*units + tap*n_half + half][ci][co_local]
// The workspace is co-fastest, the PyTorch gradient is (ci, tap)-fastest: a CTA takes a 32 co x kRedCi ci x 9 tap brick,
// reads it with 128-byte rows (co contiguous), transposes through shared memory and writes, per output channel, one
// contiguous run of kRedCi*9 floats — both sides coalesced (the direct scatter used a 9 KB stride per thread).
constexpr int kRedCi = 2;  // 32 co x 2 ci x 9 taps per CTA: 1024 CTAs for a 256x256 conv, 2-3 rows per warp (the kernel is latency-bound)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cin, int Cout,
                                                           int Cin_true, int units, int slices) {
  __shared__ float tile[32][kRedCi * 9 + 1];  // [co][ci * 9 + tap]
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * kRedCi;
  const int n_half = Cout / kTileM;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // 8 warps
  const int co = co0 + lane;
  const size_t slice_stride = (size_t)units * Cin * kTileM;
  for (int r = wid; r < kRedCi * 9; r += 8) {  // r = ci_local * 9 + tap: one 128-byte row of the brick per warp step
    const int cil = r / 9, tap = r - cil * 9, ci = ci0 + cil;
    float s = 0.f;
    if (co < Cout && ci < Cin_true) {
      const int unit = tap * n_half + co / kTileM;
      const float* p = ws + ((size_t)unit * Cin + ci) * kTileM + (co % kTileM);
#pragma unroll 8
      for (int sl = 0; sl < slices; ++sl) s += p[(size_t)sl * slice_stride];
    }
    tile[lane][r] = s;
  }
  __syncthreads();
  const int ci_n = min(kRedCi, Cin_true - ci0);
  if (ci_n <= 0) return;
  const int run = ci_n * 9;  // contiguous floats per output channel
  for (int c = wid; c < 32; c += 8) {
    if (co0 + c >= Cout) break;
    float* out = dw + ((size_t)(co0 + c) * Cin_true + ci0) * 9;
    for (int i = lane; i < run; i += 32) out[i] += tile[c][i];
  }
}

// The tcgen05 convolutions are launched with the device's HIGHEST launch priority (a per-launch attribute, independent
// of the caller's stream): when a convolution and a streaming kernel are runnable at the same time (side-stream weight
// gradients in the backward, the two-branch rollout), the one-per-SM convolution CTAs are placed first and the
// short-lived streaming CTAs fill the registers and thread slots they leave, instead of keeping the SMs full until
// their kernel drains (two-branch rollout: 133.0 k vs 130.3 k positions/s). KB_CONV_PRIO=0 switches it off.
int conv_priority() {
  static int prio = 1 << 30;
  if (prio == (1 << 30)) {
    const char* e = getenv("KB_CONV_PRIO");
    int lo = 0, hi = 0;
    if ((e && e[0] == '0') || cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { cudaGetLastError(); prio = 0; }
    else prio = hi;
  }
  return prio;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_prio(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributePriority;
  attr[0].val.priority = conv_priority();
  cfg.attrs = attr; cfg.numAttrs = conv_priority() != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;  // resolved once; benign race (same value)
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

int make_weight_map(CUtensorMap* m, const void* w, int rows, int K) {
  EncodeTiledFn enc = get_encode_fn();
  KB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return KB_OK;
}

int make_act_map(CUtensorMap* m, const void* x, int B, int C, int boards_per_box) {
  EncodeTiledFn enc = get_encode_fn();
  KB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)C, 9, 9, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 9, (cuuint64_t)C * 2 * 81};
  const cuuint32_t box[4] = {(cuuint32_t)kBlockK, 9, 9, (cuuint32_t)boards_per_box};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
  return KB_OK;
}

template <int F>
int launch_fwd(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& mx2, const CUtensorMap& mx1, bf16* out, int B,
               int Cin, int Cout, int num_tiles, const ConvEpi& epi, int grid, bool cluster, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation; idempotent
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  if (cluster) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<F, true>, mw, mx, mx2, mx1, out, B, Cin, Cout, num_tiles, epi));
    kb_count_launch();
    return KB_OK;
  }
  launch_prio(conv3x3_tc_kernel<F, false>, grid, kThreads, kSmemBytes, st, mw, mx, mx2, mx1, out, B, Cin, Cout, num_tiles, epi);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

}  // namespace

int kbk_conv3x3_tc_supported(int Cin, int Cout, int dtype) {
  return dtype == KB_BF16 && Cin % kBlockK == 0 && Cout % kTileM == 0 && Cin >= kBlockK;
}

int kbk_conv3x3_tc(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi, int num_sms,
                   cudaStream_t st) {
  KB_CHECK_ARG(kbk_conv3x3_tc_supported(Cin, Cout, KB_BF16), "conv3x3_tc: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  if (B == 0) return KB_OK;
  CUtensorMap mw, mx, mx2, mx1;
  if (int r = make_weight_map(&mw, w, Cout, 9 * Cin)) return r;
  if (int r = make_act_map(&mx, in, B, Cin, kBoards)) return r;
  if (int r = make_act_map(&mx2, in, B, Cin, 2)) return r;
  if (int r = make_act_map(&mx1, in, B, Cin, 1)) return r;
  const int groups = kb_ceil_div(B, kBoards);
  const int n_ct = Cout / kTileM;
  const int num_tiles = groups * n_ct;
  if (num_sms <= 0) num_sms = 148;
  // cluster-of-2 multicast variant: the two channel halves of a board group share the activation tile.
  // Measured on B200 (round 1): parity-green but NOT faster (0.300 vs 0.293 ms per 256->256 conv at B=4096) — the
  // MMA issuer's full-barrier stalls are TMA latency vs the 4-stage (192 KB) ring, not L2 bytes. Opt-in only.
  static int cluster_env = -1;
  if (cluster_env < 0) { const char* e = getenv("KB_CONV_CLUSTER"); cluster_env = (e && e[0] == '1') ? 1 : 0; }
  const bool cluster = cluster_env && n_ct == 2 && groups >= 2;
  int grid = num_tiles < num_sms ? num_tiles : num_sms;
  if (cluster) { grid = 2 * (groups < num_sms / 2 ? groups : num_sms / 2); }
  bf16* o = (bf16*)out;
  const int f = conv_epi_features(epi);
#define KB_FWD(FF) return launch_fwd<FF>(mw, mx, mx2, mx1, o, B, Cin, Cout, num_tiles, epi, grid, cluster, st)
  switch (f) {
    case 0: KB_FWD(0);
    case kEpiSum | kEpiSumSq: KB_FWD(kEpiSum | kEpiSumSq);
    case kEpiSum | kEpiSumSq | kEpiBoard: KB_FWD(kEpiSum | kEpiSumSq | kEpiBoard);
    case kEpiAffine | kEpiRelu | kEpiGbias: KB_FWD(kEpiAffine | kEpiRelu | kEpiGbias);
    case kEpiAffine | kEpiRelu: KB_FWD(kEpiAffine | kEpiRelu);  // plain ResNet eval: conv1 / stem
    case kEpiAffine: KB_FWD(kEpiAffine);                        // plain ResNet eval: conv2
    case kEpiAffine | kEpiRelu | kEpiPool: KB_FWD(kEpiAffine | kEpiRelu | kEpiPool);
    case kEpiAffine | kEpiBoard: KB_FWD(kEpiAffine | kEpiBoard);
    case kEpiMask | kEpiSum | kEpiDot | kEpiBoard: KB_FWD(kEpiMask | kEpiSum | kEpiDot | kEpiBoard);
    default: KB_FWD(kEpiDynamic);
  }
#undef KB_FWD
}

long long kbk_conv3x3_wgrad_tc_ws_bytes(int Cin, int Cout, int num_sms) {
  if (num_sms <= 0) num_sms = 148;
  const int units = 9 * (Cout / kTileM);
  int slices = num_sms / units;
  if (slices < 1) slices = 1;
  return (long long)slices * units * Cin * kTileM * (long long)sizeof(float);
}

int kbk_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true, float* ws,
                         long long ws_bytes, int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(kbk_conv3x3_tc_supported(Cin, Cout, KB_BF16) && Cin <= 64 * kWgMaxBGroups, "conv3x3_wgrad_tc: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  if (B == 0) return KB_OK;
  if (num_sms <= 0) num_sms = 148;
  const int units = 9 * (Cout / kTileM);
  int slices = num_sms / units;
  if (slices < 1) slices = 1;
  if (slices > B) slices = B;
  const int bps = kb_ceil_div(B, slices);
  slices = kb_ceil_div(B, bps);
  KB_CHECK_ARG(ws != nullptr && ws_bytes >= (long long)slices * units * Cin * kTileM * (long long)sizeof(float),
               "conv3x3_wgrad_tc: workspace too small");
  CUtensorMap mdy, mx;
  if (int r = make_act_map(&mdy, dy, B, Cout, 1)) return r;
  if (int r = make_act_map(&mx, x, B, Cin, 1)) return r;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_set = true;
  }
  launch_prio(conv3x3_wgrad_tc_kernel, units * slices, kThreads, kWgSmemBytes, st, mdy, mx, ws, B, Cin, Cout, units, bps);
  KB_CUDA_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<dim3(kb_ceil_div(Cout, 32), kb_ceil_div(Cin_true, kRedCi)), 256, 0, st>>>(ws, dw, Cin, Cout, Cin_true, units, slices);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
