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

// One accumulator tile (128 channels x 3 boards) through the fused epilogue: thread = output channel `et.c`, columns =
// pixels. Shared by the single-CTA and the CTA-pair kernels.
template <int F>
__device__ __forceinline__ void epilogue_tile(const ConvEpi& epi, ConvEpiThread<bf16, kBoards, F>& et, uint32_t taddr, int b0,
                                              int nb_valid, int Cout, bf16* __restrict__ out) {
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
}

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
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer: the warp runs the loop uniformly, one elected lane issues =====
    const bool issuer = elect_one_sync();
    int stage = 0; uint32_t phase = 0;
    for (int t = t_first; t < num_tiles; t += t_step) {
      const int ct = t % n_ct, grp = t / n_ct;
      int tap = 0, cc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);  // cluster: both CTAs' MMAs have retired this stage
        const uint32_t a_dst = smem_base + stage * kStageBytes;
        const int sx = tap % 3 - 1, sy = tap / 3 - 1;
        if (issuer) {
          mbar_arrive_expect_tx(full_bar(stage), kABytes + kBTxBytes);
          tma_load_2d(a_dst, &map_w, full_bar(stage), kb * kBlockK, ct * kTileM);
          if (kCl) {
            if (crank == 0)
              tma_load_4d_mc(a_dst + kABytes, &map_x2, full_bar(stage), cc * kBlockK, sx, sy, grp * kBoards, (uint16_t)3);
            else
              tma_load_4d_mc(a_dst + kABytes + 2 * 81 * 128, &map_x1, full_bar(stage), cc * kBlockK, sx, sy, grp * kBoards + 2, (uint16_t)3);
          } else {
            tma_load_4d(a_dst + kABytes, &map_x, full_bar(stage), cc * kBlockK, sx, sy, grp * kBoards);
          }
        }
        if (++cc == kb_per_tap) { cc = 0; ++tap; }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: uniform loop, one elected lane issues the MMAs and their commits =====
    const bool issuer = elect_one_sync();
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
        if (issuer) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdescF16, (kb | k) != 0 ? 1u : 0u);
          }
          if (kCl) umma_commit_mc(empty_bar(stage), (uint16_t)3);  // both producers write into this stage
          else umma_commit(empty_bar(stage));                      // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
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
      epilogue_tile<F>(epi, et, taddr, b0, nb_valid, Cout, out);
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

// ---------------------------------------------------------------- fused evaluation tail (CTA-pair kernel only)
// conv2 of a GlobalPoolBiasBlock in eval mode, everything after the convolution in the epilogue (se_resnet.py:83-98):
//   v    = acc * scale + shift                      (folded BatchNorm 2)
//   mean = board mean of v  -> SE MLP (C -> S -> 2C) -> (sigmoid(se_scale), se_shift) per (board, channel)
//   out  = relu(v * sigmoid(se_scale) + se_shift + residual);   pool = (mean, max, population std)(out)
// The accumulator is read from TMEM twice. Pass 1 sums the 81 columns of each board per thread (= channel); the 3 x 256
// means cross the CTA pair through distributed shared memory (each thread stores its three means locally and, with
// st.async, into the peer, crediting the peer's exchange barrier with transaction bytes — no cluster-scope fence); both CTAs then run the small MLP redundantly — each warp takes 4 hidden
// units (W1 slices through L1), a warp-shuffle reduction per hidden unit, W2's two rows of this thread's channel in
// registers; pass 2 applies scale / shift / residual / ReLU, stores bf16 and accumulates the global-pool statistics of
// the stored values. Residual elements are requested kPf chunks of 16 columns ahead of their use (the first ones before
// pass 1); TMEM reads are double-buffered in registers (chunk k+1 in flight while chunk k is consumed). Boards never mix: results do not depend on what
// else is in the batch. This replaces se_mlp_fwd_kernel + se_apply_col_kernel<0> and two of the block's three
// activation-tensor passes in the rollout.
constexpr int kSeC = 256, kSeS = 16;
struct SeTailSmem {
  float mean[2][kBoards][kSeC];   // [tile parity][board][channel]: board means of both CTAs' channels
  float hid[2][kBoards][kSeS];    // hidden layer of the SE MLP
};
struct SeTailRegs {               // per-thread constants, loaded once per CTA lifetime
  float b1[4];                    // b1[4*lane_grp + q]
  float b2s, b2h, sc, sh;
};

__device__ __forceinline__ void se_tail_load(const ConvEpi& epi, int c, int lane_grp, int lane, SeTailRegs& R) {
#pragma unroll
  for (int q = 0; q < 4; ++q) R.b1[q] = epi.se_b1[4 * lane_grp + q];
  R.b2s = epi.se_b2[c]; R.b2h = epi.se_b2[kSeC + c];
  R.sc = epi.scale[c]; R.sh = epi.shift[c];
}

// TMEM -> registers, two 16-column chunks in flight: the load of chunk k+1 is issued before chunk k is consumed.
// body(r, chunk) sees compile-time chunk indices after unrolling.
template <int NCHUNKS, typename Fn>
__device__ __forceinline__ void tmem_pipeline(uint32_t taddr, Fn&& body) {
  uint32_t r0[16], r1[16];
  tmem_ld16(taddr, r0);
#pragma unroll
  for (int ch = 0; ch < NCHUNKS; ch += 2) {
    tmem_ld_wait();
    if (ch + 1 < NCHUNKS) tmem_ld16(taddr + (ch + 1) * 16, r1);
    body(r0, ch);
    if (ch + 1 < NCHUNKS) {
      tmem_ld_wait();
      if (ch + 2 < NCHUNKS) tmem_ld16(taddr + (ch + 2) * 16, r0);
      body(r1, ch + 1);
    }
  }
}

// two floats -> packed bf16x2 (round to nearest even) in ONE ALU-rate instruction (F2FP); a lone cvt.rn.bf16.f32 is an F2F
// on the quarter-rate conversion pipe. `lo` lands in bits 0..15.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// kFull: all three boards of the tile exist (every tile but possibly the last): no per-column validity predicates
template <bool kFull>
__device__ __forceinline__ void se_tail_tile(const ConvEpi& epi, const SeTailRegs& R, SeTailSmem* sm, uint32_t sm_addr, uint32_t se_bar,
                                             uint32_t crank, int g, uint32_t parity, uint32_t taddr, int c, int lane_grp, int lane,
                                             int b0, int nb_valid, bf16* __restrict__ out) {
  constexpr int kChunks = (kBoards * 81 + 15) / 16;
  constexpr int kPf = 3;
  const int pb = g;   // this warp group's exchange buffers
  // residual prefetch ring: chunk k + kPf is requested while chunk k is consumed; the first kPf chunks are requested
  // here, before pass 1, so their latency hides under it
  unsigned short pf[kPf + 1][16];
  const unsigned short* rbase = (const unsigned short*)epi.res + ((size_t)b0 * 81) * kSeC + c;
  auto prefetch = [&](int ch) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = ch * 16 + i;
      if (col < kBoards * 81) pf[ch % (kPf + 1)][i] = (kFull || col / 81 < nb_valid) ? __ldg(rbase + (unsigned)col * (unsigned)kSeC) : (unsigned short)0;
    }
  };
#pragma unroll
  for (int ch = 0; ch < kPf; ++ch) prefetch(ch);
  // ---- pass 1: board sums of the BatchNorm-2 output
  float bs[kBoards];
#pragma unroll
  for (int j = 0; j < kBoards; ++j) bs[j] = 0.f;
  tmem_pipeline<kChunks>(taddr, [&](const uint32_t(&r)[16], int chunk) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = chunk * 16 + i;
      if (col < kBoards * 81) bs[col / 81] += fmaf(__uint_as_float(r[i]), R.sc, R.sh);
    }
  });
  // ---- exchange the means across the pair
  {
    // own half: plain shared stores + a CTA-scope arrive; peer half: asynchronous DSMEM stores that credit transaction
    // bytes to the PEER's exchange barrier (128 threads x 3 floats = 1536 bytes per tile, expected by one local thread)
    const uint32_t own = sm_addr + (uint32_t)(((pb * kBoards) * kSeC + c) * 4);
    const uint32_t peer = mapa_peer(own, crank ^ 1u);
    const uint32_t peer_bar = mapa_peer(se_bar, crank ^ 1u);
#pragma unroll
    for (int j = 0; j < kBoards; ++j) {
      const float m = bs[j] * (1.f / 81.f);
      sm->mean[pb][j][c] = m;
      st_async_f32(peer + (uint32_t)(j * kSeC * 4), m, peer_bar);
    }
    if (lane_grp == 0 && lane == 0) mbar_arrive_expect_tx(se_bar, (uint32_t)(kTileM * kBoards * 4));
    else mbar_arrive(se_bar);
    mbar_wait(se_bar, parity);
  }
  // ---- SE MLP, hidden layer: this warp's 4 units for the 3 boards
  {
    float part[4][kBoards];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < kBoards; ++j) part[q][j] = 0.f;
    // W1 rows of this warp's 4 hidden units, [lane + 32*k] slices: 16 KB in all, L1-resident across tiles
    const float* w1p = epi.se_w1 + (size_t)(4 * lane_grp) * kSeC + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float wq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wq[q] = __ldg(w1p + q * kSeC + 32 * k);
#pragma unroll
      for (int j = 0; j < kBoards; ++j) {
        const float m = sm->mean[pb][j][lane + 32 * k];
#pragma unroll
        for (int q = 0; q < 4; ++q) part[q][j] = fmaf(wq[q], m, part[q][j]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < kBoards; ++j) {
        float v = part[q][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm->hid[pb][j][4 * lane_grp + q] = fmaxf(v + R.b1[q], 0.f);
      }
  }
  named_bar_sync(1 + g, 128);   // the four warps of this epilogue group
  float sig[kBoards], shf[kBoards];
  {
    // W2 rows c and C + c (16 floats each, 64 contiguous bytes: four float4 loads per row, L1 / L2 resident)
    const float4* w2s = reinterpret_cast<const float4*>(epi.se_w2 + (size_t)c * kSeS);
    const float4* w2h = reinterpret_cast<const float4*>(epi.se_w2 + (size_t)(kSeC + c) * kSeS);
    float a[kBoards], b[kBoards];
#pragma unroll
    for (int j = 0; j < kBoards; ++j) { a[j] = R.b2s; b[j] = R.b2h; }
#pragma unroll
    for (int s4 = 0; s4 < kSeS / 4; ++s4) {
      const float4 ws = __ldg(w2s + s4), wh = __ldg(w2h + s4);
#pragma unroll
      for (int j = 0; j < kBoards; ++j) {
        const float4 h = *reinterpret_cast<const float4*>(&sm->hid[pb][j][4 * s4]);
        a[j] = fmaf(ws.x, h.x, a[j]); a[j] = fmaf(ws.y, h.y, a[j]); a[j] = fmaf(ws.z, h.z, a[j]); a[j] = fmaf(ws.w, h.w, a[j]);
        b[j] = fmaf(wh.x, h.x, b[j]); b[j] = fmaf(wh.y, h.y, b[j]); b[j] = fmaf(wh.z, h.z, b[j]); b[j] = fmaf(wh.w, h.w, b[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kBoards; ++j) { sig[j] = 1.f / (1.f + __expf(-a[j])); shf[j] = b[j]; }
  }
  // ---- pass 2: scale / shift / residual / ReLU, store, global-pool statistics of the stored values.
  // Columns go in pairs so that the bf16 rounding is one packed conversion per two outputs.
  float mx = -INFINITY, k0 = 0.f, ds = 0.f, dss = 0.f;
  unsigned short* optr = reinterpret_cast<unsigned short*>(out);
  auto finish_value = [&](int col, uint32_t bits) {   // bits: the stored bf16 in the low half
    const int j = col / 81, p = col % 81;
    if (kFull || j < nb_valid) {
      if (p == 0) { optr = reinterpret_cast<unsigned short*>(out) + ((size_t)(b0 + j) * 81) * kSeC + c; mx = -INFINITY; ds = 0.f; dss = 0.f; }
      optr[(unsigned)p * (unsigned)kSeC] = (unsigned short)bits;
      const float rr = __uint_as_float(bits << 16);
      mx = fmaxf(mx, rr);
      if (p == 0) k0 = rr;
      const float d = rr - k0;   // shifted data: mean = k0 + ds / 81, variance without cancellation
      ds += d;
      dss = fmaf(d, d, dss);
      if (p == 80) {
        const float dm = ds * (1.f / 81.f);
        const float mean = k0 + dm;
        const float sd = sqrtf(fmaxf(dss * (1.f / 81.f) - dm * dm, 0.f));
        float* pr = epi.pool + (size_t)(b0 + j) * 3 * kSeC;
        pr[c] = mean; pr[kSeC + c] = mx; pr[2 * kSeC + c] = sd;
        if (epi.pool_bf) {
          bf16* pq = (bf16*)epi.pool_bf + (size_t)(b0 + j) * 3 * kSeC;
          pq[c] = __float2bfloat16_rn(mean); pq[kSeC + c] = __float2bfloat16_rn(mx); pq[2 * kSeC + c] = __float2bfloat16_rn(sd);
        }
      }
    }
  };
  tmem_pipeline<kChunks>(taddr, [&](const uint32_t(&r)[16], int chunk) {
    if (chunk + kPf < kChunks) prefetch(chunk + kPf);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const int col = chunk * 16 + i;
      if (col < kBoards * 81) {
        const int j0 = col / 81, j1 = (col + 1) / 81;
        const float v0 = fmaf(__uint_as_float(r[i]), R.sc, R.sh);
        const float o0 = fmaxf(fmaf(v0, sig[j0], shf[j0]) + __uint_as_float((uint32_t)pf[chunk % (kPf + 1)][i] << 16), 0.f);
        float o1 = 0.f;
        if (col + 1 < kBoards * 81) {
          const float v1 = fmaf(__uint_as_float(r[i + 1]), R.sc, R.sh);
          o1 = fmaxf(fmaf(v1, sig[j1], shf[j1]) + __uint_as_float((uint32_t)pf[chunk % (kPf + 1)][i + 1] << 16), 0.f);
        }
        const uint32_t pk = pack_bf16x2(o0, o1);
        finish_value(col, pk & 0xffffu);
        if (col + 1 < kBoards * 81) finish_value(col + 1, pk >> 16);
      }
    }
  });
}

// ---------------------------------------------------------------- forward / dgrad kernel, CTA pair (cta_group::2)
// The two SMs of a TPC compute ONE 256-channel x 3-board tile with M = 256 UMMAs: each CTA stages its own 128 weight
// rows (A, 16 KB per K block) and HALF of the pixel columns (B, 128 rows = 16 KB), the tensor cores of both SMs read
// both halves. Per stage a CTA holds 32 KB instead of 47 KB -> 6 stages in the same 192 KB, and the L2 -> SM operand
// bytes per FLOP drop by a third (the single-CTA kernel's MMA issuer stalls on the full barriers, profiles/).
// The 256-column tile is split at column 128, which is NOT a board boundary (81 pixels per board): CTA 0 takes board 0,
// rows 0-4 of board 1 and the first 2 pixels of its row 5; CTA 1 the other 7 pixels of that row, rows 6-8 and board 2
// (115 columns; the last 13 are never written and never read back) — five TMA box shapes, each shifted by the tap.
// One thread of the leader CTA issues every MMA; both producers credit the leader's full barrier; tcgen05.commit
// multicasts the stage release / accumulator-ready arrivals to both CTAs; the epilogue warps of both CTAs (each owns
// its CTA's 128 TMEM lanes = output channels) release the accumulator on the leader's barrier.
constexpr int kBHalfBytes = 128 * kBlockK * 2;          // 16 KB
constexpr int kStageBytes2 = kABytes + kBHalfBytes;     // 32 KB
// Warp roles of the pair kernel (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue as
// TWO groups of four warps that take alternate tiles (group g always drains accumulator buffer g). One group alone is the
// pacing stage of the single-CTA kernel: with one epilogue warp per scheduler every dependent instruction exposes its
// latency (ncu: 4.5-6 cycles per issued instruction), and the epilogue of a tile takes about as long as its MMAs. Two
// groups give each tile's epilogue two MMA periods and every scheduler two warps to interleave.
constexpr int kThreads2 = 320;
template <bool kSeTail> struct PairCfg {
  static constexpr int kStages = 6;
  static constexpr int kRingBytes = kStages * kStageBytes2;
  static constexpr int kSmemBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + (kSeTail ? (int)sizeof(SeTailSmem) : 0);
};
constexpr uint32_t kTxBytes2 = 2u * kABytes + 128u * 128u + 115u * 128u;   // both CTAs' boxes (OOB-filled elements count)
constexpr uint32_t kIdescF16M256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// Two-board ("narrow") tiles, N = 176: CTA 0 holds board 0 and the first 7 pixels of board 1 (88 columns), CTA 1 the
// other 74 pixels (2 of row 0 + rows 1-8) and 14 columns nobody reads.
constexpr int kNarrowN = 176;
constexpr uint32_t kTxBytes2Narrow = 2u * kABytes + 88u * 128u + 74u * 128u;
constexpr uint32_t kIdescF16M256Narrow = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNarrowN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct PairMaps { CUtensorMap full, rows5, x2, x7, rows3, rows8; };

// Tile schedule of the pair kernel. The boards are dealt to the CTA pairs as contiguous, near-equal ranges (the pairs are
// persistent: the kernel ends with the slowest one), and a range is cut into three-board tiles plus zero, one or two
// two-board tiles (N = 176 instead of 256: 0.69 of the MMA time), so a range of 7 boards costs 256 + 176 + 176 columns
// instead of three full tiles. 512 boards on 74 pairs: 608 columns for the slowest pair instead of 768; 4096 boards: 4784
// instead of 4864. All three warp roles walk the same sequence.
struct PairSched {
  int b, end, cp, n_pairs;
  __device__ __forceinline__ PairSched(int B, int n_pairs_) : cp(0), n_pairs(n_pairs_) {
    const int P = (int)(gridDim.x >> 1), p = (int)(blockIdx.x >> 1);
    const int base = B / P, rem = B % P;
    b = p * base + min(p, rem);
    end = b + base + (p < rem ? 1 : 0);
  }
  __device__ __forceinline__ bool done() const { return b >= end; }
  __device__ __forceinline__ int boards() const {          // boards of the tile starting at b: 3, or 2 / 1 at the tail
    const int left = end - b;
    return (left == 1 || left == 2 || left == 4) ? min(left, 2) : 3;
  }
  __device__ __forceinline__ void next() { if (++cp == n_pairs) { cp = 0; b += boards(); } }
};

template <int F, bool kSeTail>
__global__ void __launch_bounds__(kThreads2, 1)   // 10 warps are allocated as 12 (groups of 4): 168 registers per thread at most
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ PairMaps mx,
                   bf16* __restrict__ out, int B, int Cin, int Cout, int num_groups, ConvEpi epi) {
  constexpr int kStages2 = PairCfg<kSeTail>::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages2 * kStageBytes2;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages2 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kStages2 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kStages2 + 2 + b); };
  const uint32_t holder = bar_base + 8u * (2 * kStages2 + 4);
  auto se_bar = [&](int g) { return bar_base + 8u * (2 * kStages2 + 5 + g); };   // means exchange per epilogue group (fused tail)
  const uint32_t se_sm_addr = bar_base + 256u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages2 * kStageBytes2 + 8 * (2 * kStages2 + 4));
  SeTailSmem* se_sm = reinterpret_cast<SeTailSmem*>(smem_gen + kStages2 * kStageBytes2 + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tap = Cin / kBlockK;
  const int num_kb = 9 * kb_per_tap;
  const uint32_t crank = cluster_ctarank();
  const int n_pairs = Cout / 256;                       // channel pairs per board tile
  (void)num_groups;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }   // 4 epilogue warps x 2 CTAs
    if (kSeTail) { mbar_init(se_bar(0), 128); mbar_init(se_bar(1), 128); }   // the 128 threads of a group (+ the peer's bytes)
    fence_barrier_init();
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&mx.full);
  }
  if (warp == 1) {  // both CTAs of the pair allocate together: all 512 columns (two 256-column accumulators) in each
    tmem_alloc_2cta(holder, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled at them
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): the warp runs the loop uniformly, one elected lane issues =====
    const bool issuer = elect_one_sync();
    int stage = 0; uint32_t phase = 0;
    for (PairSched ts(B, n_pairs); !ts.done(); ts.next()) {
      const int ct = ts.cp * 2 + (int)crank;
      const int b0 = ts.b;
      const bool wide = ts.boards() == kBoards;
      int tap = 0, cc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sx = tap % 3 - 1, sy = tap / 3 - 1, c0 = cc * kBlockK;
        mbar_wait(empty_bar(stage), phase ^ 1u);          // this CTA's copy of the stage has been consumed by the pair's MMAs
        const uint32_t a_dst = smem_base + stage * kStageBytes2;
        const uint32_t b_dst = a_dst + kABytes;
        if (issuer) {
          if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), wide ? kTxBytes2 : kTxBytes2Narrow);
          tma_load_2d_2cta(a_dst, &map_w, full_bar(stage), kb * kBlockK, ct * kTileM);
          if (!wide) {
            if (crank == 0) {
              tma_load_4d_2cta(b_dst, &mx.full, full_bar(stage), c0, sx, sy, b0);                     // board 0: columns 0..80
              tma_load_4d_2cta(b_dst + 81 * 128, &mx.x7, full_bar(stage), c0, sx, sy, b0 + 1);        // board 1, row 0, x = 0..6: 81..87
            } else {
              tma_load_4d_2cta(b_dst, &mx.x2, full_bar(stage), c0, 7 + sx, sy, b0 + 1);               // row 0, x = 7..8: 88..89
              tma_load_4d_2cta(b_dst + 2 * 128, &mx.rows8, full_bar(stage), c0, sx, 1 + sy, b0 + 1);  // rows 1..8: 90..161
            }
          } else if (crank == 0) {
            tma_load_4d_2cta(b_dst, &mx.full, full_bar(stage), c0, sx, sy, b0);                       // board 0: columns 0..80
            tma_load_4d_2cta(b_dst + 81 * 128, &mx.rows5, full_bar(stage), c0, sx, sy, b0 + 1);       // board 1 rows 0..4: 81..125
            tma_load_4d_2cta(b_dst + 126 * 128, &mx.x2, full_bar(stage), c0, sx, 5 + sy, b0 + 1);     // row 5, x = 0..1: 126..127
          } else {
            tma_load_4d_2cta(b_dst, &mx.x7, full_bar(stage), c0, 2 + sx, 5 + sy, b0 + 1);             // row 5, x = 2..8: 128..134
            tma_load_4d_2cta(b_dst + 7 * 128, &mx.rows3, full_bar(stage), c0, sx, 6 + sy, b0 + 1);    // rows 6..8: 135..161
            tma_load_4d_2cta(b_dst + 34 * 128, &mx.full, full_bar(stage), c0, sx, sy, b0 + 2);        // board 2: 162..242
          }
        }
        if (++cc == kb_per_tap) { cc = 0; ++tap; }
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one elected lane of the leader CTA's warp 1 for the pair (uniform loop) =====
    if (crank == 0) {
      const bool issuer = elect_one_sync();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (PairSched ts(B, n_pairs); !ts.done(); ts.next(), ++it) {
        const uint32_t idesc = ts.boards() == kBoards ? kIdescF16M256 : kIdescF16M256Narrow;
        const int buf = it & 1;
        const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), tphase ^ 1u);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kTileN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);      // both CTAs' boxes of this stage have landed
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * kStageBytes2;
          const uint64_t adesc = smem_desc_k128(a_addr);
          const uint64_t bdesc = smem_desc_k128(a_addr + kABytes);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_2cta(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta_mc(empty_bar(stage), (uint16_t)3);   // frees the stage in both CTAs when these MMAs retire
            if (kb == num_kb - 1) umma_commit_2cta_mc(tfull_bar(buf), (uint16_t)3);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): thread = output channel of this CTA's half; group g takes tiles it = g, g + 2, ... =====
    const int lane_grp = warp & 3;        // TMEM lanes 32*lane_grp .. +31 are the ones this warp may read
    const int g = (warp - 2) >> 2;        // epilogue group 0 / 1 = accumulator buffer
    SeTailRegs se_regs;
    if (kSeTail) se_tail_load(epi, (int)crank * kTileM + lane_grp * 32 + lane, lane_grp, lane, se_regs);
    int it = 0;
    for (PairSched ts(B, n_pairs); !ts.done(); ts.next(), ++it) {
      if ((it & 1) != g) continue;
      const int ct = ts.cp * 2 + (int)crank;
      const int buf = g;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int c = ct * kTileM + lane_grp * 32 + lane;
      const int b0 = ts.b;
      const int nb_valid = ts.boards();
      if (kSeTail) {
        // L2 prefetch of the whole residual tile of this warp (243 rows x 64 bytes = its 32 channels), 8 requests per
        // lane, issued BEFORE waiting for the tile's MMAs: when pass 2 asks for them the register prefetches find their
        // lines in L2 instead of waiting on HBM (ncu: the residual use was the fused epilogue's top stall)
        const char* wbase = reinterpret_cast<const char*>(epi.res) + (((size_t)b0 * 81) * kSeC + (size_t)(c - lane)) * 2;
#pragma unroll
        for (int k = 0; k < (kBoards * 81 + 31) / 32; ++k) {
          const int col = k * 32 + lane;
          if (col < kBoards * 81 && col / 81 < nb_valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(wbase + (size_t)col * kSeC * 2));
        }
      }
      mbar_wait(tfull_bar(buf), tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * kTileN);
      if (kSeTail) {
        if (nb_valid == kBoards)
          se_tail_tile<true>(epi, se_regs, se_sm, se_sm_addr, se_bar(g), crank, g, tphase, taddr, c, lane_grp, lane, b0, nb_valid, out);
        else
          se_tail_tile<false>(epi, se_regs, se_sm, se_sm_addr, se_bar(g), crank, g, tphase, taddr, c, lane_grp, lane, b0, nb_valid, out);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(buf));
      } else {
        ConvEpiThread<bf16, kBoards, F> et(epi, c, Cout, B);
        epilogue_tile<F>(epi, et, taddr, b0, nb_valid, Cout, out);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(buf));
        et.finish(nb_valid);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's operands / signal its barriers until it is done too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight-gradient kernel
// dW[co][ci][tap] = sum_{board, pixel} dY[board][pixel][co] * X[board][pixel + shift(tap)][ci]
//
// GEMM view per CTA: one (tap, 128-channel half of Cout) unit and one slice of the boards.
//   D[co (128 TMEM lanes), ci (N = Cin columns)] += A[co, k] * B[ci, k],  k = pixels.
// Both operands are "MN-major": in NHWC memory the channel index is contiguous for a fixed pixel k.
// K is a multiple of UMMA_K = 16 but a board has 81 pixels. Rather than padding every board to 96 rows (15.6 % of the
// tensor work multiplying zeros), a board stage holds pixels 0..79 (two TMA boxes per 64-channel group: 9 x 8 rows and the
// first 8 pixels of the last row; 80 rows x 128 B, swizzle 128B, both landing on 1024-byte boundaries) = 5 UMMA_K steps,
// and the 81st pixel (8, 8) of SIXTEEN consecutive boards arrives as one extra box (1 x 1 x 16 boards = 16 rows) = one
// more step per 16 boards: 81 steps per 16 boards... per board 5 + 1/16, no zero work. The "extras" item travels through
// the same stage ring as the boards. Slices are multiples of 16 boards; boards past B are zero-filled by TMA.
// The tap shift is the TMA box origin on X (zero fill at the board edge), exactly as in the forward.
// Partial tiles go to a workspace with coalesced stores; wgrad_reduce_kernel sums the board slices
// and scatters into the PyTorch (Cout, Cin, 3, 3) fp32 gradient.
constexpr int kWgStages = 3;
constexpr int kWgRows = 80;                          // pixels 0..79 of one board
constexpr int kWgGroupBytes = kWgRows * 128;         // one 64-channel group of one board: 10 KB (a multiple of 1024)
constexpr int kWgMainBytes = 72 * 128;               // box 9 x 8
constexpr int kWgRowBytes = 8 * 128;                 // box 8 x 1, lands at row 72
constexpr int kWgExtraBoards = 16;                   // boards whose last pixel shares one UMMA_K step
constexpr int kWgExtraBytes = kWgExtraBoards * 128;
constexpr int kWgAGroups = 2;                        // 128 output channels
constexpr int kWgMaxBGroups = 4;                     // up to 256 input channels
constexpr int kWgStageBytes = (kWgAGroups + kWgMaxBGroups) * kWgGroupBytes;  // 60 KB
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 1024 + 256;

// MN-major operand, SWIZZLE_128B: 64-element groups LBO apart, 8-row (k) groups SBO = 1024 B apart
struct WgMaps { CUtensorMap main, row, extra; };     // boxes (64c, 9, 8, 1), (64c, 8, 1, 1), (64c, 1, 1, 16)

__device__ __forceinline__ uint64_t smem_desc_mn128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(kWgGroupBytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ WgMaps map_dy, const __grid_constant__ WgMaps map_x,
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

  const int nblocks = (nboards + kWgExtraBoards - 1) / kWgExtraBoards;   // items: per block its boards, then one extras item
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dy.main); tma_prefetch_desc(&map_dy.row); tma_prefetch_desc(&map_dy.extra);
    tma_prefetch_desc(&map_x.main); tma_prefetch_desc(&map_x.row); tma_prefetch_desc(&map_x.extra);
  }
  if (warp == 1) {
    tmem_alloc(holder, 256);
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
    const int sx = tap % 3 - 1, sy = tap / 3 - 1;
    int stage = 0; uint32_t phase = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int b0 = b_begin + blk * kWgExtraBoards;
      const int nb = min(kWgExtraBoards, b_end - b0);
      for (int i = 0; i <= nb; ++i) {               // i == nb: the extras item of this block
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * kWgStageBytes;
        if (issuer) {
          if (i < nb) {
            const int b = b0 + i;
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)((kWgAGroups + b_groups) * kWgGroupBytes));
#pragma unroll
            for (int g = 0; g < kWgAGroups; ++g) {
              tma_load_4d(a_dst + g * kWgGroupBytes, &map_dy.main, full_bar(stage), half * kTileM + g * 64, 0, 0, b);
              tma_load_4d(a_dst + g * kWgGroupBytes + kWgMainBytes, &map_dy.row, full_bar(stage), half * kTileM + g * 64, 0, 8, b);
            }
            for (int g = 0; g < b_groups; ++g) {
              tma_load_4d(a_dst + (kWgAGroups + g) * kWgGroupBytes, &map_x.main, full_bar(stage), g * 64, sx, sy, b);
              tma_load_4d(a_dst + (kWgAGroups + g) * kWgGroupBytes + kWgMainBytes, &map_x.row, full_bar(stage), g * 64, sx, sy + 8, b);
            }
          } else {
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)((kWgAGroups + b_groups) * kWgExtraBytes));
#pragma unroll
            for (int g = 0; g < kWgAGroups; ++g)
              tma_load_4d(a_dst + g * kWgGroupBytes, &map_dy.extra, full_bar(stage), half * kTileM + g * 64, 8, 8, b0);
            for (int g = 0; g < b_groups; ++g)
              tma_load_4d(a_dst + (kWgAGroups + g) * kWgGroupBytes, &map_x.extra, full_bar(stage), g * 64, sx + 8, sy + 8, b0);
          }
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool issuer = elect_one_sync();
    // M = 128, N = Cin, A and B MN-major (bits 15, 16), fp32 accumulate, bf16 inputs
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Cin >> 3) << 17) |
                           ((uint32_t)(kTileM >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int nb = min(kWgExtraBoards, nboards - blk * kWgExtraBoards);
      for (int i = 0; i <= nb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * kWgStageBytes;
        const uint64_t adesc = smem_desc_mn128(a_addr);
        const uint64_t bdesc = smem_desc_mn128(a_addr + kWgAGroups * kWgGroupBytes);
        if (issuer) {
          const int ksteps = i < nb ? kWgRows / 16 : 1;     // a board: 80 pixels; the extras item: 16 last pixels
          for (int k = 0; k < ksteps; ++k) {
            // 16 pixel rows = 2048 bytes further down every group: +128 in the (addr >> 4) field
            umma_bf16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, accumulate);
            accumulate = 1u;
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
    if (issuer) umma_commit(done_bar);
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

// ---------------------------------------------------------------- weight-gradient kernel, CTA pair (cta_group::2)
// The single-CTA kernel above moves 6 operand groups (60 KB) per board into every SM for 5 UMMA_K steps: ~81 B/clk per
// SM, which is what the L2 -> shared-memory path delivers (measured: cutting the tensor work by 15.6 % bought 3.5 %).
// A CTA pair computes the whole 256 x Cin gradient tile of one tap with M = 256 UMMAs: each CTA loads only ITS 128 output
// channels of dY and ITS half of the input channels of X (4 groups, 40 KB per board), the tensor cores of both SMs read
// the B halves from both shared memories. Cluster = (tap, board slice); rank = Cout half = the workspace block the
// reduce kernel expects. Same item stream (boards of 80 pixels + one extras item per 16 boards), 5-stage ring.
constexpr int kWg2Stages = 5;
constexpr int kWg2Groups = 4;                                   // per CTA: 2 of dY + up to 2 of X
constexpr int kWg2StageBytes = kWg2Groups * kWgGroupBytes;      // 40 KB
constexpr int kWg2SmemBytes = kWg2Stages * kWg2StageBytes + 1024 + 256;

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_tc2_kernel(const __grid_constant__ WgMaps map_dy, const __grid_constant__ WgMaps map_x,
                         float* __restrict__ ws, int B, int Cin, int boards_per_slice) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kWg2Stages * kWg2StageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWg2Stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWg2Stages);
  const uint32_t holder = bar_base + 8u * (2 * kWg2Stages + 1);
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kWg2Stages * kWg2StageBytes + 8 * (2 * kWg2Stages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = (int)(blockIdx.x >> 1);                 // cluster index = slice * 9 + tap
  const int tap = cl % 9, slice = cl / 9;
  const int bg = Cin / 128;                              // X groups per CTA (its half of the input channels)
  const int b_begin = slice * boards_per_slice;
  const int b_end = min(B, b_begin + boards_per_slice);
  const int nboards = b_end - b_begin;
  const int nblocks = (nboards + kWgExtraBoards - 1) / kWgExtraBoards;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWg2Stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dy.main); tma_prefetch_desc(&map_dy.row); tma_prefetch_desc(&map_dy.extra);
    tma_prefetch_desc(&map_x.main); tma_prefetch_desc(&map_x.row); tma_prefetch_desc(&map_x.extra);
  }
  if (warp == 1) {
    tmem_alloc_2cta(holder, 256);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled at them
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // TMA producer (both CTAs): data lands in the issuing CTA, the bytes are credited to the leader's full barrier
    const bool issuer = elect_one_sync();
    const int sx = tap % 3 - 1, sy = tap / 3 - 1;
    const int co0 = (int)crank * kTileM, ci0 = (int)crank * bg * 64;
    int stage = 0; uint32_t phase = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int b0 = b_begin + blk * kWgExtraBoards;
      const int nb = min(kWgExtraBoards, b_end - b0);
      for (int i = 0; i <= nb; ++i) {               // i == nb: the extras item of this block
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * kWg2StageBytes;
        if (issuer) {
          if (i < nb) {
            const int b = b0 + i;
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(2 * (kWgAGroups + bg) * kWgGroupBytes));
#pragma unroll
            for (int g = 0; g < kWgAGroups; ++g) {
              tma_load_4d_2cta(a_dst + g * kWgGroupBytes, &map_dy.main, full_bar(stage), co0 + g * 64, 0, 0, b);
              tma_load_4d_2cta(a_dst + g * kWgGroupBytes + kWgMainBytes, &map_dy.row, full_bar(stage), co0 + g * 64, 0, 8, b);
            }
            for (int g = 0; g < bg; ++g) {
              tma_load_4d_2cta(a_dst + (kWgAGroups + g) * kWgGroupBytes, &map_x.main, full_bar(stage), ci0 + g * 64, sx, sy, b);
              tma_load_4d_2cta(a_dst + (kWgAGroups + g) * kWgGroupBytes + kWgMainBytes, &map_x.row, full_bar(stage), ci0 + g * 64, sx, sy + 8, b);
            }
          } else {
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(2 * (kWgAGroups + bg) * kWgExtraBytes));
#pragma unroll
            for (int g = 0; g < kWgAGroups; ++g)
              tma_load_4d_2cta(a_dst + g * kWgGroupBytes, &map_dy.extra, full_bar(stage), co0 + g * 64, 8, 8, b0);
            for (int g = 0; g < bg; ++g)
              tma_load_4d_2cta(a_dst + (kWgAGroups + g) * kWgGroupBytes, &map_x.extra, full_bar(stage), ci0 + g * 64, sx + 8, sy + 8, b0);
          }
        }
        if (++stage == kWg2Stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {   // one elected lane of the leader issues the pair's MMAs
      const bool issuer = elect_one_sync();
      // M = 256 over the pair, N = Cin, A and B MN-major (bits 15, 16), fp32 accumulate, bf16 inputs
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Cin >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int blk = 0; blk < nblocks; ++blk) {
        const int nb = min(kWgExtraBoards, nboards - blk * kWgExtraBoards);
        for (int i = 0; i <= nb; ++i) {
          mbar_wait(full_bar(stage), phase);        // both CTAs' boxes of this item have landed
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * kWg2StageBytes;
          const uint64_t adesc = smem_desc_mn128(a_addr);
          const uint64_t bdesc = smem_desc_mn128(a_addr + kWgAGroups * kWgGroupBytes);
          if (issuer) {
            const int ksteps = i < nb ? kWgRows / 16 : 1;
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16_2cta(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, accumulate);
              accumulate = 1u;
            }
            umma_commit_2cta_mc(empty_bar(stage), (uint16_t)3);   // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == kWg2Stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (issuer) umma_commit_2cta_mc(done_bar, (uint16_t)3);
    }
  } else {
    const int lane_grp = warp & 3;
    // ws[slice * 18 + tap * 2 + half][ci][co_local]: the layout wgrad_reduce_kernel sums
    float* dst = ws + ((size_t)(slice * 18 + tap * 2 + (int)crank) * Cin) * kTileM + lane_grp * 32 + lane;
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
  cluster_sync_all();  // the peer may still read this CTA's operands / signal its barriers until it is done too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 256);
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
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (conv_priority() != 0) { attr[n].id = cudaLaunchAttributePriority; attr[n].val.priority = conv_priority(); ++n; }
  if (kb_pdl_enabled()) { attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  cfg.attrs = attr; cfg.numAttrs = n;
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

int make_act_box_map(CUtensorMap* m, const void* x, int B, int C, int bx, int by, int bb) {
  EncodeTiledFn enc = get_encode_fn();
  KB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)C, 9, 9, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 9, (cuuint64_t)C * 2 * 81};
  const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bb};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation box %dx%dx%d) failed: %d", bx, by, bb, (int)r);
  return KB_OK;
}

template <int F, bool kSeTail = false>
int launch_pair(const CUtensorMap& mw, const PairMaps& mx, bf16* out, int B, int Cin, int Cout, int groups,
                const ConvEpi& epi, int grid, cudaStream_t st) {
  constexpr int smem = PairCfg<kSeTail>::kSmemBytes;
  static bool attr_set = false;  // per instantiation; idempotent
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc2_kernel<F, kSeTail>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[3];
  int n = 0;
  at[n].id = cudaLaunchAttributeClusterDimension;
  at[n].val.clusterDim.x = 2; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n;
  if (conv_priority() != 0) { at[n].id = cudaLaunchAttributePriority; at[n].val.priority = conv_priority(); ++n; }
  if (kb_pdl_enabled()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  cfg.attrs = at; cfg.numAttrs = n;
  KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<F, kSeTail>, mw, mx, out, B, Cin, Cout, groups, epi));
  kb_count_launch();
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

int kbk_conv3x3_se_tail_supported(int Cin, int Cout, int S, int dtype) {
  return dtype == KB_BF16 && Cout == kSeC && S == kSeS && Cin % kBlockK == 0 && Cin >= kBlockK;
}

int kbk_conv3x3_tc_supported(int Cin, int Cout, int dtype) {
  return dtype == KB_BF16 && Cin % kBlockK == 0 && Cout % kTileM == 0 && Cin >= kBlockK;
}

// mode: 0 = automatic (CTA-pair kernel whenever Cout is a multiple of 256; KB_CONV_2CTA=0 disables), 1 = single-CTA
// kernel, 2 = CTA-pair kernel (error if the shape does not allow it)
int kbk_conv3x3_tc_mode(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi, int num_sms,
                        int mode, cudaStream_t st) {
  KB_CHECK_ARG(kbk_conv3x3_tc_supported(Cin, Cout, KB_BF16), "conv3x3_tc: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  if (B == 0) return KB_OK;
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("KB_CONV_2CTA"); pair_env = (e && e[0] == '0') ? 0 : 1; }
  const bool pair_ok = Cout % 256 == 0;
  KB_CHECK_ARG(mode != 2 || pair_ok, "conv3x3_tc: the CTA-pair kernel needs Cout %% 256 == 0 (got %d)", Cout);
  if (mode == 2 || (mode == 0 && pair_env && pair_ok)) {
    CUtensorMap mw;
    PairMaps pm;
    if (int r = make_weight_map(&mw, w, Cout, 9 * Cin)) return r;
    if (int r = make_act_box_map(&pm.full, in, B, Cin, 9, 9, 1)) return r;
    if (int r = make_act_box_map(&pm.rows5, in, B, Cin, 9, 5, 1)) return r;
    if (int r = make_act_box_map(&pm.x2, in, B, Cin, 2, 1, 1)) return r;
    if (int r = make_act_box_map(&pm.x7, in, B, Cin, 7, 1, 1)) return r;
    if (int r = make_act_box_map(&pm.rows3, in, B, Cin, 9, 3, 1)) return r;
    if (int r = make_act_box_map(&pm.rows8, in, B, Cin, 9, 8, 1)) return r;
    const int groups = kb_ceil_div(B, kBoards);
    if (num_sms <= 0) num_sms = 148;
    // one contiguous board range per CTA pair (PairSched); with fewer boards than pairs every pair gets one board
    const int grid = 2 * (B < num_sms / 2 ? B : num_sms / 2);
    bf16* o = (bf16*)out;
    if (epi.res != nullptr) {   // fused evaluation tail
      KB_CHECK_ARG(Cout == kSeC && epi.scale && epi.shift && epi.se_w1 && epi.se_b1 && epi.se_w2 && epi.se_b2 && epi.pool,
                   "conv3x3_tc: the fused SE tail needs Cout == 256, folded BatchNorm and all four SE tensors");
      return launch_pair<0, true>(mw, pm, o, B, Cin, Cout, groups, epi, grid, st);
    }
    const int f = conv_epi_features(epi);
#define KB_PAIR(FF) return launch_pair<FF>(mw, pm, o, B, Cin, Cout, groups, epi, grid, st)
    switch (f) {
      case 0: KB_PAIR(0);
      case kEpiSum | kEpiSumSq: KB_PAIR(kEpiSum | kEpiSumSq);
      case kEpiSum | kEpiSumSq | kEpiBoard: KB_PAIR(kEpiSum | kEpiSumSq | kEpiBoard);
      case kEpiAffine | kEpiRelu | kEpiGbias: KB_PAIR(kEpiAffine | kEpiRelu | kEpiGbias);
      case kEpiAffine | kEpiRelu: KB_PAIR(kEpiAffine | kEpiRelu);
      case kEpiAffine: KB_PAIR(kEpiAffine);
      case kEpiAffine | kEpiRelu | kEpiPool: KB_PAIR(kEpiAffine | kEpiRelu | kEpiPool);
      case kEpiAffine | kEpiBoard: KB_PAIR(kEpiAffine | kEpiBoard);
      case kEpiMask | kEpiSum | kEpiDot | kEpiBoard: KB_PAIR(kEpiMask | kEpiSum | kEpiDot | kEpiBoard);
      default: break;   // rarely used feature sets stay on the single-CTA kernel
    }
#undef KB_PAIR
    KB_CHECK_ARG(mode != 2, "conv3x3_tc: epilogue feature set %d is not instantiated for the CTA-pair kernel", f);
  }
  return kbk_conv3x3_tc_single(in, w, out, B, Cin, Cout, epi, num_sms, st);
}

int kbk_conv3x3_tc(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi, int num_sms,
                   cudaStream_t st) {
  return kbk_conv3x3_tc_mode(in, w, out, B, Cin, Cout, epi, num_sms, 0, st);
}

int kbk_conv3x3_tc_single(const void* in, const void* w, void* out, int B, int Cin, int Cout, const ConvEpi& epi, int num_sms,
                          cudaStream_t st) {
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
  // whole 16-board blocks per slice: the 81st pixels of a block share one TMA box and one UMMA_K step
  const int bps = kb_ceil_div(kb_ceil_div(B, slices), kWgExtraBoards) * kWgExtraBoards;
  slices = kb_ceil_div(B, bps);
  KB_CHECK_ARG(ws != nullptr && ws_bytes >= (long long)slices * units * Cin * kTileM * (long long)sizeof(float),
               "conv3x3_wgrad_tc: workspace too small");
  WgMaps mdy, mx;
  if (int r = make_act_box_map(&mdy.main, dy, B, Cout, 9, 8, 1)) return r;
  if (int r = make_act_box_map(&mdy.row, dy, B, Cout, 8, 1, 1)) return r;
  if (int r = make_act_box_map(&mdy.extra, dy, B, Cout, 1, 1, kWgExtraBoards)) return r;
  if (int r = make_act_box_map(&mx.main, x, B, Cin, 9, 8, 1)) return r;
  if (int r = make_act_box_map(&mx.row, x, B, Cin, 8, 1, 1)) return r;
  if (int r = make_act_box_map(&mx.extra, x, B, Cin, 1, 1, kWgExtraBoards)) return r;
  static const bool pair_off = [] { const char* e = getenv("KB_WGRAD_2CTA"); return e && e[0] == '0'; }();
  if (!pair_off && Cout == 256 && (Cin == 128 || Cin == 256)) {
    // CTA-pair kernel: cluster = (tap, slice), the two ranks are the two Cout halves -> the same 18 workspace blocks per slice
    static bool attr2_set = false;
    if (!attr2_set) {
      KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWg2SmemBytes));
      attr2_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(units * slices)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kWg2SmemBytes; cfg.stream = st;
    cudaLaunchAttribute at[3];
    int n = 0;
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 2; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n;
    if (conv_priority() != 0) { at[n].id = cudaLaunchAttributePriority; at[n].val.priority = conv_priority(); ++n; }
    if (kb_pdl_enabled()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv3x3_wgrad_tc2_kernel, mdy, mx, ws, B, Cin, bps));
    kb_count_launch();
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      KB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
      attr_set = true;
    }
    launch_prio(conv3x3_wgrad_tc_kernel, units * slices, kThreads, kWgSmemBytes, st, mdy, mx, ws, B, Cin, Cout, units, bps);
    KB_CUDA_LAUNCH_CHECK();
  }
  wgrad_reduce_kernel<<<dim3(kb_ceil_div(Cout, 32), kb_ceil_div(Cin_true, kRedCi)), 256, 0, st>>>(ws, dw, Cin, Cout, Cin_true, units, slices);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
