// se_apply.cu — the tail of a GlobalPoolBiasBlock as ONE HBM-bound kernel (bf16 activations):
//
//   se_in  = mean_p(bn2(z2))                         (from the conv2 epilogue's board means)
//   se     = W2 relu(W1 se_in + b1) + b2             squeeze-excite MLP      (se_resnet.py:83-86)
//   x'     = relu(bn2(z2) * sigmoid(scale) + shift + x)                      (se_resnet.py:87-90)
//   pool'  = (mean, max, population std) of x' per (board, channel)          (se_resnet.py:93-98, next block's input)
//
// It replaces affine_rows + two Linear launches + apply_vec_kernel. Design for B200: persistent CTAs
// (one per SM, 512 threads) walk boards; the two 41 KB input tiles of a board (z2, x; NHWC rows are
// contiguous) are staged with 1-D TMA bulk copies into a 2-stage shared-memory ring signalled by mbarriers,
// so ~83 KB per SM is always in flight with no registers spent on it (the register-staged kernel it replaces
// ran at 44 % of HBM peak, limited by 24 KB in flight per SM). The SE weights live in shared memory for the
// whole kernel; the MLP of the next board overlaps the bulk copies. Output rows are written with 16-byte
// coalesced stores; per-(board, channel) statistics use a common shift (the pixel-0 value) so partial sums of
// the 16 pixel lanes merge by plain addition, through the drained stage buffer as scratch.
#include <stdlib.h>
#include "kb_common.cuh"
#include "kb_kernels.h"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kThreads = 512;
constexpr int kGroupThreads = 256;
constexpr int kQuant = 4;  // ds, dss, max, ties

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float round_bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Layout {
  size_t tile, stage, off_w1, off_w2t, off_b1, off_b2, off_hid, off_k0, off_bar, off_scratch, total;
  bool scratch_aliases_stage;
};
// Shared memory: [stage 0: z, x][stage 1: z, x][SE weights (prologue) / reduction scratch (main loop)][small vectors]
__host__ __device__ inline Layout make_layout(int C, int S, bool ties) {
  Layout L;
  L.tile = (size_t)81 * C * 2;
  L.stage = 2 * L.tile;
  size_t o = 2 * L.stage;
  auto take = [&](size_t bytes) { const size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
  const size_t wbytes = (size_t)S * C * 4 + (size_t)S * 2 * C * 4;
  const size_t scratch = (size_t)(ties ? 4 : 3) * kThreads * 8 * 4;  // both groups: quantities x (pixel lanes * C) floats
  // the weights are dead once the prologue is over: the scratch overlays them when it fits in the budget,
  // otherwise it aliases the drained stage (one more barrier per board)
  const size_t budget = 227 * 1024 - 2 * L.stage - 8 * 1024;
  L.scratch_aliases_stage = scratch > budget && scratch / 2 <= L.stage;
  const size_t region = L.scratch_aliases_stage ? wbytes : (wbytes > scratch ? wbytes : scratch);
  L.off_w1 = take(region);
  L.off_w2t = L.off_w1 + (size_t)S * C * 4;
  L.off_scratch = L.off_w1;
  L.off_b1 = take((size_t)S * 4);
  L.off_b2 = take((size_t)2 * C * 4);
  L.off_hid = take((size_t)(kThreads / 32) * S * 4);
  L.off_k0 = take((size_t)2 * C * 4);
  L.off_bar = take(16);
  L.total = o;
  return L;
}

// scratch layout: [quantity][pixel lane][half (channels 0-3 / 4-7 of the group)][channel group][4] floats, so the
// 16-byte stores of a warp are contiguous (conflict-free) and the combine pass reads consecutive words.
template <bool TIES>
__global__ void __launch_bounds__(kThreads, 1) se_apply_tma_kernel(SeApplyArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  // Two groups of 8 warps work on alternate boards (group = stage), each with its own named barrier, so the
  // waits of one group (tile landing, reductions) are covered by the other group's arithmetic.
  const int C = g.C, S = g.S, CG = C / 8, NPL = kGroupThreads / CG;
  const Layout L = make_layout(C, S, TIES);
  float* w1s = reinterpret_cast<float*>(smem + L.off_w1);    // [S][C]
  float* w2t = reinterpret_cast<float*>(smem + L.off_w2t);   // [S][2C]  (transposed: conflict-free per-output reads)
  float* b1s = reinterpret_cast<float*>(smem + L.off_b1);
  float* b2s = reinterpret_cast<float*>(smem + L.off_b2);
  float* hids = reinterpret_cast<float*>(smem + L.off_hid);  // [warp][S]
  float* k0s = reinterpret_cast<float*>(smem + L.off_k0);
  const uint32_t bar0 = smem_u32(smem + L.off_bar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid / kGroupThreads, gtid = tid % kGroupThreads;
  const int cg = gtid % CG, pl = gtid / CG, c0 = cg * 8;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(kGroupThreads) : "memory"); };

  const int nb = g.B > (int)blockIdx.x ? (g.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto board = [&](int i) { return (size_t)blockIdx.x + (size_t)i * gridDim.x; };
  auto issue = [&](int i) {  // thread 0: stage (i & 1) <- board i of this CTA
    const int s = i & 1;
    const size_t b = board(i);
    const uint32_t dst = smem_u32(smem + (size_t)s * L.stage), bar = bar0 + 8u * s;
    mbar_arrive_expect_tx(bar, (uint32_t)L.stage);
    bulk_load_1d(dst, g.z + b * 81 * C, (uint32_t)L.tile, bar);
    bulk_load_1d(dst + (uint32_t)L.tile, g.res + b * 81 * C, (uint32_t)L.tile, bar);
  };
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_barrier_init();
    if (nb > 0) issue(0);
    if (nb > 1) issue(1);
  }
  for (int i = tid; i < S * C; i += kThreads) w1s[i] = g.w1[i];
  for (int i = tid; i < 2 * C * S; i += kThreads) { const int h = i / (2 * C), n = i - h * 2 * C; w2t[i] = g.w2[n * S + h]; }
  for (int i = tid; i < S; i += kThreads) b1s[i] = g.b1[i];
  for (int i = tid; i < 2 * C; i += kThreads) b2s[i] = g.b2[i];
  float a_[8], b_[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a_[i] = g.a ? g.a[c0 + i] : 1.f; b_[i] = g.a ? g.b[c0 + i] : 0.f; }
  __syncthreads();

  // ---- prologue: the squeeze-excite MLP of every board of this CTA, one warp per board, no block barriers.
  //      Results go to global memory (se_out; L2-resident, 2 KB per board) while the first two tiles are in flight.
  {
    float* hw = hids + warp * S;
    const int NJ = C / 32;      // <= 8 input channels per lane
    const int NK = 2 * C / 32;  // <= 16 outputs per lane
    for (int i = warp; i < nb; i += kThreads / 32) {
      const size_t b = board(i);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        v[j] = 0.f;
        if (j < NJ) {
          const float m = g.bmean[b * C + c];
          v[j] = g.a ? fmaf(m, g.a[c], g.b[c]) : m;
          if (g.se_in_out) g.se_in_out[b * C + c] = v[j];
        }
      }
      for (int h = 0; h < S; ++h) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < NJ) d = fmaf(w1s[h * C + lane + 32 * j], v[j], d);
        d = fmaxf(kb_warp_sum(d) + b1s[h], 0.f);
        if (lane == 0) { hw[h] = d; if (g.seh_out) g.seh_out[b * S + h] = d; }
      }
      __syncwarp();
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = k < NK ? b2s[lane + 32 * k] : 0.f;
      for (int h = 0; h < S; ++h) {
        const float hv = hw[h];
#pragma unroll
        for (int k = 0; k < 16; ++k) if (k < NK) acc[k] = fmaf(w2t[h * 2 * C + lane + 32 * k], hv, acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < NK) {
          const int n = lane + 32 * k;
          // eval (se_raw == 0): the scale half is stored with the sigmoid already applied
          g.se_out[b * 2 * C + n] = (n < C && !g.se_raw) ? sigmoidf_(acc[k]) : acc[k];
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();  // se_out of this CTA's boards is visible to all its threads; the weight region is free

  const int lanes_c = NPL * C;
  float* scr_fixed = reinterpret_cast<float*>(smem + L.off_scratch) + (size_t)grp * (TIES ? 4 : 3) * lanes_c;
  k0s += grp * C;
  for (int i = grp; i < nb; i += 2) {
    const size_t b = board(i);
    const int s = i & 1;
    const uint32_t phase = (uint32_t)(i >> 1) & 1u;
    float sg[8], sf[8];
    {
      const float4* sp = reinterpret_cast<const float4*>(g.se_out + b * 2 * C + c0);
      const float4* hp = reinterpret_cast<const float4*>(g.se_out + b * 2 * C + C + c0);
      const float4 g0 = __ldcg(sp), g1 = __ldcg(sp + 1), h0 = __ldcg(hp), h1 = __ldcg(hp + 1);
      float sig[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (g.se_raw) sig[k] = sigmoidf_(sig[k]);
        sg[k] = a_[k] * sig[k];                  // (z*a+b)*sig + shift = z*(a*sig) + (b*sig + shift)
        sf[k] = fmaf(b_[k], sig[k], sh[k]);
      }
    }
    // ---- tile of this board ----
    mbar_wait(bar0 + 8u * s, phase);
    const uint8_t* zt = smem + (size_t)s * L.stage;
    const uint8_t* rt = zt + L.tile;
    // Statistics: sums on the fp32 (pre-rounding) outputs shifted by the pixel-0 value (a constant board gives an
    // exact zero variance); running max and tie counts on the STORED bf16 pairs with packed bf16x2 instructions
    // (max must be a stored value: the backward compares x == max).
    float k0[8], ds[8], dss[8];
    __nv_bfloat162 mx2[4], tie2[4];
    {
      float zv[8], rv[8];
      unpack8(*reinterpret_cast<const uint4*>(zt + (size_t)c0 * 2), zv);
      unpack8(*reinterpret_cast<const uint4*>(rt + (size_t)c0 * 2), rv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        k0[k] = fmaxf(fmaf(zv[k], sg[k], sf[k]) + rv[k], 0.f);  // pixel-0 output: the common shift of this channel
        ds[k] = 0.f; dss[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { mx2[k] = __float2bfloat162_rn(0.f); tie2[k] = __float2bfloat162_rn(0.f); }  // outputs are >= 0
    }
    bf16* orow = g.out + b * 81 * C + c0;
    const __nv_bfloat162 one2 = __float2bfloat162_rn(1.f);
#pragma unroll 2
    for (int p = pl; p < 81; p += NPL) {
      float zv[8], rv[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(zt + ((size_t)p * C + c0) * 2), zv);
      unpack8(*reinterpret_cast<const uint4*>(rt + ((size_t)p * C + c0) * 2), rv);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(zv[k], sg[k], sf[k]) + rv[k], 0.f);
      __nv_bfloat162 pk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pk[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
      *reinterpret_cast<uint4*>(orow + (size_t)p * C) = *reinterpret_cast<const uint4*>(pk);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = o[k] - k0[k];
        ds[k] += d; dss[k] = fmaf(d, d, dss[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 nm = __hmax2(mx2[k], pk[k]);
        if (TIES) {
          // count = count * [old max still the max] + [this value equals the new max]   (counts <= 6: exact in bf16)
          tie2[k] = __hfma2(tie2[k], __heq2(mx2[k], nm), __heq2(pk[k], nm));
        }
        mx2[k] = nm;
      }
    }
    float mx[8], tie[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mx[2 * k] = __low2float(mx2[k]); mx[2 * k + 1] = __high2float(mx2[k]);
      tie[2 * k] = __low2float(tie2[k]); tie[2 * k + 1] = __high2float(tie2[k]);
    }
    (void)one2;
    float* scr = scr_fixed;
    if (L.scratch_aliases_stage) {
      group_sync();  // every thread of the group is done reading its stage before it serves as reduction scratch
      scr = reinterpret_cast<float*>(smem + (size_t)s * L.stage);
    }
    {
      float4* q0 = reinterpret_cast<float4*>(scr) + (size_t)(pl * 2) * CG + cg;   // half 0; half 1 is CG float4 further
      const int qs = lanes_c / 4;                                                  // float4 per quantity
      q0[0] = make_float4(ds[0], ds[1], ds[2], ds[3]);            q0[CG] = make_float4(ds[4], ds[5], ds[6], ds[7]);
      q0[qs] = make_float4(dss[0], dss[1], dss[2], dss[3]);       q0[qs + CG] = make_float4(dss[4], dss[5], dss[6], dss[7]);
      q0[2 * qs] = make_float4(mx[0], mx[1], mx[2], mx[3]);       q0[2 * qs + CG] = make_float4(mx[4], mx[5], mx[6], mx[7]);
      if (TIES) { q0[3 * qs] = make_float4(tie[0], tie[1], tie[2], tie[3]); q0[3 * qs + CG] = make_float4(tie[4], tie[5], tie[6], tie[7]); }
      if (pl == 0) {
        float4* kq = reinterpret_cast<float4*>(k0s);
        kq[cg] = make_float4(k0[0], k0[1], k0[2], k0[3]); kq[CG + cg] = make_float4(k0[4], k0[5], k0[6], k0[7]);  // [half][cg][4]
      }
    }
    if (!L.scratch_aliases_stage) fence_proxy_async();  // this thread's generic reads of the stage precede the next bulk copy
    group_sync();  // [B] partials visible; nobody in the group reads the stage any more
    if (!L.scratch_aliases_stage && gtid == 0 && i + 2 < nb) issue(i + 2);
    if (gtid < C) {
      // thread t -> channel (half, cg_, j): consecutive threads read consecutive scratch words
      const int j = gtid & 3, cg_ = (gtid >> 2) % CG, half = gtid / (4 * CG);
      const int c = cg_ * 8 + half * 4 + j;
      const int o0 = (half * CG + cg_) * 4 + j, lstride = 2 * CG * 4;
      float D = 0.f, Q = 0.f, M = -INFINITY;
#pragma unroll 4
      for (int l = 0; l < NPL; ++l) {
        D += scr[l * lstride + o0]; Q += scr[lanes_c + l * lstride + o0]; M = fmaxf(M, scr[2 * lanes_c + l * lstride + o0]);
      }
      const float dm = D * (1.f / 81.f);
      const float mean = k0s[o0] + dm;
      const float sd = sqrtf(fmaxf(Q * (1.f / 81.f) - dm * dm, 0.f));
      float* pr = g.pool + b * 3 * C;
      pr[c] = mean; pr[C + c] = M; pr[2 * C + c] = sd;
      if (g.pool_bf) {
        bf16* pb = g.pool_bf + b * 3 * C;
        pb[c] = __float2bfloat16_rn(mean); pb[C + c] = __float2bfloat16_rn(M); pb[2 * C + c] = __float2bfloat16_rn(sd);
      }
      if (TIES) {
        float t = 0.f;
        for (int l = 0; l < NPL; ++l) if (scr[2 * lanes_c + l * lstride + o0] == M) t += scr[3 * lanes_c + l * lstride + o0];
        g.ties[b * C + c] = t;
      }
    }
    if (L.scratch_aliases_stage) fence_proxy_async();
    group_sync();  // [C] scratch (and k0s) may be rewritten
    if (L.scratch_aliases_stage && gtid == 0 && i + 2 < nb) issue(i + 2);
  }
}

}  // namespace

static int kbk_se_apply_tma_supported(int C, int S) {
  if (C % 8 != 0 || C < 64 || C > 256 || kGroupThreads % (C / 8) != 0 || kGroupThreads / (C / 8) > 81 || S < 1 || S > 64) return 0;
  return (make_layout(C, S, false).total <= 227 * 1024 && make_layout(C, S, true).total <= 227 * 1024) ? 1 : 0;
}

// Dispatch: the column pair in se_apply_col.cu by default. Training: 183 + 36 us vs 237 us for the TMA-staged kernel
// below at B = 8192. Evaluation: the same speed in isolation (95 + 24 us vs 118 us at B = 4096), but the column kernels
// need (almost) no shared memory, so in the two-branch rollout they run next to the other branch's convolution
// (+3-4 % positions/s; the TMA kernel's 166 KB ring cannot share an SM with a convolution). KB_SE_APPLY=tma / col forces one.
static int forced_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("KB_SE_APPLY"); v = !e ? 0 : (e[0] == 't' ? 1 : (e[0] == 'c' ? 2 : 0)); }
  return v;
}
int kbk_se_apply_supported(int C, int S) {
  return (kbk_se_apply_col_supported(C, S) || kbk_se_apply_tma_supported(C, S)) ? 1 : 0;
}

static int kbk_se_apply_tma(const SeApplyArgs& a, int num_sms, cudaStream_t st);

int kbk_se_apply(const SeApplyArgs& a, int num_sms, cudaStream_t st) { return kbk_se_apply_variant(a, 0, num_sms, st); }

// variant: 0 = default (KB_SE_APPLY or the column pair), 1 = TMA-staged kernel, 2 = column pair
int kbk_se_apply_variant(const SeApplyArgs& a, int variant, int num_sms, cudaStream_t st) {
  const bool col_ok = kbk_se_apply_col_supported(a.C, a.S) != 0, tma_ok = kbk_se_apply_tma_supported(a.C, a.S) != 0;
  KB_CHECK_ARG(col_ok || tma_ok, "se_apply: unsupported shape C=%d S=%d", a.C, a.S);
  const int f = variant != 0 ? variant : forced_variant();
  // ONE variant for every batch size: eval-mode results must not depend on what else is in the batch (the two kernels
  // sum in different orders, so mixing them by batch size would break bit-exact batch invariance)
  const bool want_col = f != 1;
  if (col_ok && (want_col || !tma_ok)) return kbk_se_apply_col(a, num_sms, st);
  return kbk_se_apply_tma(a, num_sms, st);
}

static int kbk_se_apply_tma(const SeApplyArgs& a, int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(kbk_se_apply_tma_supported(a.C, a.S), "se_apply: unsupported shape C=%d S=%d", a.C, a.S);
  KB_CHECK_ARG(a.z && a.res && a.out && a.bmean && a.w1 && a.b1 && a.w2 && a.b2 && a.pool && a.se_out, "se_apply: null pointer");
  if (a.B == 0) return KB_OK;
  const Layout L = make_layout(a.C, a.S, a.ties != nullptr);
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(se_apply_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    KB_CUDA_CHECK(cudaFuncSetAttribute(se_apply_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int grid = a.B < num_sms ? a.B : num_sms;
  if (a.ties) se_apply_tma_kernel<true><<<grid, kThreads, L.total, st>>>(a);
  else se_apply_tma_kernel<false><<<grid, kThreads, L.total, st>>>(a);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}
