// se_bwd.cu — backward of the squeeze-excite branch of a GlobalPoolBiasBlock in ONE launch.
//
// Reference: keisei/training/models/se_resnet.py:79-88 (autograd of
//   se = se_fc2(relu(se_fc1(mean(bn2(conv2))))); scale, shift = se.chunk(2); out = bn2 * sigmoid(scale) + shift).
//
// Replaces, per block and per training step: se_bwd_prep + 4 CUDA-core GEMMs (two weight gradients with a batch-long
// reduction, two data gradients) + 2 column sums + bn2_bwd_sums = 9 launches of latency-bound kernels. Everything here
// is per-board vector work on C <= 256 channels: thread = channel, the two weight matrices live in shared memory, the
// weight-gradient accumulators (3*S per thread) live in registers for the whole CTA lifetime and are flushed once
// with (vector) reductions into the fp32 gradient buffers.
//
//   inputs  s_du, s_duz [B][C]   per-board sums over the 81 pixels of du and du * z2 (du = gradient after the block's ReLU)
//           a2, b2 [C]           BatchNorm-2 as an affine map: zhat2 = z2 * a2 + b2
//           se [B][2C]           raw SE output (scale logits, shift); seh [B][S] hidden (post-ReLU); se_in [B][C] squeeze
//           bmean2 [B][C]        board mean of z2
//   outputs dse_in [B][C]        gradient wrt the squeeze input (consumed by block_bwd_dz2)
//           dW1 [S][C], db1 [S], dW2 [2C][S], db2 [2C]   accumulated (+=) into the pre-zeroed gradient buffers
//           sums [2][C] (double) += sum over boards and pixels of dzhat2 and dzhat2 * z2 (BatchNorm-2 backward)
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// Sum over the warp of v[idx] for idx = lane & (S-1): S-1 exchange shuffles instead of 5*S.
template <int S>
__device__ __forceinline__ float warp_multi_sum(float (&v)[S], int lane) {
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

struct SeBwdArgs {
  const float *s_du, *s_duz, *a2, *b2, *se, *seh, *se_in, *bmean2, *W1, *W2;
  float *dse_in, *dW1, *db1, *dW2, *db2;
  double* sums;
  int B, C, vec;  // vec: dW2 rows are 16-byte aligned (vector reductions)
};

template <int S>
__global__ void __launch_bounds__(256, 1) se_mlp_bwd_kernel(SeBwdArgs g) {
  extern __shared__ float sm[];
  const int C = g.C, c = threadIdx.x, lane = c & 31, warp = c >> 5, nwarps = C >> 5;
  float* W2t = sm;                    // [S][2C]: W2t[s][j] = W2[j][s]
  float* W1s = W2t + S * 2 * C;       // [S][C]
  float* seh_s = W1s + S * C;         // [2][S] (double-buffered by board parity)
  float* dseh_s = seh_s + 2 * S;      // [S]
  float* red = dseh_s + S;            // [8][S]
  for (int i = c; i < 2 * C * S; i += C) { const int j = i / S, s = i - j * S; W2t[s * 2 * C + j] = g.W2[i]; }
  for (int i = c; i < S * C; i += C) W1s[i] = g.W1[i];
  const float a2 = g.a2[c], b2 = g.b2[c];
  float accA[S], accB[S], acc1[S];
#pragma unroll
  for (int s = 0; s < S; ++s) { accA[s] = 0.f; accB[s] = 0.f; acc1[s] = 0.f; }
  float dbA = 0.f, dbB = 0.f, db1 = 0.f, bn1 = 0.f, bn2 = 0.f;

  int b = blockIdx.x;
  float du = 0.f, duz = 0.f, sc = 0.f, sin_ = 0.f, bm = 0.f, hh = 0.f;
  auto fetch = [&](int bb) {
    const size_t i = (size_t)bb * C + c;
    du = g.s_du[i]; duz = g.s_duz[i]; sc = g.se[(size_t)bb * 2 * C + c]; sin_ = g.se_in[i]; bm = g.bmean2[i];
    if (c < S) hh = g.seh[(size_t)bb * S + c];
  };
  if (b < g.B) fetch(b);
  int par = 0;
  for (; b < g.B; b += gridDim.x, par ^= 1) {
    const float cdu = du, cduz = duz, csin = sin_, cbm = bm;
    const float sg = sigmoid_f(sc);
    const float dsc = fmaf(a2, cduz, b2 * cdu) * sg * (1.f - sg);  // (sum_p du * zhat2) * sigmoid'
    const float dsh = cdu;
    float* hs = seh_s + par * S;
    if (c < S) hs[c] = hh;
    if (b + (int)gridDim.x < g.B) fetch(b + gridDim.x);  // next board's operands are in flight during this one
    __syncthreads();  // (A)
    float part[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float h = hs[s];
      accA[s] = fmaf(dsc, h, accA[s]);
      accB[s] = fmaf(dsh, h, accB[s]);
      part[s] = fmaf(dsc, W2t[s * 2 * C + c], dsh * W2t[s * 2 * C + C + c]);
    }
    dbA += dsc; dbB += dsh;
    const float ps = warp_multi_sum<S>(part, lane);
    if (lane < S) red[warp * S + lane] = ps;
    __syncthreads();  // (B)
    if (c < S) {
      float d = 0.f;
      for (int w = 0; w < nwarps; ++w) d += red[w * S + c];
      d = hs[c] > 0.f ? d : 0.f;   // ReLU of the hidden layer
      dseh_s[c] = d;
      db1 += d;
    }
    __syncthreads();  // (C)
    float dsi = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float d = dseh_s[s];
      acc1[s] = fmaf(d, csin, acc1[s]);
      dsi = fmaf(d, W1s[s * C + c], dsi);
    }
    g.dse_in[(size_t)b * C + c] = dsi;
    bn1 += fmaf(sg, cdu, dsi);
    bn2 += fmaf(sg, cduz, dsi * cbm);
  }

  // flush: one reduction per accumulator per CTA into the gradient buffers
  float* r2a = g.dW2 + (size_t)c * S;
  float* r2b = g.dW2 + (size_t)(C + c) * S;
  if (S % 4 == 0 && g.vec) {
#pragma unroll
    for (int s = 0; s < S; s += 4) {
      atomicAdd(reinterpret_cast<float4*>(r2a + s), make_float4(accA[s], accA[s + 1], accA[s + 2], accA[s + 3]));
      atomicAdd(reinterpret_cast<float4*>(r2b + s), make_float4(accB[s], accB[s + 1], accB[s + 2], accB[s + 3]));
    }
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) { atomicAdd(r2a + s, accA[s]); atomicAdd(r2b + s, accB[s]); }
  }
#pragma unroll
  for (int s = 0; s < S; ++s) atomicAdd(g.dW1 + (size_t)s * C + c, acc1[s]);
  atomicAdd(g.db2 + c, dbA);
  atomicAdd(g.db2 + C + c, dbB);
  if (c < S) atomicAdd(g.db1 + c, db1);
  atomicAdd(g.sums + c, (double)bn1);
  atomicAdd(g.sums + C + c, (double)bn2);
}

size_t smem_bytes(int C, int S) { return (size_t)(S * 2 * C + S * C + 2 * S + S + 8 * S) * sizeof(float); }

template <int S>
int launch(const SeBwdArgs& g, int num_sms, cudaStream_t st) {
  const size_t smem = smem_bytes(g.C, S);
  static bool attr_done = false;  // per instantiation; idempotent, so a benign race between host threads
  if (!attr_done) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(se_mlp_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const int grid = g.B < num_sms ? g.B : num_sms;
  se_mlp_bwd_kernel<S><<<grid, g.C, smem, st>>>(g);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

}  // namespace

int kbk_se_mlp_bwd_supported(int C, int S) {
  return C % 32 == 0 && C >= 32 && C <= 256 && (S == 4 || S == 8 || S == 16 || S == 32) && smem_bytes(C, S) <= 200 * 1024;
}

int kbk_se_mlp_bwd(const float* s_du, const float* s_duz, const float* a2, const float* b2, const float* se,
                   const float* seh, const float* se_in, const float* bmean2, const float* W1, const float* W2,
                   float* dse_in, float* dW1, float* db1, float* dW2, float* db2, double* sums, int B, int C, int S,
                   int num_sms, cudaStream_t st) {
  KB_CHECK_ARG(kbk_se_mlp_bwd_supported(C, S), "se_mlp_bwd: C=%d S=%d unsupported", C, S);
  if (B == 0) return KB_OK;
  SeBwdArgs g{s_du, s_duz, a2, b2, se, seh, se_in, bmean2, W1, W2, dse_in, dW1, db1, dW2, db2, sums, B, C,
              (S % 4 == 0 && ((uintptr_t)dW2 & 15) == 0) ? 1 : 0};
  if (num_sms <= 0) num_sms = 148;
  switch (S) {
    case 4: return launch<4>(g, num_sms, st);
    case 8: return launch<8>(g, num_sms, st);
    case 16: return launch<16>(g, num_sms, st);
    default: return launch<32>(g, num_sms, st);
  }
}
