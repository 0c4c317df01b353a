// schedule_common.cuh — host-side helpers shared by the network schedules (model.cu: SE-ResNet,
// resnet.cu: plain ResNet): workspace bump allocator and the Linear-layer forward/backward
// compositions over the SIMT GEMM.
#pragma once
#include <stdlib.h>
#include <string.h>
#include "kb_common.cuh"
#include "kb_kernels.h"

namespace kbs {

struct Bump {
  char* base; size_t off;
  explicit Bump(void* b) : base((char*)b), off(0) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f32(size_t n) { return (float*)take(n * sizeof(float)); }
};

// TF32 tensor-core inner product for the small GEMMs (GemmArgs.tf32): switched on for the duration of a bf16 / AMP
// backward schedule, where the reference runs these Linear layers in bf16 under autocast; never in the fp32 path.
inline int& gemm_tf32_flag() { static thread_local int f = 0; return f; }
struct Tf32Scope {
  int prev;
  explicit Tf32Scope(int on) : prev(gemm_tf32_flag()) { gemm_tf32_flag() = on; }
  ~Tf32Scope() { gemm_tf32_flag() = prev; }
};
inline GemmArgs gemm_base() { GemmArgs g; memset(&g, 0, sizeof(g)); g.splitk = 1; g.tf32 = gemm_tf32_flag(); return g; }
inline ConvEpi epi_base() { ConvEpi e; memset(&e, 0, sizeof(e)); e.board_scale = 1.f; return e; }

// y[M,N] = act(x[M,K] * W[N,K]^T + bias)
inline int linear_fwd(const void* x, int x_dtype, long long ldx, int M, int K, const float* W, int N, const float* bias, int relu,
               void* y, int y_dtype, long long ldy, cudaStream_t st) {
  GemmArgs g = gemm_base();
  g.A = x; g.a_dtype = x_dtype; g.lda = ldx;
  g.B = W; g.b_dtype = KB_F32; g.ldb = K; g.transB = 1;
  g.C = y; g.c_dtype = y_dtype; g.ldc = ldy; g.bias = bias; g.relu = relu;
  g.M = M; g.N = N; g.K = K;
  return kbk_gemm(g, st);
}
// dx[M,K] = (dy[M,N] * W[N,K]) masked by mask_src > 0 ; accumulate -> atomic add onto dx
inline int linear_bwd_x(const void* dy, int dy_dtype, long long ldy, int M, int N, const float* W, int K, void* dx, int dx_dtype,
                 long long lddx, const float* mask_src, long long ld_mask, int accumulate, cudaStream_t st) {
  GemmArgs g = gemm_base();
  g.A = dy; g.a_dtype = dy_dtype; g.lda = ldy;
  g.B = W; g.b_dtype = KB_F32; g.ldb = K; g.transB = 0;
  g.C = dx; g.c_dtype = dx_dtype; g.ldc = lddx; g.mask_src = mask_src; g.ld_mask = ld_mask;
  g.M = M; g.N = K; g.K = N; g.splitk = accumulate ? 2 : 1;
  return kbk_gemm(g, st);
}
// dW[N,K] += dy[M,N]^T * x[M,K] ; db[N] += colsum(dy)
inline int linear_bwd_w(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, int M, int N, int K,
                 float* dW, float* db, cudaStream_t st) {
  GemmArgs g = gemm_base();
  g.A = dy; g.a_dtype = dy_dtype; g.lda = ldy; g.transA = 1;
  g.B = x; g.b_dtype = x_dtype; g.ldb = ldx; g.transB = 0;
  g.C = dW; g.c_dtype = KB_F32; g.ldc = K;
  g.M = N; g.N = K; g.K = M;
  const int tiles = kb_ceil_div(N, 64) * kb_ceil_div(K, 64);
  static const int ctas_per_sm = [] { const char* e = getenv("KB_WGEMM_CTAS"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
  // K slices for ~2 CTAs per SM. Measured on B200 with the lean main loop (same box, interleaved; KB_WGEMM_CTAS): 8192-sample
  // step 196.0 / 197.6 ms at 5 per SM, 194.7 / 195.4 ms at 2, 196.6 / 195.8 ms at 3; 1024 samples 28.0 / 28.1 / 27.7 ms —
  // fewer, longer slices mean fewer atomic partial tiles (a 128 x 768 gradient was being written 31 times).
  int sk = kb_ceil_div(148 * ctas_per_sm, tiles);
  const int max_sk = kb_ceil_div(M, 64);
  if (sk > max_sk) sk = max_sk;
  g.splitk = sk < 2 ? 2 : sk;  // always the atomic epilogue: dW accumulates into the pre-zeroed gradient
  if (int r = kbk_gemm(g, st)) return r;
  if (db) return kbk_colsum(dy, dy_dtype, ldy, 0, 0, M, N, db, st);
  return KB_OK;
}

// RAII wrapper of the grouped-launch deferral (kb_kernels.h): problems issued inside are collected and launched together
// by flush() / close(); an early return through KB_TRY still ends the group (without leaving it open on the thread).
struct GemmGroupScope {
  cudaStream_t st; int rc; bool open;
  explicit GemmGroupScope(cudaStream_t s) : st(s), rc(kbk_gemm_group_begin()), open(rc == KB_OK) {}
  int status() const { return rc; }
  int flush() { return kbk_gemm_group_flush(st); }
  int close() { open = false; return kbk_gemm_group_end(st); }
  ~GemmGroupScope() { if (open) kbk_gemm_group_end(st); }
};

#define KB_TRY(expr) do { int r__ = (expr); if (r__ != KB_OK) return r__; } while (0)

// Side stream for the backward schedules: the weight-gradient convolutions (tensor-bound, 45 registers, no dependants
// until the optimiser) run on it while the caller's stream carries the data-gradient chain — so the HBM-bound
// elementwise passes of that chain share the SMs with a tensor-bound kernel instead of alternating with it.
// One lazily created stream + event pool PER HOST THREAD AND DEVICE (concurrent callers never share
// events); everything is joined back into the caller's stream before the schedule returns, so callers see plain
// stream-ordered semantics. KB_BWD_OVERLAP=0 disables it; it is also off while the caller's stream is being captured.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaStream_t s2 = nullptr;   // second side stream: the small global-pool-MLP backward GEMMs (model.cu), off the main chain
  cudaEvent_t* ev = nullptr;
  int n_ev = 0, device = -1;
  ~SideStream() {
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    free(ev);
    if (s) cudaStreamDestroy(s);
    if (s2) cudaStreamDestroy(s2);
  }
  int ensure(int device_id, int events) {
    if (s == nullptr || device != device_id) {
      if (s) { for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]); free(ev); ev = nullptr; n_ev = 0; cudaStreamDestroy(s); s = nullptr; if (s2) { cudaStreamDestroy(s2); s2 = nullptr; } }
      int lo = 0, hi = 0;
      KB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      // Lowest priority (= the default stream's): measured on B200, a high-priority side stream does not help — the
      // convolutions and the streaming kernels both live off L2 -> SM bandwidth (a conv moves ~17 TB/s of operands), so
      // co-residency buys little; what the side stream does buy is that the tail of one tensor-bound kernel overlaps
      // the head of the next (KB_SIDE_PRIO=hi to compare).
      static int prio_hi = -1;
      if (prio_hi < 0) { const char* e = getenv("KB_SIDE_PRIO"); prio_hi = (e && e[0] == 'h') ? 1 : 0; }
      KB_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio_hi ? hi : lo));
      KB_CUDA_CHECK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, lo));
      device = device_id;
    }
    if (events > n_ev) {
      cudaEvent_t* ne = (cudaEvent_t*)realloc(ev, sizeof(cudaEvent_t) * events);
      if (!ne) { kb_set_error("out of host memory for the event pool"); return KB_ERR_INVALID; }
      ev = ne;
      for (; n_ev < events; ++n_ev) KB_CUDA_CHECK(cudaEventCreateWithFlags(&ev[n_ev], cudaEventDisableTiming));
    }
    return KB_OK;
  }
};
inline SideStream& side_stream_tls() { static thread_local SideStream ss; return ss; }
inline bool bwd_overlap_enabled(cudaStream_t st) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("KB_BWD_OVERLAP"); env = (e && e[0] == '0') ? 0 : 1; }
  if (!env) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs == cudaStreamCaptureStatusNone;
}


}  // namespace kbs
