// peer_sync.cu — the SyncBatchNorm statistic exchange as ONE kernel over NVLink peer memory.
//
// Reference: torch.nn.SyncBatchNorm under DDP (katago_loop.py:494-497): per BatchNorm layer an all-gather / all-reduce
// of (2*C,) statistics in the forward and in the backward. With NCCL that is 164 collectives of 4 KB per step for the
// 40-block model — pure latency (~21 us each: two stream hand-offs plus a collective kernel launch).
//
// Here every rank owns one peer-mapped buffer (cudaMalloc + CUDA IPC, mapped by all ranks of the node):
//     data  [n_slots][world][slot_doubles]  double      flags [n_slots][world]  uint64
// and an exchange is one CTA on the caller's stream:
//   1. store this rank's n doubles into slot (seq % n_slots), row `rank`, of EVERY rank's buffer (NVLink P2P stores);
//   2. fence.sys, then release-store seq+1 into flags[slot][rank] of every rank's buffer;
//   3. acquire-spin on the `world` flags of the LOCAL buffer until all read >= seq+1. The wait is bounded in TIME
//      (ctx->timeout_ms, default 120 s — rank skew of many seconds is normal: checkpointing, evaluation, first-call
//      builds): a lost peer then poisons the result with NaN instead of hanging the GPU AND raises the host-visible
//      status word (ctx->status, pinned mapped host memory) so the trainer reports "peer rank lost" instead of
//      training on NaN; the BatchNorm finalize kernel does not commit non-finite statistics to the running buffers;
//   4. sum the `world` rows of the local slot in rank order — every rank adds the same numbers in the same order, so
//      the statistics are bit-identical across ranks — and write the result over the input.
// A rank can be at most one exchange ahead of the slowest (it needs that rank's flag to finish the next one), so
// n_slots >= 2 rules out overwriting a slot that is still being read; the default is 4.
// The BatchNorm finalize kernel that consumes the sums follows on the same stream.
#include "kb_common.cuh"
#include "../../include/keisei_b200.h"

namespace {

constexpr int kMaxWorld = 16;

struct PeerPtrs { double* data[kMaxWorld]; unsigned long long* flags[kMaxWorld]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void peer_exchange(double* buf, int n, const PeerPtrs& pp, int rank, int world,
                                              unsigned long long seq, int n_slots, long long slot_doubles,
                                              unsigned long long timeout_ns, unsigned long long* status) {
  __shared__ int timed_out;
  const int slot = (int)(seq % (unsigned long long)n_slots);
  const size_t row = ((size_t)slot * world + rank) * (size_t)slot_doubles;
  if (threadIdx.x == 0) timed_out = 0;
  // 1. publish
  for (int p = 0; p < world; ++p) {
    double* dst = pp.data[p] + row;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal
  if (threadIdx.x < world) st_release_sys(pp.flags[threadIdx.x] + (size_t)slot * world + rank, seq + 1ull);
  // 3. wait for every rank's contribution to land HERE
  if (threadIdx.x < world) {
    const unsigned long long* f = pp.flags[rank] + (size_t)slot * world + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(f) < seq + 1ull) {
      __nanosleep(64);
      if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > timeout_ns) {   // a peer is gone
        timed_out = 1;
        if (status != nullptr) {   // first failure wins; the host reads this word (mapped pinned memory) once per step
          atomicCAS_system(status, 0ull, ((unsigned long long)(threadIdx.x + 1) << 48) | ((seq + 1ull) & 0xffffffffffffull));
          __threadfence_system();
        }
        break;
      }
    }
  }
  __syncthreads();
  // 4. reduce in rank order (L2 reads: the rows were written by remote stores)
  const double* mine = pp.data[rank] + (size_t)slot * world * (size_t)slot_doubles;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcg(mine + (size_t)r * slot_doubles + i);
    buf[i] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(double* buf, int n, PeerPtrs pp, int rank, int world,
                                                                unsigned long long seq, int n_slots, long long slot_doubles,
                                                                unsigned long long timeout_ns, unsigned long long* status) {
  peer_exchange(buf, n, pp, rank, world, seq, n_slots, slot_doubles, timeout_ns, status);
}

// Single-GPU emulation for tests: ONE cooperative launch whose block r plays rank r on that rank's data and buffers.
// (Ranks emulated as separate launches that wait on one another are not guaranteed to run at the same time.)
struct EmuArgs { double* buf[kMaxWorld]; PeerPtrs pp[kMaxWorld]; unsigned long long* status[kMaxWorld]; };
__global__ void __launch_bounds__(256) peer_allreduce_emulate_kernel(EmuArgs a, int n, int world, unsigned long long seq, int n_slots,
                                                                    long long slot_doubles, unsigned long long timeout_ns) {
  const int r = blockIdx.x;
  peer_exchange(a.buf[r], n, a.pp[r], r, world, seq, n_slots, slot_doubles, timeout_ns, a.status[r]);
}

}  // namespace

extern "C" long long kb_peer_buffer_bytes(int world, int n_slots, long long slot_doubles) {
  if (world < 1 || world > kMaxWorld || n_slots < 2 || slot_doubles < 1) return -1;
  return (long long)n_slots * world * slot_doubles * 8 + (long long)n_slots * world * 8;
}

extern "C" int kb_peer_buffer_create(long long bytes, void** ptr, unsigned char* handle64) {
  KB_CHECK_ARG(bytes > 0 && ptr && handle64, "kb_peer_buffer_create: bad arguments");
  void* p = nullptr;
  KB_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  KB_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  KB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    kb_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p);
    return KB_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return KB_OK;
}

extern "C" int kb_peer_buffer_open(const unsigned char* handle64, void** ptr) {
  KB_CHECK_ARG(handle64 && ptr, "kb_peer_buffer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  KB_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KB_OK;
}

extern "C" int kb_peer_buffer_close(void* ptr) {
  if (ptr) KB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return KB_OK;
}

extern "C" int kb_peer_buffer_destroy(void* ptr) {
  if (ptr) KB_CUDA_CHECK(cudaFree(ptr));
  return KB_OK;
}

// Host-visible failure word of the exchange kernels: pinned, device-mapped host memory (the kernel writes it with a
// system-scope atomic on timeout; the host reads it without a copy). 0 = no failure.
extern "C" int kb_peer_status_create(unsigned long long** word) {
  KB_CHECK_ARG(word != nullptr, "kb_peer_status_create: null pointer");
  void* p = nullptr;
  KB_CUDA_CHECK(cudaHostAlloc(&p, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
  *(volatile unsigned long long*)p = 0ull;
  *word = (unsigned long long*)p;
  return KB_OK;
}

extern "C" int kb_peer_status_destroy(unsigned long long* word) {
  if (word) KB_CUDA_CHECK(cudaFreeHost(word));
  return KB_OK;
}

extern "C" int kb_peer_allreduce_f64(void* buf, long long n, const kb_peer_ctx* ctx, unsigned long long seq, kb_stream_t stream) {
  KB_CHECK_ARG(buf && ctx, "kb_peer_allreduce_f64: null pointer");
  KB_CHECK_ARG(ctx->world >= 1 && ctx->world <= kMaxWorld && ctx->rank >= 0 && ctx->rank < ctx->world && ctx->n_slots >= 2,
               "kb_peer_allreduce_f64: bad context (world=%d rank=%d slots=%d)", ctx->world, ctx->rank, ctx->n_slots);
  KB_CHECK_ARG(n >= 0 && n <= ctx->slot_doubles, "kb_peer_allreduce_f64: %lld doubles exceed the slot size %lld", n, ctx->slot_doubles);
  if (n == 0 || ctx->world == 1) return KB_OK;
  PeerPtrs pp;
  const size_t data_doubles = (size_t)ctx->n_slots * ctx->world * (size_t)ctx->slot_doubles;
  for (int r = 0; r < ctx->world; ++r) {
    KB_CHECK_ARG(ctx->peers[r] != nullptr, "kb_peer_allreduce_f64: peer %d is not mapped", r);
    pp.data[r] = (double*)ctx->peers[r];
    pp.flags[r] = (unsigned long long*)((double*)ctx->peers[r] + data_doubles);
  }
  const unsigned long long timeout_ns = (unsigned long long)(ctx->timeout_ms > 0 ? ctx->timeout_ms : 120000) * 1000000ull;
  peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((double*)buf, (int)n, pp, ctx->rank, ctx->world, seq, ctx->n_slots,
                                                                ctx->slot_doubles, timeout_ns, ctx->status);
  KB_CUDA_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_peer_allreduce_emulate(void* const* bufs, long long n, kb_peer_ctx* const* ctxs, int world, kb_stream_t stream) {
  KB_CHECK_ARG(bufs && ctxs && world >= 1 && world <= kMaxWorld, "kb_peer_allreduce_emulate: bad arguments");
  EmuArgs a; memset(&a, 0, sizeof(a));
  const kb_peer_ctx* c0 = ctxs[0];
  KB_CHECK_ARG(c0 && n >= 1 && n <= c0->slot_doubles && c0->n_slots >= 2, "kb_peer_allreduce_emulate: bad context / size");
  const size_t data_doubles = (size_t)c0->n_slots * world * (size_t)c0->slot_doubles;
  for (int r = 0; r < world; ++r) {
    const kb_peer_ctx* c = ctxs[r];
    KB_CHECK_ARG(c && c->world == world && c->rank == r && c->seq == c0->seq && c->n_slots == c0->n_slots && c->slot_doubles == c0->slot_doubles,
                 "kb_peer_allreduce_emulate: context %d does not describe rank %d of %d", r, r, world);
    a.buf[r] = (double*)bufs[r];
    a.status[r] = c->status;
    for (int q = 0; q < world; ++q) {
      KB_CHECK_ARG(c->peers[q] != nullptr, "kb_peer_allreduce_emulate: peer %d of rank %d is not mapped", q, r);
      a.pp[r].data[q] = (double*)c->peers[q];
      a.pp[r].flags[q] = (unsigned long long*)((double*)c->peers[q] + data_doubles);
    }
  }
  int nn = (int)n, w = world, slots = c0->n_slots;
  unsigned long long seq = c0->seq;
  long long sd = c0->slot_doubles;
  unsigned long long timeout_ns = (unsigned long long)(c0->timeout_ms > 0 ? c0->timeout_ms : 120000) * 1000000ull;
  void* args[] = {&a, &nn, &w, &seq, &slots, &sd, &timeout_ns};
  KB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)peer_allreduce_emulate_kernel, dim3((unsigned)world), dim3(256), args, 0,
                                            (cudaStream_t)stream));
  kb_count_launch();
  for (int r = 0; r < world; ++r) ctxs[r]->seq += 1ull;
  return KB_OK;
}

// kb_allreduce_hook-compatible entry: `user` is a kb_peer_ctx whose `seq` advances by one per exchange (every rank makes
// the same sequence of calls, so the counters agree without communication).
extern "C" int kb_peer_allreduce_hook(void* user, void* buf, long long n_doubles, kb_stream_t stream) {
  kb_peer_ctx* ctx = (kb_peer_ctx*)user;
  if (!ctx) { kb_set_error("kb_peer_allreduce_hook: null context"); return KB_ERR_INVALID; }
  const int r = kb_peer_allreduce_f64(buf, n_doubles, ctx, ctx->seq, stream);
  if (r == KB_OK) ctx->seq += 1ull;
  return r;
}
