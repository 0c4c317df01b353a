/* keisei_b200.h — C-ABI of libkeisei_b200.so: the B200 (sm_100a) kernels behind Keisei's
 * data-parallel hot path (SE-ResNet forward/backward + KataGo-PPO update + GAE).
 *
 * The reference (tachyon-beep/keisei) has NO FFI on this path — every device op is a PyTorch
 * library call from Python. Each entry point below therefore names the reference Python call
 * site(s) it replaces (file:line relative to the reference root); INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer
 * unless stated otherwise; the caller owns all memory including workspaces; nothing is allocated
 * and no global mutable state is kept (thread-safe; kernels go to the caller's `stream`);
 * return 0 on success or a negative code, never throw; kb_last_error() gives the message for the
 * calling thread. dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef KEISEI_B200_H
#define KEISEI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* kb_stream_t;

#define KB_DTYPE_F32 0
#define KB_DTYPE_BF16 1

/* ---- library ---- */
int kb_abi_version(void);
int kb_compiled_sm(void);                     /* 100 */
const char* kb_last_error(void);              /* thread-local message of the last failing call */
unsigned long long kb_launch_count(void);     /* kernels launched by this library so far */
int kb_device_sm_count(int device);

/* ---- GAE: keisei/training/gae.py:8-73 compute_gae, :76-148 compute_gae_padded,
 *      :151-218 compute_gae_gpu, :221-296 compute_gae_padded_gpu  (one launch instead of 2*T) ----
 * (T,N) row-major. terminated: term_kind 0 = uint8/bool, 1 = same float type as the values.
 * override_nv: optional (T,N), NaN = "no override". lengths: optional (N,) int32 -> padded variant
 * (bootstrap stamped at lengths[n]-1). dtype_is_f64: 0 = float32 tensors, 1 = float64. */
int kb_gae_scan(const void* rewards, const void* values, const void* terminated, int term_kind,
                const void* next_value, const void* override_nv, const int* lengths, void* adv,
                int T, int N, double gamma, double lam, int dtype_is_f64, kb_stream_t stream);
/* keisei/training/katago_ppo.py:797-798: (A - mean) / (std_unbiased + eps), in place, no-op for n <= 1 */
int kb_advantage_normalize(float* adv, long long n, float eps, kb_stream_t stream);

/* ---- rollout policy: keisei/training/katago_ppo.py:589-613 (also katago_loop.py:345-357, 407-418;
 *      match_utils.py:211-224): legal-count guard, mask -> softmax -> sample -> log_prob, scalar
 *      value P(W)-P(L) with optional score blend (value_adapter.py:79-96) ----
 * logits: (B, A) with row stride `row_stride` elements. mask + mask_kind + mask_pitch (all three policy entry points):
 *   KB_MASK_BYTES (0): (B, A) uint8 / bool, `mask_pitch` bytes per row (the reference's buffer layout, katago_ppo.py:160);
 *   KB_MASK_BITS  (1): bit-packed, `mask_pitch` 32-bit words per row, action i = bit (i & 31) of word i >> 5 (1,408 B instead
 *                      of 11,259 B per row: the device-resident rollout buffer, kb_pack_mask_bits);
 *   KB_MASK_NONE  (2): every action legal, mask may be NULL (supervised policy cross-entropy, sl/trainer.py:147-149).
 * logprob_mode 0 = float32 Categorical semantics, 1 = bfloat16-autocast semantics (eps = 2^-7).
 * forced_actions (optional, (B,) int64): report the log-prob of these instead of sampling.
 * flags[0] += rows with zero legal actions; legal_count (B,) int32. */
#define KB_MASK_BYTES 0
#define KB_MASK_BITS 1
#define KB_MASK_NONE 2
int kb_policy_sample(const void* logits, int logits_dtype, long long row_stride, const void* mask,
                     const float* value_logits, const float* score_lead, float alpha, int B, int A,
                     unsigned long long seed, unsigned long long offset, int logprob_mode,
                     const long long* forced_actions, long long* actions, float* logp, float* values,
                     int* legal_count, int* flags, int mask_kind, long long mask_pitch, kb_stream_t stream);
/* (rows, A) uint8 / bool -> (rows, words) bit-packed legal masks (words * 32 >= A; padding bits are 0) */
int kb_pack_mask_bits(const void* mask_bytes, void* bits, long long rows, int A, int words, kb_stream_t stream);
/* One launch gathers a shuffled minibatch from the device-resident rollout storage (katago_ppo.py:829-841 does eight
 * index-gathers): observations (rows of `obs_floats` fp32, even), bit-packed masks (`words` per row) and the per-sample
 * scalars of rows idx[0..n_out) out of n_src stored samples. */
int kb_gather_minibatch(const float* obs, const void* mask_bits, const long long* actions, const float* old_lp,
                        const float* adv, const long long* cats, const float* score, const float* returns,
                        const long long* idx, long long n_src, int n_out, int obs_floats, int words, float* o_obs,
                        void* o_bits, long long* o_actions, float* o_old_lp, float* o_adv, long long* o_cats,
                        float* o_score, float* o_returns, kb_stream_t stream);

/* ---- update losses: keisei/training/katago_ppo.py:858-888 (NaN / zero-legal guards, masked
 *      log_softmax, gather, entropy) and :33-43 ppo_clip_loss ----
 * out2 = [policy_loss, entropy]; dlogp (B,) = d policy_loss / d new_logp; flags[0] zero-legal rows,
 * flags[1] rows with NaN raw logits. */
int kb_ppo_policy_fwd(const void* logits, int logits_dtype, long long row_stride, const void* mask,
                      const long long* actions, const float* old_logp, const float* adv, int B, int A,
                      float clip_eps, float* new_logp, float* row_entropy, float* row_lse, float* dlogp,
                      float* out2, int* flags, int mask_kind, long long mask_pitch, kb_stream_t stream);
/* dlogits = g_policy[0] * d policy_loss + g_entropy[0] * d entropy, written once (zeros on illegal) */
int kb_ppo_policy_bwd(const void* logits, int logits_dtype, long long row_stride, const void* mask,
                      const long long* actions, int B, int A, const float* row_lse,
                      const float* row_entropy, const float* dlogp, const float* g_policy,
                      const float* g_entropy, void* dlogits, long long d_row_stride, int mask_kind, long long mask_pitch,
                      kb_stream_t stream);
/* keisei/training/katago_ppo.py:46-57 wdl_cross_entropy_loss (ignore_index=-1, all-ignored -> 0),
 * :910-912 score MSE; value_adapter.py:98-126. out3 = [value_loss, score_loss, n_valid] */
int kb_value_losses_fwd(const float* value_logits, const long long* cats, const float* score_pred,
                        const float* score_tgt, int B, float* out3, kb_stream_t stream);
int kb_value_losses_bwd(const float* value_logits, const long long* cats, const float* score_pred,
                        const float* score_tgt, int B, const float* out3, const float* g_value,
                        const float* g_score, float* dvalue_logits, float* dscore, kb_stream_t stream);

/* ---- optimiser tail: keisei/training/katago_ppo.py:926-933 (GradScaler.unscale_, clip_grad_norm_, Adam.step) as two
 *      passes over the flat fp32 gradient the backward produces (csrc/optim.cu) ----
 * out2 (device, 2 doubles) = [sum of squares of flat_grad, 1.0 if any element is non-finite]; zeroed by the call. */
int kb_flat_grad_stats(const float* flat_grad, long long n, double* out2, int num_sms, kb_stream_t stream);
/* One launch over every parameter: g = flat_grad * inv_scale / grad_div, clipped to max_norm by the global 2-norm taken
 * from `stats`, then torch.optim.Adam's update (no weight decay / amsgrad) in place on PyTorch's own parameter and moment
 * tensors (DEVICE arrays of pointers p/m/v, [n_tensors]); `steps` (device, [n_tensors] floats) are the optimizer's `step`
 * scalars, incremented unless the gradient was non-finite, in which case nothing is touched (GradScaler skip).
 * Chunking: chunk c covers elements [chunk_start[c], chunk_start[c] + chunk) of tensor chunk_tensor[c]; g_off[t] is the
 * tensor's offset in the flat gradient. grad_norm_out / found_inf_out: device floats. */
int kb_adam_step_flat(const float* flat_grad, void* const* p_ptrs, void* const* m_ptrs, void* const* v_ptrs, float* steps,
                      const long long* g_off, const long long* sizes, const int* chunk_tensor, const long long* chunk_start,
                      int n_tensors, int n_chunks, int chunk, const double* stats, const float* inv_scale, float grad_div,
                      float max_norm, float lr, float beta1, float beta2, float eps, float* grad_norm_out,
                      float* found_inf_out, kb_stream_t stream);

/* ---- SE-ResNet: keisei/training/models/se_resnet.py:132-159 SEResNetModel._forward_impl,
 *      :68-90 GlobalPoolBiasBlock.forward, :93-98 _global_pool, and their autograd ----
 * params / buffers / grads are HOST arrays of device pointers in the reference's registration
 * order (see keisei_b200/csrc/model.cu header): params[16 + 14*num_blocks] float32,
 * buffers[6 + 6*num_blocks] (running_mean, running_var float32; num_batches_tracked int64). */
typedef struct {
  int num_blocks, channels, se_hidden, gpool_channels, policy_channels, value_fc, score_fc, obs_channels;
} kb_seresnet_desc;

long long kb_seresnet_num_params(const kb_seresnet_desc* d);
long long kb_seresnet_num_buffers(const kb_seresnet_desc* d);
long long kb_seresnet_wpack_bytes(const kb_seresnet_desc* d, int dtype);
long long kb_seresnet_workspace_bytes(const kb_seresnet_desc* d, int B, int training, int dtype);
/* repack conv weights (activation dtype, forward + dgrad layouts) and fold eval-mode BatchNorm */
int kb_seresnet_pack_weights(const kb_seresnet_desc* d, const void* const* params, const void* const* buffers,
                             int dtype, void* wpack, long long wpack_bytes, kb_stream_t stream);
/* obs: (B, obs_channels, 9, 9) float32 NCHW. policy_out: (B, policy_pitch >= 11259) rows of
 * (9,9,139) logits in the activation dtype; value_out (B,3), score_out (B,1) float32.
 * training != 0: batch-statistics BatchNorm, activations saved in `workspace` for
 * kb_seresnet_backward; running stats (momentum 0.1, unbiased var) are updated in place in `buffers`
 * (and num_batches_tracked incremented) when new_stats == NULL, else written to new_stats
 * ([2*num_blocks+2][2][max(channels, policy_channels)] float32: mean row, var row per BN layer in
 * registration order) leaving `buffers` untouched. use_tc: 1 = tcgen05 convolutions when shape/dtype allow. */
int kb_seresnet_forward(const kb_seresnet_desc* d, const void* const* params, void* const* buffers,
                        float* new_stats, const void* wpack, const float* obs, int B, int training, int dtype, void* workspace,
                        long long ws_bytes, void* policy_out, long long policy_pitch, float* value_out,
                        float* score_out, int use_tc, int num_sms, kb_stream_t stream);
/* SyncBatchNorm variants — the reference's data-parallel default (`sync_batchnorm = true`: katago_loop.py:494-508 wraps
 * the model in torch.nn.SyncBatchNorm before DDP). `hook(user, buf, n_doubles, stream)` must sum `buf` (device doubles)
 * over the `world` ranks in stream order and return 0; it is called once per BatchNorm layer in the forward (batch
 * sum / sum of squares) and once in the backward (sum dz, sum dz*z; dgamma/dbeta keep the rank's own share, as
 * torch.nn.SyncBatchNorm does). hook == NULL or world == 1 reproduces kb_seresnet_forward / _backward. */
typedef int (*kb_allreduce_hook)(void* user, void* buf, long long n_doubles, kb_stream_t stream);
int kb_seresnet_forward_sync(const kb_seresnet_desc* d, const void* const* params, void* const* buffers,
                             float* new_stats, const void* wpack, const float* obs, int B, int training, int dtype,
                             void* workspace, long long ws_bytes, void* policy_out, long long policy_pitch,
                             float* value_out, float* score_out, int use_tc, int num_sms, kb_allreduce_hook hook,
                             void* hook_user, int world, kb_stream_t stream);
int kb_seresnet_backward_sync(const kb_seresnet_desc* d, const void* const* params, const void* wpack, int B,
                              int dtype, void* workspace, long long ws_bytes, const void* dpolicy,
                              long long policy_pitch, const float* dvalue, const float* dscore, void* const* grads,
                              int use_tc, int num_sms, kb_allreduce_hook hook, void* hook_user, int world,
                              kb_stream_t stream);
/* ---- SyncBatchNorm exchange over NVLink peer memory (csrc/peer_sync.cu): one kernel per exchange instead of an NCCL
 *      collective. Every rank of the node creates one buffer, ships its 64-byte CUDA IPC handle to the others (any host
 *      transport: torch.distributed all_gather_object in keisei_b200/distributed.py), opens theirs, and fills a
 *      kb_peer_ctx struct on the host (caller-owned). These are the only entry points of the library that allocate:
 *      kb_peer_buffer_create = cudaMalloc (IPC needs a whole allocation), released by kb_peer_buffer_destroy. */
typedef struct kb_peer_ctx {
  void* peers[16];             /* device pointers of every rank's buffer as mapped in THIS process; peers[rank] = own buffer */
  int rank, world, n_slots, reserved;
  long long slot_doubles;      /* capacity of one exchange (>= 2 * max channels) */
  unsigned long long seq;      /* exchanges done so far; advanced by kb_peer_allreduce_hook */
  unsigned long long* status;  /* kb_peer_status_create word or NULL: set non-zero by an exchange that timed out waiting for a
                                  peer (bits 0-47: seq+1 of the exchange, bits 48-63: 1 + the rank that never arrived) */
  long long timeout_ms;        /* how long an exchange waits for the slowest rank; <= 0: 120 000 ms */
} kb_peer_ctx;
long long kb_peer_buffer_bytes(int world, int n_slots, long long slot_doubles);
int kb_peer_buffer_create(long long bytes, void** ptr, unsigned char* handle64);   /* zero-filled; handle64: 64 bytes out */
int kb_peer_buffer_open(const unsigned char* handle64, void** ptr);                /* map a peer's buffer */
int kb_peer_buffer_close(void* ptr);
int kb_peer_buffer_destroy(void* ptr);
int kb_peer_status_create(unsigned long long** word);   /* pinned, device-mapped host word, zeroed */
int kb_peer_status_destroy(unsigned long long* word);
/* buf (n doubles, device) <- sum over ranks, in rank order (bit-identical on every rank), stream-ordered */
int kb_peer_allreduce_f64(void* buf, long long n, const kb_peer_ctx* ctx, unsigned long long seq, kb_stream_t stream);
/* test helper: all `world` ranks of one exchange emulated on ONE GPU as a single cooperative launch (block r = rank r on
 * bufs[r] / ctxs[r]; the contexts' buffers are plain device allocations of this process); advances every ctx->seq */
int kb_peer_allreduce_emulate(void* const* bufs, long long n, kb_peer_ctx* const* ctxs, int world, kb_stream_t stream);
/* a kb_allreduce_hook: user = kb_peer_ctx* */
int kb_peer_allreduce_hook(void* user, void* buf, long long n_doubles, kb_stream_t stream);
/* Gradient-bucket hook of the backward schedules (replaces DDP's bucketed overlap, katago_loop.py:498-504). When set (per
 * host thread; NULL clears), kb_seresnet_backward[_sync] calls it right after the kernels producing the gradients of
 * parameters [first_param, first_param + n_params) have been enqueued: once for the heads (bucket = num_blocks), then for
 * every residual block from the last (bucket = num_blocks - 1) down to 0. The stem (parameters 0..2) finishes with the
 * call itself. `main_stream` / `side_stream` are the two streams those kernels were enqueued on (record an event on
 * both). A non-zero return aborts the backward with an error. */
typedef int (*kb_bucket_hook)(void* user, int bucket, int first_param, int n_params, kb_stream_t main_stream, kb_stream_t side_stream);
int kb_seresnet_set_bucket_hook(kb_bucket_hook hook, void* user);
/* grads: float32 buffers shaped like params, PRE-ZEROED by the caller */
int kb_seresnet_backward(const kb_seresnet_desc* d, const void* const* params, const void* wpack, int B,
                         int dtype, void* workspace, long long ws_bytes, const void* dpolicy,
                         long long policy_pitch, const float* dvalue, const float* dscore,
                         void* const* grads, int use_tc, int num_sms, kb_stream_t stream);

/* ---- plain ResNet baseline: keisei/training/models/resnet.py:25-84 ResidualBlock / ResNetModel.forward
 *      (the `resnet` registry entry, scalar value contract; BASELINE.json configs[3]) and its autograd ----
 * params[15 + 6*num_layers] float32 and buffers[9 + 6*num_layers] in the reference's registration order
 * (see keisei_b200/csrc/resnet.cu header). policy_out: (B, policy_pitch >= 11259) flat logits in the
 * activation dtype; value_out (B,1) float32, tanh-activated. new_stats: [2*num_layers+3][2][hidden_size]
 * float32 (same convention as kb_seresnet_forward). */
typedef struct {
  int num_layers, hidden_size, obs_channels;
} kb_resnet_desc;

long long kb_resnet_num_params(const kb_resnet_desc* d);
long long kb_resnet_num_buffers(const kb_resnet_desc* d);
long long kb_resnet_wpack_bytes(const kb_resnet_desc* d, int dtype);
long long kb_resnet_workspace_bytes(const kb_resnet_desc* d, int B, int training, int dtype);
int kb_resnet_pack_weights(const kb_resnet_desc* d, const void* const* params, const void* const* buffers, int dtype,
                           void* wpack, long long wpack_bytes, kb_stream_t stream);
int kb_resnet_forward(const kb_resnet_desc* d, const void* const* params, void* const* buffers, float* new_stats,
                      const void* wpack, const float* obs, int B, int training, int dtype, void* workspace,
                      long long ws_bytes, void* policy_out, long long policy_pitch, float* value_out, int use_tc,
                      int num_sms, kb_stream_t stream);
/* Gradient-bucket hook of the backward schedules (replaces DDP's bucketed overlap, katago_loop.py:498-504). When set (per
 * host thread; NULL clears), kb_seresnet_backward[_sync] calls it right after the kernels producing the gradients of
 * parameters [first_param, first_param + n_params) have been enqueued: once for the heads (bucket = num_blocks), then for
 * every residual block from the last (bucket = num_blocks - 1) down to 0. The stem (parameters 0..2) finishes with the
 * call itself. `main_stream` / `side_stream` are the two streams those kernels were enqueued on (record an event on
 * both). A non-zero return aborts the backward with an error. */
typedef int (*kb_bucket_hook)(void* user, int bucket, int first_param, int n_params, kb_stream_t main_stream, kb_stream_t side_stream);
int kb_seresnet_set_bucket_hook(kb_bucket_hook hook, void* user);
/* grads: float32 buffers shaped like params, PRE-ZEROED by the caller */
int kb_resnet_backward(const kb_resnet_desc* d, const void* const* params, const void* wpack, int B, int dtype,
                       void* workspace, long long ws_bytes, const void* dpolicy, long long policy_pitch,
                       const float* dvalue, void* const* grads, int use_tc, int num_sms, kb_stream_t stream);

/* ---- single 3x3 convolution on NHWC 9x9 boards (unit-test / profiling entry points):
 *      F.conv2d(padding=1, bias=False) at se_resnet.py:50,52,110 ----
 * in (B,81,Cin), w (Cout,9,Cin) packed, out (B,81,Cout), all `dtype`. backend 0 = SIMT fp32-accumulate,
 * 1 = tcgen05 (bf16 only; kernel chosen automatically), 2 = tcgen05 single-CTA kernel, 3 = tcgen05 CTA-pair kernel
 * (cta_group::2, Cout %% 256 == 0). Optional fused epilogue pieces (null = off): per-channel scale/shift,
 * relu, per-(board,channel) bias, channel sums (double[2*Cout]: sum, sum of squares),
 * board_mean (B,Cout), pool (B,3*Cout: mean,max,std). */
int kb_conv3x3_forward(const void* in, const void* w, void* out, int B, int Cin, int Cout, int dtype, int backend,
                       const float* scale, const float* shift, int relu, const float* gbias, double* ch_sums,
                       float* board_mean, float* pool, int num_sms, kb_stream_t stream);
/* conv2 of a GlobalPoolBiasBlock in evaluation mode with the rest of the block fused into the convolution's epilogue
 * (reference se_resnet.py:79-90): out = relu(bn2(conv(in)) * sigmoid(se_scale) + se_shift + res), where
 * (se_scale, se_shift) = se_fc2(relu(se_fc1(board mean of bn2(conv(in))))); pool (B,3*Cout) / pool_bf16 receive the
 * global-pool statistics (mean, max, population std) of `out` (se_resnet.py:93-98). tcgen05 CTA-pair kernel:
 * bf16, Cout == 256, S == 16, Cin %% 64 == 0, B >= 3. scale/shift = folded eval BatchNorm. */
int kb_conv3x3_se_tail(const void* in, const void* w, void* out, int B, int Cin, int Cout, const float* scale,
                       const float* shift, const void* res, const float* se_w1, const float* se_b1, const float* se_w2,
                       const float* se_b2, int S, float* pool, void* pool_bf16, int num_sms, kb_stream_t stream);
/* dw (Cout,Cin_true,3,3) float32 += sum over boards/pixels of dy (B,81,Cout) x shifted x (B,81,Cin) */
/* backend 1 (tcgen05) needs a scratch buffer of kb_conv3x3_wgrad_ws_bytes() bytes for the per-slice partial tiles */
long long kb_conv3x3_wgrad_ws_bytes(int Cin, int Cout, int num_sms);
int kb_conv3x3_wgrad(const void* x, const void* dy, float* dw, int B, int Cin, int Cout, int Cin_true, int dtype,
                     int backend, void* ws, long long ws_bytes, int num_sms, kb_stream_t stream);
/* w (Cout,Cin,3,3) float32 -> wf (Cout,9,Cinp) and optional wd (Cinp,9,Cout) (flipped taps) in `dtype` */
int kb_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cinp, int dtype, kb_stream_t stream);

/* ---- tail of a GlobalPoolBiasBlock as one kernel (unit-test / profiling entry point): se_resnet.py:83-90 + the
 *      next block's _global_pool (:93-98). bf16 NHWC tiles staged by TMA bulk copies; 64 <= C <= 256, C % 8 == 0.
 * se_in = board_mean * bn_a + bn_b (bn_a/bn_b null: identity); se = W2 relu(W1 se_in + b1) + b2 (w1 (S,C), w2 (2C,S));
 * out = relu((z*bn_a+bn_b) * sigmoid(se[:C]) + se[C:] + res); pool (B,3C) = mean,max,std of out; ties (B,C) optional.
 * se_out (B,2C) is required (kernel-internal hand-off): raw logits when se_raw != 0, else its first half holds sigmoid(scale). */
int kb_se_block_tail(const void* z, const void* res, void* out, const float* bn_a, const float* bn_b,
                     const float* board_mean, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* se_in_out, float* seh_out, float* se_out, int se_raw, float* pool, void* pool_bf16,
                     float* ties, int B, int C, int S, int num_sms, kb_stream_t stream);
/* Same, with the implementation chosen explicitly: 0 = the library default, 1 = one persistent kernel staging board tiles
 * with TMA bulk copies (csrc/se_apply.cu), 2 = SE-MLP kernel + column-layout streaming pass (csrc/se_apply_col.cu). The
 * two sum in different orders (results agree to rounding, not bit for bit), so the network schedule uses ONE of them. */
int kb_se_block_tail_variant(const void* z, const void* res, void* out, const float* bn_a, const float* bn_b,
                             const float* board_mean, const float* w1, const float* b1, const float* w2, const float* b2,
                             float* se_in_out, float* seh_out, float* se_out, int se_raw, float* pool, void* pool_bf16,
                             float* ties, int B, int C, int S, int variant, int num_sms, kb_stream_t stream);

/* ---- Linear / 1x1-conv layer on the tcgen05 path (nn.Linear at se_resnet.py:57-66, heads :119-130) ----
 * w (N,K) float32 -> bf16 (Np,Kp) zero padded, Np % 128 == 0, Kp % 64 == 0. */
int kb_pack_linear_weight(const float* w, void* out_bf16, int N, int K, int Np, int Kp, kb_stream_t stream);
/* Y[m][n] = act((X[m][:] . W[n][:]) * scale[n] + bias[n]); X bf16 (M,Kp); out_f32 (M,ld_f) features < N and/or
 * out_bf16 (M,ld_b) features < nb_store (zeros for n >= N); group_rows > 0: bf16 row m lives at
 * (m / group_rows) * group_pitch + (m % group_rows) * ld_b (the padded policy-logit buffer). */
int kb_linear_tc(const void* x_bf16, long long M, int Kp, const void* w_bf16, int N, int Np, const float* scale,
                 const float* bias, int relu, float* out_f32, long long ld_f, void* out_bf16, long long ld_b,
                 int nb_store, int group_rows, long long group_pitch, int num_sms, kb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KEISEI_B200_H */
