/* b200rl — C-ABI of the B200 (sm_100a) kernels behind SheepRL's Dreamer-V3 / PPO / SAC update paths.
 *
 * The reference (Eclectic-Sheep/sheeprl) is pure Python and has NO native interface for this path
 * (SURVEY.md §0 F1, §2.2); these entry points are what a reference-side binding (ctypes / a
 * torch.utils.cpp_extension shim) calls from `sheeprl.algos.<algo>.train()`.
 * Each declaration cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer owned by the caller (borrowed);
 *     nothing is allocated or freed by the library; no host synchronisation; all work is enqueued on
 *     `stream` (pass the caller's current stream) — hence CUDA-graph capturable;
 *   - fp32, row-major; `ld*` = row stride in elements of a 2-D view whose inner stride is 1;
 *   - return 0 on success; non-zero => b200rl_last_error() (thread-local message);
 *   - built for sm_100a only (b200rl_device_check refuses other devices). There is no CPU fallback.
 */
#ifndef B200RL_H_
#define B200RL_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* b200rl_last_error(void);
int b200rl_abi_version(void);
const char* b200rl_build_arch(void);
int b200rl_device_check(void);

/* ---- dense layers ------------------------------------------------------------------------------
 * nn.Linear forward/backward everywhere in sheeprl/models/models.py:16-119 (MLP), agent.py:281-341
 * (RecurrentModel), agent.py:1021-1051 (representation / transition).  C[M,N] = op(A) op(B) (+bias) (+C).
 * A is [M,K] (lda) or [K,M] if transA; B is [K,N] (ldb) or [N,K] if transB (nn.Linear weight layout). */
int b200rl_gemm_f32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                    int ldc, int transA, int transB, int accumulate, cudaStream_t stream);
/* Tensor-core path used by b200rl_gemm_f32 for large NT products (transA = 0, transB = 1, 16-byte aligned
 * operands): 3xTF32 split-precision on tcgen05.mma with TMA-fed 128B-swizzled tiles and TMEM accumulators
 * (gemm_tc.cu).  Same contract as b200rl_gemm_f32; `_supported` tells whether a shape is eligible. */
/* Precision of every tensor-core product (GEMM, conv forward / input gradient / weight gradient), process-wide like
 * torch.set_float32_matmul_precision (the reference sets it from configs/config.yaml:18, default "high"):
 * 3 = three TF32 products per k-step on x = hi + lo (fp32-accurate, default; the 1e-4 parity tests run this),
 * 1 = one TF32 product ("high": the numerics of the reference's own GPU runs, ~3x the throughput). */
int b200rl_set_matmul_precision(int tf32_passes);
int b200rl_get_matmul_precision(void);
int b200rl_gemm_tc_supported(const float* A, const float* B, int M, int N, int K, int lda, int ldb, int transA,
                             int transB);
int b200rl_gemm_tc(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                   int ldc, int transA, int transB, int accumulate, cudaStream_t stream);
/* Dense block Linear(bias=False) -> LayerNorm(eps) -> activation in two launches (sheeprl/utils/model.py:34-88 miniblock,
 * as built by MLP at sheeprl/models/models.py:23-119): the tcgen05 product leaves split-K partial tiles, one kernel sums
 * them in split order, normalises the row in registers and applies the activation.  `pre` (optional) keeps W.x for the
 * backward.  mode 1 (N = 3R): the tail is LayerNormGRUCell's gate instead (sheeprl/models/models.py:396-403): h_out (and
 * h_out2, optional second copy, e.g. the next step's [h | x] input) = u * tanh(r * c) + (1 - u) * h_prev; `out` (optional)
 * keeps LN(W.x).  A [M, K] (lda), W [N, K] (ldw): y = A W^T.  `_supported`: 16-byte aligned NT operands, N % 4 == 0,
 * N <= 1536 (mode 1: N % 384 == 0). */
int b200rl_gemm_ln_supported(const float* A, const float* W, int M, int N, int K, int lda, int ldw, int mode);
int b200rl_gemm_ln(const float* A, const float* W, int M, int N, int K, int lda, int ldw, const float* gamma,
                   const float* beta, float eps, int act, float* pre, long long ldpre, float* out, long long ldout,
                   int mode, const float* h_prev, long long ldh, float* h_out, long long ldho, float* h_out2,
                   long long ldho2, cudaStream_t stream);
/* nn.LayerNorm(eps) (+ nn.SiLU): miniblock sheeprl/utils/model.py:34-88; LayerNormChannelLast
 * sheeprl/models/models.py:507-518 (channel-last is native here).  act: 0 none, 1 SiLU, 2 tanh, 3 ReLU (the last two:
 * PPO MLPs with layer_norm=True, sheeprl/algos/ppo/agent.py:58-66,152-176). */
int b200rl_ln_act_fwd(const float* X, const float* gamma, const float* beta, float* Y, long long M, int C,
                      long long ldx, long long ldy, float eps, int act, cudaStream_t stream);
int b200rl_ln_act_bwd(const float* X, const float* gamma, const float* beta, const float* dY, float* dX, float* dgamma,
                      float* dbeta, long long M, int C, long long ldx, long long lddy, long long lddx, float eps,
                      int act, int accumulate, cudaStream_t stream);
int b200rl_col_sum(const float* X, float* out, long long M, int C, long long ldx, int accumulate, cudaStream_t stream);

/* ---- image encoder / decoder ---------------------------------------------------------------------
 * CNNEncoder agent.py:42-97 (Conv2d k4 s2 p1), CNNDecoder agent.py:154-226 (ConvTranspose2d k4 s2 p1),
 * observation normalisation dreamer_v3.py:98.  Images are NHWC; W is [C_small, C_big, 4, 4] (the
 * reference layouts of both layer kinds).  (h, w) = SMALL image size; big image is (2h, 2w). */
int b200rl_obs_prep(const void* obs_nchw, int is_uint8, float* out_nhwc, long long NB, int C, int HW,
                    cudaStream_t stream);
int b200rl_transpose_batched(const float* X, float* Y, int NB, int a, int b, cudaStream_t stream);
/* Y[j][i] = X[i][j] for a strided [rows][cols] view (used to present K-major operands to the tensor-core GEMM:
 * W^T for input gradients, dY^T / X^T for weight gradients). */
int b200rl_transpose2d(const float* X, float* Y, int rows, int cols, long long ldx, long long ldy, cudaStream_t stream);
int b200rl_conv_down(const float* big, const float* W, float* small_, int NB, int h, int w, int Cs, int Cb,
                     cudaStream_t stream);
int b200rl_conv_up(const float* small_, const float* W, float* big, const float* bias, int NB, int h, int w, int Cs,
                   int Cb, cudaStream_t stream);
int b200rl_conv_wgrad(const float* small_, const float* big, float* dW, int NB, int h, int w, int Cs, int Cb,
                      int accumulate, cudaStream_t stream);
/* Tensor-core implicit-GEMM versions of conv_down / conv_up (gemm_tc.cu: 4-D TMA boxes gather the taps, no
 * im2col buffer, 3xTF32 tcgen05).  `Wpacked` is a caller-owned 16*Cs*Cb-float workspace filled by
 * b200rl_conv_pack (down: [Cs][tap][Cb]; up: [parity][Cb][tap][Cs]) after every weight update.  Eligible when the
 * gathered image has a multiple of 32 channels and the small grid tiles by 128 pixels (`_supported`). */
/* Weight gradient as one tensor-core GEMM over all pixels: small^T [Cs][P] times the transposed im2col of `big`
 * [16*Cb][P] (both K-major, split-K).  `workspace`: b200rl_conv_wgrad_tc_workspace(...) floats, caller-owned. */
long long b200rl_conv_wgrad_tc_workspace(int NB, int h, int w, int Cs, int Cb);
int b200rl_conv_wgrad_tc(const float* small_, const float* big, float* dW, float* workspace, int NB, int h, int w,
                         int Cs, int Cb, int accumulate, cudaStream_t stream);
int b200rl_conv_tc_supported(int mode_up, int NB, int h, int w, int Cs, int Cb);
/* floats of workspace b200rl_conv_pack writes (16*Cs*Cb; 36*Cs*Cb for ConvTranspose2d layers with 32 output channels, whose
 * four output-parity classes are computed as one 128-column tile over the 9 shifted input windows) */
long long b200rl_conv_pack_floats(int mode_up, int Cs, int Cb);
int b200rl_conv_pack(const float* W, float* Wpacked, int mode_up, int Cs, int Cb, cudaStream_t stream);
int b200rl_conv_down_tc(const float* big, const float* Wpacked, float* small_, int NB, int h, int w, int Cs, int Cb,
                        cudaStream_t stream);
int b200rl_conv_up_tc(const float* small_, const float* Wpacked, float* big, const float* bias, int NB, int h, int w,
                      int Cs, int Cb, cudaStream_t stream);

/* ---- RSSM ------------------------------------------------------------------------------------------
 * LayerNormGRUCell gates models.py:399-403; is_first masking agent.py:425-430; unimix agent.py:437-449;
 * straight-through categorical sampling dreamer_v2/utils.py:44-61 (noise q ~ Exp(1): sample =
 * argmax(probs / q), noise == NULL -> mode); KL balancing + free nats loss.py:61-75. */
int b200rl_gru_gate_fwd(const float* G, const float* Hin, float* Hout, long long M, int R, long long ldg,
                        long long ldhi, long long ldho, cudaStream_t stream);
int b200rl_gru_gate_bwd(const float* G, const float* Hin, const float* dH, float* dG, float* dHin, long long M, int R,
                        long long ldg, long long ldhi, long long lddh, long long lddg, long long lddhi,
                        cudaStream_t stream);
int b200rl_mask_mix(const float* prev, const float* init_row, const float* first, float* out, long long M, int C,
                    long long ldp, long long ldo, cudaStream_t stream);
int b200rl_mask_bwd(const float* dIn, const float* first, float* dPrev, float* dInit, int M, int C, long long ldi,
                    long long ldp, cudaStream_t stream);
int b200rl_cat_sample(const float* raw, const float* noise, float* onehot, float* mix_out, long long M, int groups,
                      int classes, long long ldr, long long ldn, long long ldo, long long ldm, float unimix,
                      cudaStream_t stream);
/* Policy head + straight-through sample in one launch (Actor.mlp_heads[i] + OneHotCategoricalStraightThrough.rsample,
 * sheeprl/algos/dreamer_v3/agent.py:793-818): raw [M, A] = X W^T + bias, onehot = sample(unimix(raw), noise) with
 * b200rl_cat_sample's rule.  A <= 32, Kin <= 1024, Kin % 4 == 0, 16-byte aligned rows. */
int b200rl_head_sample(const float* X, const float* W, const float* bias, const float* noise, float* raw, float* onehot,
                       long long M, int Kin, int A, long long ldx, long long ldw, long long ldr, long long ldn,
                       long long ldo, float unimix, cudaStream_t stream);
int b200rl_cat_sample_bwd(const float* raw, const float* dz, const float* dmix, float* draw, long long M, int groups,
                          int classes, long long ldr, long long lddz, long long lddm, long long lddr, float unimix,
                          cudaStream_t stream);
int b200rl_kl_loss_grad(const float* post_mix, const float* prior_mix, float* d_post, float* d_prior, float* rows,
                        long long M, int groups, int classes, long long ldp, long long ldq, long long lddp,
                        long long lddq, float kl_dyn, float kl_rep, float free_nats, float regularizer, float scale,
                        cudaStream_t stream);
/* Persistent fused scan: all T steps of the POSTERIOR recurrence of RSSM.dynamic (dreamer_v3.py:131-145 ->
 * agent.py:396-435: recurrent model, representation model, unimix, straight-through sample) in ONE cooperative
 * kernel; weight slices resident in shared memory, cross-SM hand-offs as flag-carrying data exchanges instead of grid
 * barriers (rssm_scan.cu).  The prior (transition model on the finished h sequence, agent.py:433) is NOT computed here:
 * it is off the recurrence and runs as batched products over all T*B rows (tr_pre / tr_act / prior_raw / prior_mix
 * below are unused by the kernels; the fields stay so that the per-step path and this one fill the same set of saved
 * activations).  Requires B <= 16, classes <= 32, even layer widths and the per-CTA weight slices to fit in 227 KB
 * (returns non-zero otherwise; the caller then uses the per-step ops). */
typedef struct b200rl_rssm_scan_args {
  int T, B, S, D, R, A, Dx, Dt, Dr, ld_lat, ld_wr1;
  float eps, unimix;
  const float *W_in, *lnx_g, *lnx_b, *W_g, *lng_g, *lng_b, *W_t1, *lnt_g, *lnt_b, *W_t2, *b_t2;
  const float *W_r1, *lnr_g, *lnr_b, *W_r2, *b_r2;
  const float *h0, *z0;                       /* [R] tanh(initial_recurrent_state); [S*D] one-hot initial posterior */
  const float *pe, *actions, *first, *noise;  /* [T,B,Dr] embed projection; [T,B,A] shifted; [T,B]; [T,B,S*D] Exp(1) */
  float* latent;                              /* [T*B, ld_lat]: z (S*D) | h (R) */
  float *z_in, *h_in, *a_in, *x_pre, *x_act, *g_pre, *g_ln, *tr_pre, *tr_act, *rp_pre, *rp_act;
  float *post_raw, *prior_raw, *post_mix, *prior_mix;
  void* workspace;
  long long workspace_bytes;
} b200rl_rssm_scan_args;
/* Gradient buffers of the persistent BPTT kernel (same meaning as the per-step path's buffers):
 * inputs d_latent [T*B, ld_lat] (grad wrt z|h from decoder + heads + the batched prior backward), d_post_mix (KL seed
 * grads); outputs: d_post_raw and the per-step ACTIVATION gradients d_rp_act / d_g_ln / d_x_act (the batched
 * LayerNorm-backward kernels turn them into d_rp_pre / d_g_pre / d_x_pre afterwards), and d_h0 [R].
 * d_prior_mix / d_prior_raw / d_tr_act / d_tr_pre / d_*_pre are unused by the kernel. */
typedef struct b200rl_rssm_scan_grads {
  const float *d_latent, *d_post_mix, *d_prior_mix;
  float *d_post_raw, *d_prior_raw, *d_rp_act, *d_rp_pre, *d_tr_act, *d_tr_pre, *d_g_ln, *d_g_pre, *d_x_act, *d_x_pre;
  float* d_h0;
  /* pre-activation x weight products over all T*B rows (batched, before the kernel; they let the consumer of a
   * LayerNorm gradient apply the LayerNorm-backward correction by linearity, rssm_scan.cu):
   * q_r = rp_pre W_r1[:, :R] [T*B, R];  q_g = g_pre W_g [T*B, R+Dx];  q_x = x_pre W_in[:, :S*D] [T*B, S*D] */
  const float *q_r, *q_g, *q_x;
} b200rl_rssm_scan_grads;
long long b200rl_rssm_scan_workspace_bytes(int T, int B, int S, int D, int Dx, int R, int Dr);
int b200rl_rssm_scan_fwd(const b200rl_rssm_scan_args* args, cudaStream_t stream);
/* BPTT over the same scan (autograd replay inside fabric.backward, dreamer_v3.py:191); must follow
 * b200rl_rssm_scan_fwd on the same workspace (uses its saved LayerNorm statistics and class indices). */
int b200rl_rssm_scan_bwd(const b200rl_rssm_scan_args* args, const b200rl_rssm_scan_grads* grads, cudaStream_t stream);
/* envelope check of the backward kernel for `args` (non-zero + last_error if it cannot run); launches nothing */
int b200rl_rssm_scan_bwd_check(const b200rl_rssm_scan_args* args);
int b200rl_rssm_scan_error(const void* workspace, cudaStream_t stream);
/* cycle counters (2 x 32 int64: CTA 0, CTA 1) accumulated per phase by the last launch on `workspace` */
int b200rl_rssm_scan_profile(const void* workspace, long long* out64, cudaStream_t stream);

/* ---- losses (value + seed gradient) -----------------------------------------------------------------
 * distribution.py:212-276 (MSE, two-hot on symlog), Bernoulli continue head loss.py:77, lambda returns
 * dreamer_v3/utils.py:66-77 + dreamer_v3.py:244-260, Moments dreamer_v3/utils.py:40-63, discrete policy
 * loss dreamer_v3.py:272-297. */
int b200rl_mse_loss_grad(const float* pred, const float* target, float* loss_row, float* grad, long long M, int P,
                         float scale, cudaStream_t stream);
int b200rl_twohot_loss_grad(const float* logits, const float* x, const float* weight, float* loss_row, float* dlogits,
                            long long M, int nbins, long long ldl, long long ldd, float low, float high, float scale,
                            int accumulate, cudaStream_t stream);
int b200rl_bce_loss_grad(const float* logit, const float* target, float* loss_row, float* dlogit, long long M,
                         float loss_scale, float scale, cudaStream_t stream);
int b200rl_twohot_mean(const float* logits, float* out, long long M, int nbins, long long ldl, float low, float high,
                       cudaStream_t stream);
int b200rl_lambda_returns(const float* rew, const float* val, const float* cont_logit, const float* true_cont,
                          float* lam, float* discount, int H, int N, float gamma, float lmbda, cudaStream_t stream);
int b200rl_moments_update(const float* x, long long n, float* state_low_high, float* out_offset_invscale, float decay,
                          float max_, float p_low, float p_high, cudaStream_t stream);
int b200rl_actor_loss_grad(const float* raw, const float* actions, const float* lam, const float* val,
                           const float* discount, const float* moments, float* rows, float* draw, long long M,
                           const int* head_dims_host, int n_heads, float unimix, float ent_coef, float scale,
                           cudaStream_t stream);
int b200rl_sum_rows(const float* X, float* out, long long M, int C, long long ldx, float scale, cudaStream_t stream);
int b200rl_weighted_mean(const float* x, const float* w, float* out, long long n, float scale, cudaStream_t stream);

/* ---- optimiser -----------------------------------------------------------------------------------------
 * clip_grad_norm_ + torch.optim.Adam.step fused (dreamer_v3.py:191-200,298-304,318-327); target-critic
 * EMA dreamer_v3.py:674-680; Exp(1) noise for categorical sampling (torch.multinomial). */
int b200rl_sumsq(const float* x, long long n, double* out, cudaStream_t stream);
int b200rl_adam_step(float* p, const float* g, float* m, float* v, const double* normsq, const int* step_dev,
                     float* norm_out, long long n, float max_norm, float lr, float b1, float b2, float eps,
                     cudaStream_t stream);
int b200rl_ema(float* target, const float* src, long long n, float tau, cudaStream_t stream);
int b200rl_fill_exponential(float* out, long long n, unsigned long long seed, unsigned int stream_id,
                            const int* counter_dev, cudaStream_t stream);
int b200rl_zero(float* x, long long n, cudaStream_t stream);
int b200rl_copy2d(const float* src, float* dst, long long M, int C, long long lds, long long ldd, cudaStream_t stream);
int b200rl_axpy(const float* x, float* y, long long n, float alpha, cudaStream_t stream);
int b200rl_affine(const float* x, float* y, long long n, float alpha, float beta, cudaStream_t stream);
/* y[M,C] = symlog(x[M,C]) = sign(x) log(1 + |x|)  (sheeprl/utils/utils.py:148): the squashing of vector observations in
 * MLPEncoder.forward (dreamer_v3/agent.py:150) and the regression target of SymlogDistribution (utils/distribution.py:180). */
int b200rl_symlog(const float* x, float* y, long long M, int C, long long ldx, long long ldy, cudaStream_t stream);
int b200rl_tanh_fwd(const float* x, float* y, long long n, cudaStream_t stream);
int b200rl_tanh_bwd(const float* y, const float* dy, float* dx, long long n, int accumulate, cudaStream_t stream);
int b200rl_increment(int* p, cudaStream_t stream);

/* ---- replay storage / PPO -------------------------------------------------------------------------------
 * SequentialReplayBuffer._get_samples buffers.py:467-526 (+ get_tensor :1158-1180), ReplayBuffer.add
 * buffers.py:145-221, gae utils/utils.py:63-100.  idx: int64 flat row indices in (sample, batch, time)
 * order exactly as the reference computes them on the host; out rows in (sample, time, batch) order. */
int b200rl_replay_gather(const void* storage, const long long* idx, void* out, int n_samples, int batch, int seq_len,
                         long long row_bytes, cudaStream_t stream);
int b200rl_replay_scatter(const void* src, const long long* dst_rows, void* storage, long long n_rows,
                          long long row_bytes, cudaStream_t stream);
int b200rl_gae(const float* rewards, const float* values, const float* dones, const float* next_value, float* returns,
               float* advantages, int T, int E, float gamma, float lmbda, cudaStream_t stream);

/* ---- SAC / PPO dense layers and SAC element-wise stages ---------------------------------------------------
 * b200rl_bgemm: `nets` independent products C[n] = epi(A[n] * B[n] + bias[n]) in one launch; A(m,k) = A[m*sam+k*sak],
 * B(k,n) = B[k*sbk+n*sbn] (covers NN/NT/TN), C row-major with ldc.  epilogue: 0 none, 1 ReLU, 2 Tanh, 3 multiply by
 * ReLU'(aux), 4 multiply by Tanh'(aux) = 1-aux^2 (aux = the layer's saved activation output).  rsum (optional):
 * rsum[m] = sum_k A(m,k), i.e. the bias gradient when A = dY^T.  Replaces nn.Linear + activation forward/backward
 * of models/models.py:16-119 as used by sac/agent.py:19-108 and ppo/agent.py:84-177. */
int b200rl_bgemm(const float* A, long long sam, long long sak, long long strideA, const float* B, long long sbk,
                 long long sbn, long long strideB, float* C, long long ldc, long long strideC, const float* bias,
                 long long strideBias, const float* aux, long long ldaux, long long strideAux, float* rsum,
                 long long strideRsum, int M, int N, int K, int nets, int epilogue, int accumulate, cudaStream_t stream);
/* SACActor._get_actions_and_log_probs sac/agent.py:110-142: head = [mean | log_std] (B x 2A), eps ~ N(0,1);
 * writes the rescaled tanh action into `action` (row stride ld_action: straight into the critics' input buffer),
 * logp[B], and tanh(x_t) for the backward. */
int b200rl_sac_sample_fwd(const float* head, const float* eps, const float* scale, const float* abias, float* action,
                          long long ld_action, float* logp, float* tanh_out, int B, int A, cudaStream_t stream);
/* autograd of the above for the policy loss (loss.py:9-11): d(action) = sum over `nets` critics' input gradients,
 * d(logp) = exp(log_alpha)/B. */
int b200rl_sac_sample_bwd(const float* head, const float* eps, const float* tanh_y, const float* scale,
                          const float* dact, long long stride_net, int nets, const float* log_alpha, float* dhead, int B,
                          int A, cudaStream_t stream);
/* SACAgent.get_next_target_q_values sac/agent.py:254-262 (alpha read from the device log_alpha) */
int b200rl_sac_target(const float* q_target, long long stride_net, int nets, const float* logp, const float* rewards,
                      const float* terminated, const float* log_alpha, float gamma, float* y, int B, cudaStream_t stream);
/* critic_loss sac/loss.py:14-20 + its gradient w.r.t. q */
int b200rl_sac_critic_loss(const float* q, long long stride_net, int nets, const float* y, float* dq, float* loss_out,
                           int B, cudaStream_t stream);
/* policy_loss + entropy_loss sac/loss.py:9-11,23-26 with torch.min over critics (sac.py:62-63): gradients w.r.t. q
 * and log_alpha */
int b200rl_sac_actor_loss(const float* q, long long stride_net, int nets, const float* logp, const float* log_alpha,
                          float target_entropy, float* dq, float* actor_loss, float* alpha_loss, float* dlog_alpha, int B,
                          cudaStream_t stream);
/* N(0,1) noise (Philox4x32-10 + Box-Muller), same counter scheme as b200rl_fill_exponential; replaces
 * Normal.rsample's torch.normal draw (sac/agent.py:126) */
int b200rl_fill_normal(float* out, long long n, unsigned long long seed, unsigned int stream_id, const int* counter_dev,
                       cudaStream_t stream);

/* ---- PPO --------------------------------------------------------------------------------------------------
 * Channel-last patch gather / scatter for NatureCNN's unpadded convolutions (models/models.py:288-328):
 * col[(b,oy,ox),(ky,kx,c)] = x[b,oy*s+ky,ox*s+kx,c]; col2im is its transpose (sum over overlapping patches), optionally
 * masked by ReLU'(act) of the activation that fed the convolution. */
int b200rl_im2col(const float* x, float* col, int B, int H, int W, int C, int k, int stride, cudaStream_t stream);
int b200rl_col2im(const float* dcol, const float* act, float* dx, int B, int H, int W, int C, int k, int stride,
                  cudaStream_t stream);
/* PPO objective on one minibatch: log-prob + entropy of the taken actions from the actor head (OneHotCategorical per
 * head / Independent Normal, ppo/agent.py:179-239), optional advantage normalisation (utils/utils.py:121-130),
 * policy / value / entropy losses (ppo/loss.py:6-75, reduction mean) and the gradients of
 * policy + vf_coef*value + ent_coef*entropy w.r.t. the head outputs and the values.  losses[3].
 * is_continuous: 0 discrete, 1 `normal`, 2 `tanh_normal` (stored actions are tanh-squashed, agent.py:194-206). */
int b200rl_ppo_loss(const float* head, const float* actions, const float* old_logp, const float* adv,
                    const float* values, const float* old_values, const float* returns, float* dhead, float* dvalues,
                    float* losses, int B, const int* head_dims, int n_heads, int is_continuous, int clip_vloss,
                    int normalize_adv, float clip_coef, float vf_coef, float ent_coef, cudaStream_t stream);

/* Linear([z, a]) for a one-hot z (S groups of K classes, straight-through categorical sample) as a gather-sum over the
 * transposed weight WT [S*K + A, N]: RecurrentModel.mlp's first Linear (agent.py:328-341) inside the imagination
 * rollout.  z: [M, S*K] (row stride ldz), act: [M, A], out: [M, N]. */
int b200rl_onehot_linear(const float* z, const float* act, const float* WT, float* out, long long M, int S, int K, int A,
                         int N, long long ldz, long long lda, long long ldo, cudaStream_t stream);
/* The same gather followed, in the same launch, by the miniblock's LayerNorm(eps) + SiLU: out = SiLU(LN(Linear([z, a])));
 * `pre` (optional) keeps the Linear output.  N a multiple of 128 up to 1024 (one float4 of the row per thread), 16-byte
 * aligned WT / gamma / beta / out / pre rows. */
int b200rl_onehot_linear_ln(const float* z, const float* act, const float* WT, const float* gamma, const float* beta,
                            float eps, float* pre, long long ldpre, float* out, long long M, int S, int K, int A, int N,
                            long long ldz, long long lda, long long ldo, cudaStream_t stream);

/* ---- Dreamer-V3 continuous actions (policy gradient through the imagined rollout) -------------------------
 * Actor.forward `scaled_normal` branch agent.py:803-825: head = [mean | std_raw] (M x 2A), eps ~ N(0,1);
 * action (row stride lda) = clip-rescaled tanh(mean) + std*eps, ent[M] = Independent(Normal).entropy(). */
int b200rl_cont_action_fwd(const float* head, const float* eps, float* action, long long lda, float* ent, long long M,
                           int A, float min_std, float max_std, float init_std, float clip, cudaStream_t stream);
/* its backward: d_action (row stride ldd) and the entropy bonus d_ent[m] = ent_scale * discount[m] -> dhead */
int b200rl_cont_action_bwd(const float* head, const float* eps, const float* d_action, long long ldd,
                           const float* discount, float* dhead, long long M, int A, float min_std, float max_std,
                           float init_std, float clip, float ent_scale, cudaStream_t stream);
/* continuous objective dreamer_v3.py:276-296 + backward of compute_lambda_values (dreamer_v3/utils.py:66-77):
 * rows[H,N] = discount*(advantage + ent_coef*entropy); d_val / d_rew [H+1,N] = d(policy_loss)/d(values, rewards) */
int b200rl_lambda_returns_bwd(const float* cont_logit, const float* discount, const float* moments, const float* lam,
                              const float* val, const float* ent, float* d_val, float* d_rew, float* rows, int H, int N,
                              float gamma, float lmbda, float ent_coef, float scale, cudaStream_t stream);
/* backward of TwoHotEncodingDistribution.mean (distribution.py:245-247): d_logits = d_mean * dsymexp * softmax' */
int b200rl_twohot_mean_bwd(const float* logits, const float* d_mean, float* d_logits, long long M, int nb, long long ldl,
                           long long ldd, float low, float high, cudaStream_t stream);

/* Convolution weight gradient with both operands read in place as MN-major tcgen05 operands (no im2col, no transposes):
 * G[(tap, cb), cs] = sum over small pixels p of big[patch(p)][tap][cb] * small[p][cs]; big [NB,2h,2w,Cb] and small
 * [NB,h,w,Cs] channel-last, G [16*Cb][Cs] (unpacked to the reference's [Cs][Cb][4][4] by b200rl_conv_wgrad_tc).
 * Replaces the autograd weight-gradient of CNNEncoder / CNNDecoder convolutions (agent.py:78-91, :199-222). */
int b200rl_conv_wgrad_mn_supported(int NB, int h, int w, int Cs, int Cb);
int b200rl_conv_wgrad_mn(const float* small_, const float* big, float* G, int NB, int h, int w, int Cs, int Cb,
                         cudaStream_t stream);

/* PPOPlayer.forward / get_actions (ppo/agent.py:269-322): per-head categorical sample (Exp(1) noise; mode when greedy or
 * noise == NULL) or Normal sample (N(0,1) noise; mean when greedy) from the actor head, its log-probability logp[B];
 * actions: one-hot [B, sum(head_dims)] or [B, A].  is_continuous: 0 discrete, 1 `normal`, 2 `tanh_normal` as
 * forward() returns it (safetanh + corrected log-prob, :257-268), 3 `tanh_normal` as get_actions() returns it (:306-311). */
int b200rl_ppo_act(const float* head, const float* noise, float* actions, float* logp, int B, const int* head_dims,
                   int n_heads, int is_continuous, int greedy, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H_ */
