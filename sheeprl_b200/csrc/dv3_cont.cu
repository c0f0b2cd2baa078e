// Dreamer-V3 with CONTINUOUS actions: the pieces of the policy gradient that flows through the imagined rollout.
//
// Replaces (reference): Actor.forward `scaled_normal` branch sheeprl/algos/dreamer_v3/agent.py:803-825 (+ its autograd
// backward), the continuous objective of train() dreamer_v3.py:276-296 (objective = advantage, entropy bonus,
// discount weighting), the autograd backward of compute_lambda_values dreamer_v3/utils.py:66-77 and of
// TwoHotEncodingDistribution.mean sheeprl/utils/distribution.py:245-247 (symexp of the softmax-weighted bins).
// All are element-wise / per-row kernels on <= (H+1)*N rows: HBM-trivial, one launch each.
#include "common.cuh"

namespace {

constexpr float kHalfLog2PiE = 1.4189385332046727f;      // 0.5 + 0.5*log(2*pi): entropy of a unit Normal

__device__ __forceinline__ float bin_value(int i, int nb, float low, float high) {   // torch.linspace, as losses.cu
  const float step = (high - low) / (float)(nb - 1);
  return (i < nb / 2) ? (low + step * (float)i) : (high - step * (float)(nb - 1 - i));
}

// head [M, 2A] = [mean | std_raw].  std = (max-min)*sigmoid(std_raw + init) + min; a = tanh(mean) + std*eps;
// a *= (clip / max(clip, |a|))  (the factor is detached in the reference); entropy = sum_j 0.5+0.5log(2pi)+log std
__global__ void cont_action_fwd_kernel(const float* __restrict__ head, const float* __restrict__ eps,
                                       float* __restrict__ action, long long lda, float* __restrict__ ent, long long M,
                                       int A, float min_std, float max_std, float init_std, float clip) {
  const int lane = threadIdx.x & 31;
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  float e = 0.f;
  for (int j = lane; j < A; j += 32) {
    const float mean = head[m * 2 * A + j], sr = head[m * 2 * A + A + j];
    const float std = (max_std - min_std) * sigmoidf_(sr + init_std) + min_std;
    float a = tanhf(mean) + std * eps[m * A + j];
    if (clip > 0.f) a = a * (clip / fmaxf(clip, fabsf(a)));
    action[m * lda + j] = a;
    e += kHalfLog2PiE + logf(std);
  }
  e = warp_sum(e);
  if (lane == 0 && ent) ent[m] = e;
}

// d(head) from d(action) and the entropy bonus: d_ent[m] = ent_scale * discount[m]  (rows of the first H steps)
__global__ void cont_action_bwd_kernel(const float* __restrict__ head, const float* __restrict__ eps,
                                       const float* __restrict__ dact, long long ldd, const float* __restrict__ discount,
                                       float* __restrict__ dhead, long long M, int A, float min_std, float max_std,
                                       float init_std, float clip, float ent_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * A) return;
  const long long m = i / A;
  const int j = (int)(i - m * A);
  const float mean = head[m * 2 * A + j], sr = head[m * 2 * A + A + j];
  const float sg = sigmoidf_(sr + init_std);
  const float std = (max_std - min_std) * sg + min_std;
  const float th = tanhf(mean), e = eps[i];
  const float a_raw = th + std * e;
  const float f = (clip > 0.f) ? clip / fmaxf(clip, fabsf(a_raw)) : 1.f;
  const float da = dact[m * ldd + j] * f;
  const float dent = ent_scale * discount[m];
  const float dstd = da * e + dent / std;
  dhead[m * 2 * A + j] = da * (1.f - th * th);
  dhead[m * 2 * A + A + j] = dstd * (max_std - min_std) * sg * (1.f - sg);
}

// thread per column n.  Policy objective rows and the gradient of
//   policy_loss = -scale * sum_{t<H} D_t * ((lam_t - v_t)/invscale + ent_coef * ent_t)
// w.r.t. the predicted rewards / values (through the lambda recursion L_t = r_{t+1} + c_{t+1}((1-l) v_{t+1} + l L_{t+1})).
__global__ void lambda_returns_bwd_kernel(const float* __restrict__ cont_logit, const float* __restrict__ discount,
                                          const float* __restrict__ moments, const float* __restrict__ lam,
                                          const float* __restrict__ val, const float* __restrict__ ent,
                                          float* __restrict__ d_val, float* __restrict__ d_rew, float* __restrict__ rows,
                                          int H, int N, float gamma, float lmbda, float ent_coef, float scale) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float inv = 1.f / moments[1];
  float G = 0.f, c_prev = 0.f;
  d_rew[n] = 0.f;
  float carry_v = 0.f;                      // gradient into v_t coming from L_{t-1}'s (1-lambda) term
  for (int t = 0; t < H; ++t) {
    const long long i = (long long)t * N + n, i1 = (long long)(t + 1) * N + n;
    const float D = discount[i];
    rows[i] = D * ((lam[i] - val[i]) * inv + ent_coef * ent[i]);
    const float dlam = -scale * D * inv;
    G = dlam + c_prev * lmbda * G;                                       // dLoss/dL_t
    const float c = ((sigmoidf_(cont_logit[i1]) > 0.5f) ? 1.f : 0.f) * gamma;   // continues[t+1] * gamma
    d_rew[i1] = G;
    d_val[i] = scale * D * inv + carry_v;                                // baseline v_t (+ from L_{t-1})
    carry_v = G * c * (1.f - lmbda);
    c_prev = c;
  }
  d_val[(long long)H * N + n] = carry_v + c_prev * lmbda * G;            // v_H: (1-lambda) term of L_{H-1} + the seed L_H
}

// V = symexp(m), m = sum_j softmax(l)_j b_j  ->  dl_j = dV * exp(|m|) * p_j * (b_j - m)
__global__ void __launch_bounds__(256)
twohot_mean_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ d_mean, float* __restrict__ d_logits,
                       long long M, int nb, long long ldl, long long ldd, float low, float high) {
  const int lane = threadIdx.x & 31;
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const float* l = logits + m * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < nb; c += 32) mx = fmaxf(mx, l[c]);
  mx = warp_max(mx);
  float se = 0.f, sb = 0.f;
  for (int c = lane; c < nb; c += 32) se += expf(l[c] - mx);
  se = warp_sum(se);
  for (int c = lane; c < nb; c += 32) sb = fmaf(expf(l[c] - mx) / se, bin_value(c, nb, low, high), sb);
  sb = warp_sum(sb);
  const float g = d_mean[m] * expf(fabsf(sb));
  for (int c = lane; c < nb; c += 32)
    d_logits[m * ldd + c] = g * (expf(l[c] - mx) / se) * (bin_value(c, nb, low, high) - sb);
}

}  // namespace

extern "C" int b200rl_cont_action_fwd(const float* head, const float* eps, float* action, long long lda, float* ent,
                                      long long M, int A, float min_std, float max_std, float init_std, float clip,
                                      cudaStream_t st) {
  RL_CHECK_ARG(head && eps && action, "null pointer");
  RL_CHECK_ARG(A > 0 && lda >= A, "bad dims");
  if (M <= 0) return B200RL_OK;
  cont_action_fwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(head, eps, action, lda, ent, M, A, min_std, max_std, init_std, clip);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

extern "C" int b200rl_cont_action_bwd(const float* head, const float* eps, const float* d_action, long long ldd,
                                      const float* discount, float* dhead, long long M, int A, float min_std,
                                      float max_std, float init_std, float clip, float ent_scale, cudaStream_t st) {
  RL_CHECK_ARG(head && eps && d_action && discount && dhead, "null pointer");
  RL_CHECK_ARG(A > 0 && ldd >= A, "bad dims");
  if (M <= 0) return B200RL_OK;
  cont_action_bwd_kernel<<<ceil_div(M * A, 256), 256, 0, st>>>(head, eps, d_action, ldd, discount, dhead, M, A, min_std,
                                                              max_std, init_std, clip, ent_scale);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

extern "C" int b200rl_lambda_returns_bwd(const float* cont_logit, const float* discount, const float* moments,
                                         const float* lam, const float* val, const float* ent, float* d_val, float* d_rew,
                                         float* rows, int H, int N, float gamma, float lmbda, float ent_coef, float scale,
                                         cudaStream_t st) {
  RL_CHECK_ARG(cont_logit && discount && moments && lam && val && ent && d_val && d_rew && rows, "null pointer");
  if (N <= 0 || H <= 0) return B200RL_OK;
  lambda_returns_bwd_kernel<<<ceil_div(N, 128), 128, 0, st>>>(cont_logit, discount, moments, lam, val, ent, d_val, d_rew,
                                                              rows, H, N, gamma, lmbda, ent_coef, scale);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

extern "C" int b200rl_twohot_mean_bwd(const float* logits, const float* d_mean, float* d_logits, long long M, int nb,
                                      long long ldl, long long ldd, float low, float high, cudaStream_t st) {
  RL_CHECK_ARG(logits && d_mean && d_logits, "null pointer");
  RL_CHECK_ARG(nb > 1 && ldl >= nb && ldd >= nb, "bad dims");
  if (M <= 0) return B200RL_OK;
  twohot_mean_bwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(logits, d_mean, d_logits, M, nb, ldl, ldd, low, high);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}
