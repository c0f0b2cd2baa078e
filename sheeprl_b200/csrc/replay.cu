// Device-resident replay storage: row gather (sequence sampling) + ring write, and the PPO GAE scan.
//
// Replaces (reference): SequentialReplayBuffer._get_samples' np.take + reshape + swapaxes + H2D copy
// (sheeprl/data/buffers.py:467-526, :1158-1180), ReplayBuffer._get_samples (buffers.py:270-288),
// ReplayBuffer.add ring write (buffers.py:145-221), and `gae` (sheeprl/utils/utils.py:63-100).
// Index generation stays on the host (numpy PCG64 Generator, bit-exact with the reference); these kernels
// only move bytes: algorithmic traffic = 2 x row_bytes per sampled row, coalesced 16-byte accesses.
#include "common.cuh"

namespace {

// out[(s*T + t)*B + b, :] = storage[idx[(s*B + b)*T + t], :]   (rows of row_bytes bytes)
// One warp per output row when rows are large; 16-byte vector path when row_bytes % 16 == 0.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const unsigned char* __restrict__ storage, const long long* __restrict__ idx,
                   unsigned char* __restrict__ out, long long n_rows, int S, int B, int T, long long row_bytes) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_rows; r += nwarps) {
    // r enumerates output rows in (s, t, b) order
    const int b = (int)(r % B);
    const long long q = r / B;
    const int t = (int)(q % T);
    const int s = (int)(q / T);
    const long long src_row = idx[((long long)s * B + b) * T + t];
    const unsigned char* src = storage + src_row * row_bytes;
    unsigned char* dst = out + r * row_bytes;
    if ((row_bytes & 15) == 0) {
      const int4* s4 = reinterpret_cast<const int4*>(src);
      int4* d4 = reinterpret_cast<int4*>(dst);
      const long long n16 = row_bytes >> 4;
      for (long long i = lane; i < n16; i += 32) d4[i] = __ldg(s4 + i);
    } else {
      for (long long i = lane; i < row_bytes; i += 32) dst[i] = src[i];
    }
  }
}

// storage[dst_rows[i], :] = src[i, :]
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ dst_rows,
                    unsigned char* __restrict__ storage, long long n_rows, long long row_bytes) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_rows; r += nwarps) {
    const unsigned char* s = src + r * row_bytes;
    unsigned char* d = storage + dst_rows[r] * row_bytes;
    if ((row_bytes & 15) == 0) {
      const int4* s4 = reinterpret_cast<const int4*>(s);
      int4* d4 = reinterpret_cast<int4*>(d);
      for (long long i = lane; i < (row_bytes >> 4); i += 32) d4[i] = s4[i];
    } else {
      for (long long i = lane; i < row_bytes; i += 32) d[i] = s[i];
    }
  }
}

// thread per environment column; reverse scan over T (sheeprl/utils/utils.py:63-100)
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                           const float* __restrict__ dones, const float* __restrict__ next_value,
                           float* __restrict__ returns, float* __restrict__ advantages, int T, int E, float gamma,
                           float lmbda) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  // The recurrence is a dependent chain but its inputs are not: load CH steps' rewards / values / dones first (CH
  // independent L2 requests in flight per thread), then run the chain from registers.  Same arithmetic as before.
  constexpr int CH = 16;
  float lastgaelam = 0.f;
  float nextvalues = next_value[e];
  for (int t1 = T; t1 > 0; t1 -= CH) {
    float r[CH], v[CH], d[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int t = t1 - 1 - k;
      if (t >= 0) {
        const long long i = (long long)t * E + e;
        r[k] = rewards[i];
        v[k] = values[i];
        d[k] = dones[i];
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int t = t1 - 1 - k;
      if (t >= 0) {
        const long long i = (long long)t * E + e;
        const float nextnonterminal = (d[k] != 0.f) ? 0.f : 1.f;     // dones[t] masks the bootstrap from step t+1
        const float delta = r[k] + nextvalues * nextnonterminal * gamma - v[k];
        lastgaelam = delta + nextnonterminal * lastgaelam * gamma * lmbda;
        advantages[i] = lastgaelam;
        returns[i] = lastgaelam + v[k];
        nextvalues = v[k];
      }
    }
  }
}

}  // namespace

extern "C" int b200rl_replay_gather(const void* storage, const long long* idx, void* out, int n_samples, int batch,
                                    int seq_len, long long row_bytes, cudaStream_t st) {
  RL_CHECK_ARG(storage && idx && out, "null pointer");
  RL_CHECK_ARG(n_samples > 0 && batch > 0 && seq_len > 0 && row_bytes > 0, "bad dims");
  const long long n_rows = (long long)n_samples * batch * seq_len;
  long long blocks = (n_rows + 7) / 8;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const unsigned char*)storage, idx, (unsigned char*)out, n_rows,
                                                       n_samples, batch, seq_len, row_bytes);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

extern "C" int b200rl_replay_scatter(const void* src, const long long* dst_rows, void* storage, long long n_rows,
                                     long long row_bytes, cudaStream_t st) {
  RL_CHECK_ARG(storage && dst_rows && src, "null pointer");
  if (n_rows <= 0) return B200RL_OK;
  long long blocks = (n_rows + 7) / 8;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  scatter_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const unsigned char*)src, dst_rows, (unsigned char*)storage,
                                                        n_rows, row_bytes);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

extern "C" int b200rl_gae(const float* rewards, const float* values, const float* dones, const float* next_value,
                          float* returns, float* advantages, int T, int E, float gamma, float lmbda, cudaStream_t st) {
  RL_CHECK_ARG(rewards && values && dones && next_value && returns && advantages, "null pointer");
  RL_CHECK_ARG(T > 0 && E > 0, "bad dims");
  gae_kernel<<<ceil_div(E, 128), 128, 0, st>>>(rewards, values, dones, next_value, returns, advantages, T, E, gamma,
                                               lmbda);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}
