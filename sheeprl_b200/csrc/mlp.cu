// Batched strided small-GEMM with fused epilogues: the Linear layers of the SAC / PPO updates.
//
// Replaces (reference): every nn.Linear + activation of sheeprl/models/models.py:16-119 (MLP) as used by
// SACActor / SACCritic (sheeprl/algos/sac/agent.py:19-108) and PPOAgent (sheeprl/algos/ppo/agent.py:84-177), forward
// and the three autograd products of each layer.  These updates are launch-latency bound (B = 64..256 rows, 64..256
// units: ~40 MFLOP per layer), so the design goal is FEW launches, not tensor-core tiles:
//   * `nets` independent networks (the twin critics and their targets) run in ONE launch (blockIdx.z) with a constant
//     stride between their parameter blocks in the flat group;
//   * bias, ReLU / Tanh, and the activation derivative of the backward-data product are epilogues;
//   * the bias gradient (column sum of dY) is produced by the weight-gradient launch itself (row sums of its A
//     operand), so a layer's backward is exactly two launches.
// All operands are addressed through (row stride, column stride) pairs, which covers NN / NT / TN without copies.
#include "common.cuh"

namespace {

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_TANH = 2, EPI_DRELU = 3, EPI_DTANH = 4 };

struct BG {
  const float* A; long long sam, sak, strideA;      // A(m,k) = A[m*sam + k*sak]
  const float* B; long long sbk, sbn, strideB;      // B(k,n) = B[k*sbk + n*sbn]
  float* C; long long ldc, strideC;                 // C(m,n) = C[m*ldc + n]
  const float* bias; long long strideBias;          // + bias[n]            (may be null)
  const float* aux; long long ldaux, strideAux;     // activation output of the layer, for EPI_D*
  float* rsum; long long strideRsum;                // rsum[m] = sum_k A(m,k) (may be null)
  int M, N, K, epi, accumulate;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
bgemm_kernel(const BG g) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int net = blockIdx.z;
  const float* __restrict__ A = g.A + net * g.strideA;
  const float* __restrict__ B = g.B + net * g.strideB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float rs = 0.f;                                    // row sum of A for row m0 + tid (tid < BM)
  const bool want_rsum = g.rsum != nullptr && blockIdx.x == 0;
  const bool a_kfast = g.sak == 1, b_nfast = g.sbn == 1;
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    // stage A tile (BM x BK): consecutive threads walk the unit-stride dimension
    for (int e = tid; e < BM * BK; e += NT) {
      const int m = a_kfast ? e / BK : e % BM;
      const int k = a_kfast ? e % BK : e / BM;
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < g.K) ? __ldg(A + gm * g.sam + gk * g.sak) : 0.f;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      const int n = b_nfast ? e % BN : e / BK;
      const int k = b_nfast ? e / BN : e % BK;
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < g.N && gk < g.K) ? __ldg(B + gk * g.sbk + gn * g.sbn) : 0.f;
    }
    __syncthreads();
    if (want_rsum && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) rs += As[k][tid];
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ C = g.C + net * g.strideC;
  const float* __restrict__ bias = g.bias ? g.bias + net * g.strideBias : nullptr;
  const float* __restrict__ aux = g.aux ? g.aux + net * g.strideAux : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (g.epi == EPI_RELU) v = fmaxf(v, 0.f);
      else if (g.epi == EPI_TANH) v = tanhf(v);
      else if (g.epi == EPI_DRELU) v = (aux[m * g.ldaux + n] > 0.f) ? v : 0.f;
      else if (g.epi == EPI_DTANH) { const float y = aux[m * g.ldaux + n]; v *= (1.f - y * y); }
      float* c = C + m * g.ldc + n;
      *c = g.accumulate ? *c + v : v;
    }
  }
  if (want_rsum && tid < BM && m0 + tid < g.M) {
    float* r = g.rsum + net * g.strideRsum + m0 + tid;
    *r = g.accumulate ? *r + rs : rs;
  }
}

}  // namespace

extern "C" int b200rl_bgemm(const float* A, long long sam, long long sak, long long strideA, const float* B,
                            long long sbk, long long sbn, long long strideB, float* C, long long ldc, long long strideC,
                            const float* bias, long long strideBias, const float* aux, long long ldaux,
                            long long strideAux, float* rsum, long long strideRsum, int M, int N, int K, int nets,
                            int epilogue, int accumulate, cudaStream_t st) {
  RL_CHECK_ARG(A && B && C, "null pointer");
  RL_CHECK_ARG(M > 0 && N > 0 && K > 0 && nets > 0, "bad dims");
  RL_CHECK_ARG(epilogue >= EPI_NONE && epilogue <= EPI_DTANH, "unknown epilogue");
  RL_CHECK_ARG(epilogue < EPI_DRELU || aux, "derivative epilogue needs the saved activation");
  BG g{A, sam, sak, strideA, B, sbk, sbn, strideB, C, ldc, strideC, bias, strideBias, aux, ldaux, strideAux,
       rsum, strideRsum, M, N, K, epilogue, accumulate};
  const long long tiles64 = (long long)((M + 63) / 64) * ((N + 63) / 64) * nets;
  if (tiles64 >= 2 * kNumSMs) {
    dim3 grid((N + 63) / 64, (M + 63) / 64, nets);
    bgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((N + 31) / 32, (M + 31) / 32, nets);
    bgemm_kernel<32, 32, 2, 2><<<grid, 256, 0, st>>>(g);
  }
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}
