// Persistent RSSM scan: all T steps of RSSM.dynamic (forward) and their BPTT (backward), each as ONE
// cooperative kernel.
//
// Replaces the Python loop `for i in range(sequence_length): rssm.dynamic(...)`
// (sheeprl/algos/dreamer_v3/dreamer_v3.py:131-145 -> agent.py:396-435: is_first masking, RecurrentModel +
// LayerNormGRUCell models.py:396-403, transition / representation MLPs, unimix, straight-through sampling)
// and the autograd replay of it inside `fabric.backward(rec_loss)` (dreamer_v3.py:191).
//
// Design (B200): the batch is tiny (B <= 16 rows) and the steps are strictly sequential, so the scan is
// latency-bound.  One CTA per SM (SCAN_G CTAs, cooperative launch) owns a fixed slice of OUTPUT COLUMNS of
// every weight matrix and keeps that slice resident in shared memory for the whole scan (S size: ~150 KB of
// weights per CTA; weights are read from HBM once per scan instead of once per step).  Per step the
// dependent skinny GEMM stages are separated by grid barriers; the state rows of all B sequences are staged in
// shared memory ([16][K] block) for each stage.  A warp takes a (4-column group) x (K-slice) work item:
// 16 rows x 4 cols accumulators per lane with 128-bit shared loads along K, then a 62-shuffle
// reduce-scatter.  z_{t-1} is one-hot per group, so z W_in^T is a gather of S weight columns.  LayerNorm
// statistics over rows that are spread across CTAs are merged from per-CTA partials (Chan / Welford in the
// forward, plain sums of the two backward reductions) exchanged through L2.  The 32 classes of a categorical
// sit on the 32 lanes of a warp (softmax / unimix / sampling / argmax by shuffles).
#include "b200rl.h"
#include "common.cuh"

namespace {

constexpr int SCAN_G = 128;    // CTAs (one per SM; 148 SMs available)
constexpr int SCAN_NT = 256;   // threads per CTA (8 warps: 255 registers per thread for the 64-accumulator tiles)
constexpr int SCAN_NW = SCAN_NT / 32;
constexpr int MAXB = 16;

__host__ __device__ inline int r4(int x) { return (x + 3) / 4 * 4; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

__host__ __device__ inline int owned_groups(int width, int cta) {
  const int groups = (width + 3) / 4;
  return (groups > cta) ? (groups - cta + SCAN_G - 1) / SCAN_G : 0;
}
// number of valid columns of `width` owned by `cta`
__host__ __device__ inline int owned_cols(int width, int cta) {
  int n = 0;
  for (int gi = cta; gi * 4 < width; gi += SCAN_G) n += (width - gi * 4 < 4) ? width - gi * 4 : 4;
  return n;
}

struct Workspace {
  unsigned* counter;   // grid barrier arrivals
  int* error;
  float* stats;        // [2][MAXB][SCAN_G][4]
  int* zidx;           // [T][B][S] sampled class per group
  float* ln_stats;     // [4][T*B][2] (mean, rstd) of the x / g / transition / representation LayerNorms
  float* dz_carry;     // [MAXB][Z]
};

__host__ __device__ inline size_t ws_bytes(int T, int B, int S, int D) {
  return 256 + sizeof(float) * 2 * MAXB * SCAN_G * 4 + sizeof(int) * (size_t)T * B * S +
         sizeof(float) * 4 * (size_t)T * B * 2 + sizeof(float) * MAXB * (size_t)S * D + 256;
}

__device__ inline Workspace carve(void* ws, int T, int B, int S, int D) {
  Workspace w;
  char* p = (char*)ws;
  w.counter = (unsigned*)p;
  w.error = (int*)(p + 64);
  p += 256;
  w.stats = (float*)p;      p += sizeof(float) * 2 * MAXB * SCAN_G * 4;
  w.zidx = (int*)p;         p += sizeof(int) * (size_t)T * B * S;
  w.ln_stats = (float*)p;   p += sizeof(float) * 4 * (size_t)T * B * 2;
  w.dz_carry = (float*)p;
  return w;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier (all SCAN_G CTAs are co-resident: cooperative launch). `target` advances by gridDim.x.
__device__ __forceinline__ void grid_barrier(const Workspace& w, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release: orders this CTA's prior global writes (made visible CTA-wide by the bar.sync above) before the arrival
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(w.counter) : "memory");
    unsigned polls = 0;
    long long t0 = 0;
    while (ld_acquire(w.counter) < target) {
      if ((++polls & 1023u) == 0u) {
        if (t0 == 0) t0 = clock64();
        if (clock64() - t0 > 4000000000LL) { atomicExch(w.error, 1); break; }   // ~2 s: never hang the device
        if (ld_acquire((const unsigned*)w.error) != 0u) break;                   // another CTA gave up
      }
    }
    // the acquire load orders the other CTAs' released writes before everything after the bar.sync below;
    // cross-CTA data is always read through L2 (__ldcg), so no L1 invalidation is needed
  }
  __syncthreads();
}

__device__ __forceinline__ bool scan_failed(const Workspace& w, int* flag_smem) {
  if (threadIdx.x == 0) *flag_smem = (int)ld_acquire((const unsigned*)w.error);
  __syncthreads();
  const bool f = *flag_smem != 0;
  __syncthreads();
  return f;
}

// fast transcendental form for the redundantly-evaluated (every CTA, full rows) SiLU; rel. error ~1e-6
__device__ __forceinline__ float fsilu(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// v[N] = acc[N/4 rows][4 cols] per lane; sum across the 32 lanes; lane ends up owning N/32 consecutive values
// starting at the returned base: N - N/32 shuffles instead of 5N.
template <int N>
__device__ __forceinline__ int reduce_scatter(float (&v)[N], int lane) {
#pragma unroll
  for (int off = 16, n = N; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float lo = v[i], hi = v[i + n / 2];
      const float send = upper ? lo : hi;
      const float keep = upper ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return ((lane >> 4) & 1) * (N / 2) + ((lane >> 3) & 1) * (N / 4) + ((lane >> 2) & 1) * (N / 8) +
         ((lane >> 1) & 1) * (N / 16) + (lane & 1) * (N / 32);
}

// out[b][c0 + j] += sum_{k in [k0,k1)} X[b][k] * W_j[k] for NR rows and the 4 columns (weight rows w + j*wst).
// X: smem [NR][KS] (KS % 4 == 0, zero padded); weight rows in smem (stride wst % 4 == 0, zero padded);
// k0, k1 multiples of 4.  One warp; a lane handles 4 consecutive k per 128-wide sweep (LDS.128).
template <int NR>
__device__ __forceinline__ void warp_item(const float* __restrict__ X, int KS, const float* __restrict__ w, int wst,
                                          int k0, int k1, float* out, int ldo, int c0, int lane) {
  float acc[NR * 4];
#pragma unroll
  for (int i = 0; i < NR * 4; ++i) acc[i] = 0.f;
  for (int k = k0 + 4 * lane; k < k1; k += 128) {
    const float4 a0 = *reinterpret_cast<const float4*>(w + k);
    const float4 a1 = *reinterpret_cast<const float4*>(w + wst + k);
    const float4 a2 = *reinterpret_cast<const float4*>(w + 2 * wst + k);
    const float4 a3 = *reinterpret_cast<const float4*>(w + 3 * wst + k);
#pragma unroll
    for (int b = 0; b < NR; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(X + b * KS + k);
      acc[b * 4 + 0] = fmaf(x.w, a0.w, fmaf(x.z, a0.z, fmaf(x.y, a0.y, fmaf(x.x, a0.x, acc[b * 4 + 0]))));
      acc[b * 4 + 1] = fmaf(x.w, a1.w, fmaf(x.z, a1.z, fmaf(x.y, a1.y, fmaf(x.x, a1.x, acc[b * 4 + 1]))));
      acc[b * 4 + 2] = fmaf(x.w, a2.w, fmaf(x.z, a2.z, fmaf(x.y, a2.y, fmaf(x.x, a2.x, acc[b * 4 + 2]))));
      acc[b * 4 + 3] = fmaf(x.w, a3.w, fmaf(x.z, a3.z, fmaf(x.y, a3.y, fmaf(x.x, a3.x, acc[b * 4 + 3]))));
    }
  }
  const int base = reduce_scatter<NR * 4>(acc, lane);
#pragma unroll
  for (int i = 0; i < NR / 8; ++i) {
    const int idx = base + i, b = idx >> 2, j = idx & 3;
    atomicAdd(&out[b * ldo + c0 + j], acc[i]);
  }
}

// Runs all (column group, K slice) items of one stage over the CTA's 16 warps; sums land in
// OUT[b][cbase + 4*group + j], b < NR.  `wbase`: first weight row of group 0; group g starts at wbase + g*4*wst.
template <int NR = MAXB>
__device__ __forceinline__ void run_stage(const float* X, int KS, const float* wbase, int wst, int ngroups, int K,
                                          float* OUT, int ldo, int cbase, bool zero, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  if (zero) {
    for (int e = tid; e < NR * ldo; e += SCAN_NT) OUT[e] = 0.f;
  }
  __syncthreads();
  if (ngroups > 0) {
    const int Kp = r4(K);
    int ks = SCAN_NW / ngroups;
    if (ks < 1) ks = 1;
    const int kchunk = ((Kp + ks - 1) / ks + 127) / 128 * 128;
    ks = (Kp + kchunk - 1) / kchunk;
    for (int item = wid; item < ngroups * ks; item += SCAN_NW) {
      const int cg = item % ngroups, sl = item / ngroups;
      const int k0 = sl * kchunk, k1 = min(Kp, k0 + kchunk);
      warp_item<NR>(X, KS, wbase + (size_t)cg * 4 * wst, wst, k0, k1, OUT, ldo, cbase + cg * 4, lane);
    }
  }
  __syncthreads();
}

// Loads `n` floats of a global row (L2 path: produced by other CTAs during the kernel) into smem, zero pads to np.
// All L2 requests of a lane are issued before the first use (one L2 round trip per row, not one per element).
__device__ __forceinline__ void warp_load_row(float* dst, const float* src, int n, int np, bool valid, int lane) {
  if (!valid) {
    for (int k = lane; k < np; k += 32) dst[k] = 0.f;
    return;
  }
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i0 = 0; i0 < n4; i0 += 128) {   // up to 4 x 128-bit loads in flight per lane
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i < n4) v[u] = __ldcg(s4 + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i < n4) d4[i] = v[u];
      }
    }
    done = n4 << 2;
  }
  for (int k0 = done; k0 < n; k0 += 128) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < n) v[u] = __ldcg(src + k);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < n) dst[k] = v[u];
    }
  }
  for (int k = n + lane; k < np; k += 32) dst[k] = 0.f;
}

// LayerNorm(+SiLU) of a row held in smem (length n), in place; one warp.  Returns (mean, rstd).
__device__ __forceinline__ float2 warp_ln_row(float* x, int n, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, bool silu, int lane) {
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s += x[k];
  const float mu = warp_sum(s) / (float)n;
  float v = 0.f;
  for (int k = lane; k < n; k += 32) { const float d = x[k] - mu; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / (float)n + eps);
  for (int k = lane; k < n; k += 32) {
    float o = (x[k] - mu) * rstd * gamma[k] + beta[k];
    if (silu) o = fsilu(o);
    x[k] = o;
  }
  return make_float2(mu, rstd);
}

// Copies 4 rows (row0..row0+3) of a row-major [rows][ld] weight into smem rows of stride `wst`, zero padded.
__device__ __forceinline__ void load_rows4(float* dst, int wst, const float* W, size_t ld, int row0, int nrows_valid,
                                           int K, int tid) {
  for (int j = 0; j < 4; ++j)
    for (int k = tid; k < wst; k += SCAN_NT)
      dst[j * wst + k] = (row0 + j < nrows_valid && k < K) ? W[(size_t)(row0 + j) * ld + k] : 0.f;
}
// Copies 4 COLUMNS (col0..col0+3) of a row-major [K][ld] weight into 4 smem rows (transposed slice) at offset koff.
__device__ __forceinline__ void load_cols4(float* dst, int wst, const float* W, size_t ld, int col0, int ncols_valid,
                                           int K, int koff, int tid) {
  for (int e = tid; e < K * 4; e += SCAN_NT) {
    const int k = e >> 2, j = e & 3;
    dst[j * wst + koff + k] = (col0 + j < ncols_valid) ? W[(size_t)k * ld + col0 + j] : 0.f;
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
struct GeoF {
  int ngx, ngh, ngt, ngr, nunit;
  int sKG, sR, sD2, KS, D4, KIN;
  int oWin, oWg, oW1, oW2, oX, oOut, oMisc, oZp, oNc, total;
};

__host__ __device__ inline GeoF make_geo_f(const b200rl_rssm_scan_args& a, int cta) {
  GeoF g;
  const int Z = a.S * a.D;
  g.ngx = owned_groups(a.Dx, cta);
  g.ngh = owned_groups(a.R, cta);
  g.ngt = owned_groups(a.Dt, cta);
  g.ngr = owned_groups(a.Dr, cta);
  g.nunit = (4 * a.S + SCAN_G - 1) / SCAN_G;   // (group, row-half) work units per CTA
  g.KIN = Z + a.A;
  g.sKG = r4(a.R + a.Dx);
  g.sR = r4(a.R);
  g.sD2 = r4(imax(a.Dt, a.Dr));
  g.D4 = r4(a.D);
  g.KS = imax(g.sKG, g.sD2);
  const int mx = owned_groups(a.Dx, 0), mh = owned_groups(a.R, 0), mt = owned_groups(a.Dt, 0), mr = owned_groups(a.Dr, 0);
  int o = 0;
  g.oWin = o;  o += r4(mx * 4 * g.KIN);
  g.oWg = o;   o += mh * 12 * g.sKG;
  g.oW1 = o;   o += (mt + mr) * 4 * g.sR;   // transition groups then representation groups (contiguous per CTA)
  g.oW2 = o;   o += g.nunit * g.D4 * g.sD2;
  g.oX = o;    o += MAXB * g.KS;
  const int outc = imax(imax(mh * 12, mx * 4), imax((mt + mr) * 4, 32));
  g.oOut = o;  o += MAXB * outc;
  g.oMisc = o; o += 4 * MAXB + 64;
  g.oZp = o;   o += r4(MAXB * a.S);
  g.oNc = o;   o += SCAN_G;
  g.total = o;
  return g;
}

__global__ void __launch_bounds__(SCAN_NT, 1) rssm_scan_fwd_kernel(const b200rl_rssm_scan_args a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const int T = a.T, B = a.B, S = a.S, D = a.D, Z = S * D, R = a.R, A = a.A, Dx = a.Dx, Dt = a.Dt, Dr = a.Dr;
  const int KG = R + Dx, NB = T * B;
  const GeoF g = make_geo_f(a, cta);
  const Workspace ws = carve(a.workspace, T, B, S, D);
  float* Win = sm + g.oWin;                 // [ngx*4][KIN]
  float* Wg = sm + g.oWg;                   // [ngh*12][sKG]  rows: (group, part r/c/u, col-in-group)
  float* Wt1 = sm + g.oW1;                  // [ngt*4][sR]
  float* Wr1 = Wt1 + (size_t)g.ngt * 4 * g.sR;   // [ngr*4][sR] directly after this CTA's transition rows
  float* W2s = sm + g.oW2;                  // [unit][D4][sD2]
  float* X = sm + g.oX;                     // [MAXB][KS]
  float* OUT = sm + g.oOut;                 // [MAXB][ldo]
  float* misc = sm + g.oMisc;   // [0,16): first flags; [16,32): mean; [32,48): rstd; 48: fail flag; [64,..): z0idx
  int* z0idx = (int*)(misc + 64);
  int* zprev = (int*)(sm + g.oZp);          // [B][S] class indices of z_{t-1}
  int* nctab = (int*)(sm + g.oNc);          // valid R-columns owned by every CTA (LayerNorm merge weights)
  const int KS = g.KS, KIN = g.KIN;
  unsigned bar_target = 0;

  // ---------------- prologue: weight slices -> shared memory (read from HBM once per scan)
  for (int gi = 0; gi < g.ngx; ++gi) {
    const int c0 = (cta + gi * SCAN_G) * 4;
    for (int j = 0; j < 4; ++j)
      for (int k = tid; k < KIN; k += SCAN_NT)
        Win[(gi * 4 + j) * KIN + k] = (c0 + j < Dx) ? a.W_in[(size_t)(c0 + j) * KIN + k] : 0.f;
  }
  for (int gi = 0; gi < g.ngh; ++gi)
    for (int part = 0; part < 3; ++part)
      load_rows4(Wg + (size_t)(gi * 3 + part) * 4 * g.sKG, g.sKG, a.W_g + (size_t)part * R * KG, KG,
                 (cta + gi * SCAN_G) * 4, R, KG, tid);
  for (int gi = 0; gi < g.ngt; ++gi)
    load_rows4(Wt1 + (size_t)gi * 4 * g.sR, g.sR, a.W_t1, R, (cta + gi * SCAN_G) * 4, Dt, R, tid);
  for (int gi = 0; gi < g.ngr; ++gi)
    load_rows4(Wr1 + (size_t)gi * 4 * g.sR, g.sR, a.W_r1, a.ld_wr1, (cta + gi * SCAN_G) * 4, Dr, R, tid);
  {
    int ui = 0;
    for (int hu = cta; hu < 4 * S; hu += SCAN_G, ++ui) {
      const int u = hu % (2 * S);
      const bool post = u < S;
      const int gq = post ? u : u - S, Dh = post ? Dr : Dt;
      const float* W2 = (post ? a.W_r2 : a.W_t2) + (size_t)gq * D * Dh;
      for (int jg = 0; jg < g.D4 / 4; ++jg)
        load_rows4(W2s + ((size_t)ui * g.D4 + jg * 4) * g.sD2, g.sD2, W2, Dh, jg * 4, D, Dh, tid);
    }
  }
  for (int c = tid; c < SCAN_G; c += SCAN_NT) nctab[c] = owned_cols(R, c);
  if (wid == 0) {  // class index of the learned initial posterior (one-hot `z0`)
    for (int gq = 0; gq < S; ++gq) {
      int best = 0;
      for (int d = lane; d < D; d += 32)
        if (a.z0[gq * D + d] > 0.5f) best = d;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0) z0idx[gq] = best;
    }
  }
  __syncthreads();

  const int n_units = 2 * S;  // stage-4 units: [0,S) posterior groups, [S,2S) prior groups

  for (int t = 0; t < T; ++t) {
    const size_t row0 = (size_t)t * B;
    if (tid < MAXB) misc[tid] = (tid < B) ? a.first[row0 + tid] : 0.f;
    if (t > 0)
      for (int e = tid; e < B * S; e += SCAN_NT) zprev[e] = __ldcg(&ws.zidx[(row0 - B) * S + e]);
    __syncthreads();
    const float* fl = misc;

    // ============ stage 1: x_pre = [z_in, a_in] W_in^T  (z_in one-hot -> column gather)
    {
      const int nout = B * g.ngx * 4;
      for (int o = tid; o < nout; o += SCAN_NT) {
        const int b = o / (g.ngx * 4), cj = o - b * (g.ngx * 4);
        const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
        if (col >= Dx) continue;
        const float* wrow = Win + cj * KIN;
        const float f = fl[b];
        float acc = 0.f;
        for (int gq = 0; gq < S; ++gq) {
          // z_in = (1-f) z_prev + f z0 (mask-multiply form of agent.py:430), both one-hot
          if (t > 0) acc = fmaf(1.f - f, wrow[gq * D + zprev[b * S + gq]], acc);
          acc = fmaf(f, wrow[gq * D + z0idx[gq]], acc);
        }
        for (int qq = 0; qq < A; ++qq) acc = fmaf((1.f - f) * a.actions[(row0 + b) * A + qq], wrow[Z + qq], acc);
        a.x_pre[(row0 + b) * Dx + col] = acc;
      }
      // dense saves for the deferred weight-gradient GEMMs: z_in / a_in rows, spread over the CTAs
      for (int e = cta * SCAN_NT + tid; e < B * Z; e += SCAN_G * SCAN_NT) {
        const int b = e / Z, k = e - b * Z;
        const int gq = k / D, d = k - gq * D;
        const float f = fl[b];
        float zp = 0.f;
        if (t > 0) zp = (zprev[b * S + gq] == d) ? 1.f : 0.f;
        a.z_in[(row0 + b) * Z + k] = (1.f - f) * zp + f * ((z0idx[gq] == d) ? 1.f : 0.f);
      }
      if (cta == (t % SCAN_G))
        for (int e = tid; e < B * A; e += SCAN_NT) {
          const int b = e / A;
          a.a_in[row0 * A + e] = (1.f - fl[b]) * a.actions[row0 * A + e];
        }
    }
    grid_barrier(ws, bar_target);  // B1: x_pre complete

    // ============ stage 2: g_pre = [h_in, SiLU(LN(x_pre))] W_g^T for the owned (r,c,u) column triples
    for (int b = wid; b < MAXB; b += SCAN_NW) {
      float* xr = X + b * KS;
      if (b < B) {
        const float f = fl[b];
        const size_t prow = (t > 0) ? row0 - B + b : 0;
        warp_load_row(xr, a.latent + prow * a.ld_lat + Z, R, R, t > 0, lane);
        warp_load_row(xr + R, a.x_pre + (row0 + b) * Dx, Dx, KS - R, true, lane);
        __syncwarp();
        for (int k = lane; k < R; k += 32) xr[k] = (1.f - f) * xr[k] + f * a.h0[k];   // agent.py:428 mask-mix
        __syncwarp();
        const float2 st = warp_ln_row(xr + R, Dx, a.lnx_g, a.lnx_b, a.eps, true, lane);
        if (cta == ((t + 1) % SCAN_G)) {
          for (int k = lane; k < R; k += 32) a.h_in[(row0 + b) * R + k] = xr[k];
          for (int k = lane; k < Dx; k += 32) a.x_act[(row0 + b) * Dx + k] = xr[R + k];
          if (lane == 0) { ws.ln_stats[((size_t)0 * NB + row0 + b) * 2] = st.x; ws.ln_stats[((size_t)0 * NB + row0 + b) * 2 + 1] = st.y; }
        }
      } else {
        for (int k = lane; k < KS; k += 32) xr[k] = 0.f;
      }
    }
    const int ldo2 = imax(g.ngh * 12, 4);
    run_stage(X, KS, Wg, g.sKG, g.ngh * 3, KG, OUT, ldo2, 0, true, tid);
    // save g_pre columns; per-row partial statistics (mean, M2) over the owned valid columns
    {
      const int par = t & 1;
      for (int b = wid; b < B; b += SCAN_NW) {
        float s = 0.f;
        int cnt = 0;
        for (int c = lane; c < g.ngh * 12; c += 32) {
          const int gi = c / 12, part = (c % 12) / 4, j = c & 3;
          const int col = (cta + gi * SCAN_G) * 4 + j;
          if (col < R) {
            const float v = OUT[b * ldo2 + c];
            a.g_pre[(row0 + b) * 3 * R + part * R + col] = v;
            s += v;
            ++cnt;
          }
        }
        s = warp_sum(s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const float mean = cnt > 0 ? s / (float)cnt : 0.f;
        float m2 = 0.f;
        for (int c = lane; c < g.ngh * 12; c += 32) {
          const int gi = c / 12, j = c & 3;
          const int col = (cta + gi * SCAN_G) * 4 + j;
          if (col < R) { const float d = OUT[b * ldo2 + c] - mean; m2 = fmaf(d, d, m2); }
        }
        m2 = warp_sum(m2);
        if (lane == 0) {
          float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + cta) * 4;
          st[0] = mean;
          st[1] = m2;
        }
      }
    }
    grid_barrier(ws, bar_target);  // B2: partial LN statistics complete

    // ============ stage 2b: merge statistics, LayerNorm, GRU gate -> h_t for the owned columns
    {
      const int par = t & 1;
      for (int b = wid; b < B; b += SCAN_NW) {
        float2 pv[SCAN_G / 32];   // all partials of this row in flight at once (one L2 round trip)
#pragma unroll
        for (int i = 0; i < SCAN_G / 32; ++i)
          pv[i] = __ldcg(reinterpret_cast<const float2*>(ws.stats + (((size_t)par * MAXB + b) * SCAN_G + lane + 32 * i) * 4));
        float sm_ = 0.f;
#pragma unroll
        for (int i = 0; i < SCAN_G / 32; ++i) sm_ += (float)(3 * nctab[lane + 32 * i]) * pv[i].x;
        const float mean = warp_sum(sm_) / (float)(3 * R);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < SCAN_G / 32; ++i) {
          const float d = pv[i].x - mean;
          m2 += pv[i].y + (float)(3 * nctab[lane + 32 * i]) * d * d;
        }
        m2 = warp_sum(m2);
        if (lane == 0) {
          const float rstd = rsqrtf(m2 / (float)(3 * R) + a.eps);
          misc[16 + b] = mean;
          misc[32 + b] = rstd;
          if (cta == ((t + 2) % SCAN_G)) {
            ws.ln_stats[((size_t)1 * NB + row0 + b) * 2] = mean;
            ws.ln_stats[((size_t)1 * NB + row0 + b) * 2 + 1] = rstd;
          }
        }
      }
      __syncthreads();
      for (int e = tid; e < B * g.ngh * 4; e += SCAN_NT) {
        const int b = e / (g.ngh * 4), cj = e - b * (g.ngh * 4);
        const int gi = cj >> 2, j = cj & 3;
        const int col = (cta + gi * SCAN_G) * 4 + j;
        if (col >= R) continue;
        const float mu = misc[16 + b], rstd = misc[32 + b];
        float gl[3];
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const float v = OUT[b * ldo2 + gi * 12 + part * 4 + j];
          gl[part] = (v - mu) * rstd * a.lng_g[part * R + col] + a.lng_b[part * R + col];
          a.g_ln[(row0 + b) * 3 * R + part * R + col] = gl[part];
        }
        const float r = sigmoidf_(gl[0]);
        const float c = tanhf(r * gl[1]);
        const float u = sigmoidf_(gl[2] - 1.f);
        const float hin = X[b * KS + col];
        a.latent[(row0 + b) * a.ld_lat + Z + col] = u * c + (1.f - u) * hin;
      }
    }
    grid_barrier(ws, bar_target);  // B3: h_t complete

    // ============ stage 3: tr_pre = h W_t1^T ; rp_pre = h W_r1[:, :R]^T + pe
    for (int b = wid; b < MAXB; b += SCAN_NW)
      warp_load_row(X + b * KS, a.latent + (row0 + b) * a.ld_lat + Z, R, g.sR, b < B, lane);
    const int ldo3 = imax((g.ngt + g.ngr) * 4, 4);
    run_stage(X, KS, Wt1, g.sR, g.ngt + g.ngr, R, OUT, ldo3, 0, true, tid);
    for (int e = tid; e < B * (g.ngt + g.ngr) * 4; e += SCAN_NT) {
      const int nc = (g.ngt + g.ngr) * 4;
      const int b = e / nc, c = e - b * nc;
      const int cg = c >> 2, j = c & 3;
      if (cg < g.ngt) {
        const int col = (cta + cg * SCAN_G) * 4 + j;
        if (col < Dt) a.tr_pre[(row0 + b) * Dt + col] = OUT[b * ldo3 + c];
      } else {
        const int col = (cta + (cg - g.ngt) * SCAN_G) * 4 + j;
        if (col < Dr) a.rp_pre[(row0 + b) * Dr + col] = OUT[b * ldo3 + c] + a.pe[(row0 + b) * Dr + col];
      }
    }
    grid_barrier(ws, bar_target);  // B4: tr_pre / rp_pre complete

    // ============ stage 4: logits of one categorical group per (unit, row-half), unimix, sample (posterior only).
    // Every CTA takes one half-unit: 8 of the 16 rows of one group, so all 128 SMs work here.
    int unit_i = 0;
    for (int hu = cta; hu < 2 * n_units; hu += SCAN_G, ++unit_i) {
      const int u = hu % n_units, rb = (hu / n_units) * 8;     // unit and first row of this half
      const bool post = u < S;
      const int gq = post ? u : u - S;
      const int Dh = post ? Dr : Dt;
      const float* pre = post ? a.rp_pre : a.tr_pre;
      float* act_save = post ? a.rp_act : a.tr_act;
      const float* lg_ = post ? a.lnr_g : a.lnt_g;
      const float* lb_ = post ? a.lnr_b : a.lnt_b;
      const float* b2 = post ? a.b_r2 : a.b_t2;
      __syncthreads();
      for (int bb = wid; bb < 8; bb += SCAN_NW) {
        const int b = rb + bb;
        float* xr = X + bb * KS;
        warp_load_row(xr, pre + (row0 + b) * Dh, Dh, g.sD2, b < B, lane);
        if (b < B) {
          __syncwarp();
          const float2 st = warp_ln_row(xr, Dh, lg_, lb_, a.eps, true, lane);
          if (gq == 0) {
            for (int k = lane; k < Dh; k += 32) act_save[(row0 + b) * Dh + k] = xr[k];
            const size_t which = post ? 3 : 2;
            if (lane == 0) { ws.ln_stats[(which * NB + row0 + b) * 2] = st.x; ws.ln_stats[(which * NB + row0 + b) * 2 + 1] = st.y; }
          }
        }
      }
      const int ldo4 = g.D4;
      run_stage<8>(X, KS, W2s + (size_t)unit_i * g.D4 * g.sD2, g.sD2, g.D4 / 4, Dh, OUT, ldo4, 0, true, tid);
      // one warp per row: the D classes of the group live on the lanes (D <= 32)
      for (int bb = wid; bb < 8; bb += SCAN_NW) {
        const int b = rb + bb;
        if (b >= B) continue;
        const bool on = lane < D;
        const float raw = on ? OUT[bb * ldo4 + lane] + b2[gq * D + lane] : -INFINITY;
        const size_t o = (row0 + b) * Z + (size_t)gq * D + lane;
        const float mx = warp_max(raw);
        const float ex = on ? expf(raw - mx) : 0.f;
        const float se = warp_sum(ex);
        float l = raw;
        if (a.unimix > 0.f && on) {
          const float pm = (1.f - a.unimix) * (ex / se) + a.unimix / (float)D;
          l = logf(fminf(fmaxf(pm, kFp32Eps), 1.f - kFp32Eps));
        }
        if (on) {
          (post ? a.post_raw : a.prior_raw)[o] = raw;
          (post ? a.post_mix : a.prior_mix)[o] = l;
        }
        if (!post) continue;
        // torch Categorical: lg = l - logsumexp(l); probs = softmax(lg); sample = argmax(probs / q)
        const float lmx = warp_max(on ? l : -INFINITY);
        const float lse = lmx + logf(warp_sum(on ? expf(l - lmx) : 0.f));
        const float lgmax = warp_max(on ? l - lse : -INFINITY);
        const float pe_ = on ? expf(l - lse - lgmax) : 0.f;
        const float psum = warp_sum(pe_);
        float best = on ? (pe_ / psum) / a.noise[o] : -INFINITY;
        int besti = on ? lane : 0x7fffffff;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, s);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, s);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (on) a.latent[(row0 + b) * a.ld_lat + (size_t)gq * D + lane] = (lane == besti) ? 1.f : 0.f;
        if (lane == 0) ws.zidx[(row0 + b) * S + gq] = besti;
      }
    }
    grid_barrier(ws, bar_target);  // B5: z_t complete
    if ((t & 15) == 15 && scan_failed(ws, (int*)(misc + 48))) return;   // barrier time-out: bail out, never hang
  }
}

// =====================================================================================================
// backward (BPTT); consumes the activations + LayerNorm statistics saved by the forward kernel above
// =====================================================================================================
struct GeoB {
  int ngx, ngh, ngt, ngr, ngz;
  int sZ, sH1, s3R, sDx, KS;
  int oW2, oW1, oWg, oWin, oX, oOut, oDxh, oXh, oDhin, oDhc, oDh0, oMisc, total;
};

__host__ __device__ inline GeoB make_geo_b(const b200rl_rssm_scan_args& a, int cta) {
  GeoB g;
  const int Z = a.S * a.D;
  g.ngx = owned_groups(a.Dx, cta);
  g.ngh = owned_groups(a.R, cta);
  g.ngt = owned_groups(a.Dt, cta);
  g.ngr = owned_groups(a.Dr, cta);
  g.ngz = owned_groups(Z, cta);
  g.sZ = r4(Z);
  g.sH1 = r4(a.Dr) + r4(a.Dt);
  g.s3R = r4(3 * a.R);
  g.sDx = r4(a.Dx);
  g.KS = imax(imax(g.sZ, g.sH1), imax(g.s3R, g.sDx));
  const int mx = owned_groups(a.Dx, 0), mh = owned_groups(a.R, 0), mt = owned_groups(a.Dt, 0), mr = owned_groups(a.Dr, 0),
            mz = owned_groups(Z, 0);
  int o = 0;
  g.oW2 = o;   o += (mr + mt) * 4 * g.sZ;    // W_r2[:, c] slices then W_t2[:, c] slices (transposed)
  g.oW1 = o;   o += mh * 4 * g.sH1;          // [W_r1[:, j] | W_t1[:, j]] for dh columns
  g.oWg = o;   o += (mh + mx) * 4 * g.s3R;   // W_g[:, j] (h part) then W_g[:, R + c] (x part)
  g.oWin = o;  o += mz * 4 * g.sDx;          // W_in[:, z]
  g.oX = o;    o += MAXB * g.KS;
  const int outc = imax(imax((mr + mt) * 4, (mh + mx) * 4), imax(mz * 4, 32));
  g.oOut = o;  o += MAXB * outc;
  const int stc = imax(imax((mr + mt) * 4, mh * 12), imax(mx * 4, 4));
  g.oDxh = o;  o += MAXB * stc;
  g.oXh = o;   o += MAXB * stc;
  g.oDhin = o; o += MAXB * imax(mh * 4, 4);
  g.oDhc = o;  o += MAXB * imax(mh * 4, 4);
  g.oDh0 = o;  o += imax(mh * 4, 4);
  g.oMisc = o; o += 8 * MAXB + 64;
  g.total = o;
  return g;
}

__global__ void __launch_bounds__(SCAN_NT, 1)
rssm_scan_bwd_kernel(const b200rl_rssm_scan_args a, const b200rl_rssm_scan_grads q) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const int T = a.T, B = a.B, S = a.S, D = a.D, Z = S * D, R = a.R, Dx = a.Dx, Dt = a.Dt, Dr = a.Dr;
  const int KG = R + Dx, KIN = Z + a.A, NB = T * B;
  const GeoB g = make_geo_b(a, cta);
  const Workspace ws = carve(a.workspace, T, B, S, D);
  float* W2r = sm + g.oW2;                        // [ngr*4][sZ]
  float* W2t = W2r + (size_t)g.ngr * 4 * g.sZ;    // [ngt*4][sZ]
  float* W1 = sm + g.oW1;                         // [ngh*4][sH1]
  float* WgT = sm + g.oWg;                        // [(ngh+ngx)*4][s3R]
  float* WinT = sm + g.oWin;                      // [ngz*4][sDx]
  float* X = sm + g.oX;
  float* OUT = sm + g.oOut;
  float* DXH = sm + g.oDxh;     // dxh = dLN * gamma for the owned columns of the current LayerNorm
  float* XH = sm + g.oXh;       // normalised pre-activation for the same columns
  float* DHIN = sm + g.oDhin;   // [MAXB][ngh*4] dh wrt h_in (gate part)
  float* DHC = sm + g.oDhc;     // [MAXB][ngh*4] dh carried to step t-1
  float* DH0 = sm + g.oDh0;     // [ngh*4] accumulated grad of tanh(initial_recurrent_state)
  float* misc = sm + g.oMisc;   // [0,16) first; [16,32) S1; [32,48) S2; [48,64) S1b; [64,80) S2b; 80: fail flag
  const int KS = g.KS;
  const int nh4 = g.ngh * 4;
  unsigned bar_target = 0;

  // ---------------- prologue: zero everything, then transposed weight slices -> shared memory
  for (int e = tid; e < g.total; e += SCAN_NT) sm[e] = 0.f;
  __syncthreads();
  for (int gi = 0; gi < g.ngr; ++gi)
    load_cols4(W2r + (size_t)gi * 4 * g.sZ, g.sZ, a.W_r2, Dr, (cta + gi * SCAN_G) * 4, Dr, Z, 0, tid);
  for (int gi = 0; gi < g.ngt; ++gi)
    load_cols4(W2t + (size_t)gi * 4 * g.sZ, g.sZ, a.W_t2, Dt, (cta + gi * SCAN_G) * 4, Dt, Z, 0, tid);
  for (int gi = 0; gi < g.ngh; ++gi) {
    const int c0 = (cta + gi * SCAN_G) * 4;
    load_cols4(W1 + (size_t)gi * 4 * g.sH1, g.sH1, a.W_r1, a.ld_wr1, c0, R, Dr, 0, tid);
    load_cols4(W1 + (size_t)gi * 4 * g.sH1, g.sH1, a.W_t1, R, c0, R, Dt, r4(Dr), tid);
    load_cols4(WgT + (size_t)gi * 4 * g.s3R, g.s3R, a.W_g, KG, c0, R, 3 * R, 0, tid);
  }
  for (int gi = 0; gi < g.ngx; ++gi)
    load_cols4(WgT + (size_t)(g.ngh + gi) * 4 * g.s3R, g.s3R, a.W_g + R, KG, (cta + gi * SCAN_G) * 4, Dx, 3 * R, 0, tid);
  for (int gi = 0; gi < g.ngz; ++gi)
    load_cols4(WinT + (size_t)gi * 4 * g.sDx, g.sDx, a.W_in, KIN, (cta + gi * SCAN_G) * 4, Z, Dx, 0, tid);
  __syncthreads();

  const int n_units = 2 * S;
  const float invDr = 1.f / (float)Dr, invDt = 1.f / (float)Dt, inv3R = 1.f / (float)(3 * R), invDx = 1.f / (float)Dx;
  const int ld2 = imax((g.ngr + g.ngt) * 4, 4);
  const int ld3 = imax(nh4, 4);
  const int ld12 = imax(g.ngh * 12, 4);
  const int ld4 = imax((g.ngh + g.ngx) * 4, 4);
  const int nx4 = g.ngx * 4, ldx = imax(nx4, 4);
  const int ld5 = imax(g.ngz * 4, 4);

  for (int t = T - 1; t >= 0; --t) {
    const size_t row0 = (size_t)t * B;
    const bool last = (t == T - 1);
    const int par = t & 1;
    if (tid < MAXB) misc[tid] = (tid < B) ? a.first[row0 + tid] : 0.f;
    __syncthreads();
    const float* fl = misc;

    // ============ BS1: raw-logit gradients of every categorical group (straight-through + KL), group-local
    for (int u = cta; u < n_units; u += SCAN_G) {
      const bool post = u < S;
      const int gq = post ? u : u - S;
      for (int b = wid; b < B; b += SCAN_NW) {
        const bool on = lane < D;
        const size_t o = (row0 + b) * Z + (size_t)gq * D + lane;
        const float raw = on ? (post ? a.post_raw : a.prior_raw)[o] : -INFINITY;
        const float mx = warp_max(raw);
        const float ex = on ? expf(raw - mx) : 0.f;
        const float se = warp_sum(ex);
        const float s = ex / se;
        float pm = 0.f, l = raw;
        if (a.unimix > 0.f && on) {
          pm = (1.f - a.unimix) * s + a.unimix / (float)D;
          l = logf(fminf(fmaxf(pm, kFp32Eps), 1.f - kFp32Eps));
        }
        float gg = on ? (post ? q.d_post_mix : q.d_prior_mix)[o] : 0.f;
        if (post) {
          const float lmx = warp_max(on ? l : -INFINITY);
          const float lse = lmx + logf(warp_sum(on ? expf(l - lmx) : 0.f));
          const float p = on ? expf(l - lse) : 0.f;
          float dz = 0.f;
          if (on) {
            dz = q.d_latent[(row0 + b) * a.ld_lat + (size_t)gq * D + lane];
            if (!last) dz += __ldcg(&ws.dz_carry[(size_t)b * Z + (size_t)gq * D + lane]);
          }
          const float pdz = warp_sum(p * dz);
          gg += p * (dz - pdz);
        }
        if (a.unimix > 0.f) {
          const bool inside = on && pm >= kFp32Eps && pm <= 1.f - kFp32Eps;
          const float ds = inside ? gg * (1.f - a.unimix) / pm : 0.f;
          const float sds = warp_sum(s * ds);
          gg = s * (ds - sds);
        }
        if (on) (post ? q.d_post_raw : q.d_prior_raw)[o] = gg;
      }
    }
    grid_barrier(ws, bar_target);  // b1

    // ============ BS2: d_rp_act = d_post_raw W_r2 ; d_tr_act = d_prior_raw W_t2 (owned columns) + LN/SiLU backward
    for (int b = wid; b < MAXB; b += SCAN_NW) warp_load_row(X + b * KS, q.d_post_raw + (row0 + b) * Z, Z, g.sZ, b < B, lane);
    run_stage(X, KS, W2r, g.sZ, g.ngr, Z, OUT, ld2, 0, true, tid);
    for (int b = wid; b < MAXB; b += SCAN_NW) warp_load_row(X + b * KS, q.d_prior_raw + (row0 + b) * Z, Z, g.sZ, b < B, lane);
    run_stage(X, KS, W2t, g.sZ, g.ngt, Z, OUT, ld2, g.ngr * 4, false, tid);
    for (int b = wid; b < B; b += SCAN_NW) {
      float s1r = 0.f, s2r = 0.f, s1t = 0.f, s2t = 0.f;
      for (int c = lane; c < (g.ngr + g.ngt) * 4; c += 32) {
        const bool isr = c < g.ngr * 4;
        const int cc = isr ? c : c - g.ngr * 4;
        const int col = (cta + (cc >> 2) * SCAN_G) * 4 + (cc & 3);
        const int Dh = isr ? Dr : Dt;
        float dxh = 0.f, xh = 0.f;
        if (col < Dh) {
          const float dact = OUT[b * ld2 + c];
          (isr ? q.d_rp_act : q.d_tr_act)[(row0 + b) * Dh + col] = dact;
          const float* stp = ws.ln_stats + ((size_t)(isr ? 3 : 2) * NB + row0 + b) * 2;
          const float pre = (isr ? a.rp_pre : a.tr_pre)[(row0 + b) * Dh + col];
          xh = (pre - stp[0]) * stp[1];
          const float gam = (isr ? a.lnr_g : a.lnt_g)[col], bet = (isr ? a.lnr_b : a.lnt_b)[col];
          const float ln = xh * gam + bet;
          const float sg = sigmoidf_(ln);
          dxh = dact * sg * (1.f + ln * (1.f - sg)) * gam;
          if (isr) { s1r += dxh; s2r = fmaf(dxh, xh, s2r); } else { s1t += dxh; s2t = fmaf(dxh, xh, s2t); }
        }
        DXH[b * ld2 + c] = dxh;
        XH[b * ld2 + c] = xh;
      }
      s1r = warp_sum(s1r); s2r = warp_sum(s2r); s1t = warp_sum(s1t); s2t = warp_sum(s2t);
      if (lane == 0) {
        float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + cta) * 4;
        st[0] = s1r; st[1] = s2r; st[2] = s1t; st[3] = s2t;
      }
    }
    grid_barrier(ws, bar_target);  // b2
    for (int b = wid; b < B; b += SCAN_NW) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      for (int c = lane; c < SCAN_G; c += 32) {
        const float4 s = __ldcg(reinterpret_cast<const float4*>(ws.stats + (((size_t)par * MAXB + b) * SCAN_G + c) * 4));
        v0 += s.x; v1 += s.y; v2 += s.z; v3 += s.w;
      }
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
      if (lane == 0) { misc[16 + b] = v0 * invDr; misc[32 + b] = v1 * invDr; misc[48 + b] = v2 * invDt; misc[64 + b] = v3 * invDt; }
    }
    __syncthreads();
    for (int e = tid; e < B * (g.ngr + g.ngt) * 4; e += SCAN_NT) {
      const int ncol = (g.ngr + g.ngt) * 4;
      const int b = e / ncol, c = e - b * ncol;
      const bool isr = c < g.ngr * 4;
      const int cc = isr ? c : c - g.ngr * 4;
      const int col = (cta + (cc >> 2) * SCAN_G) * 4 + (cc & 3);
      const int Dh = isr ? Dr : Dt;
      if (col >= Dh) continue;
      const float rstd = ws.ln_stats[((size_t)(isr ? 3 : 2) * NB + row0 + b) * 2 + 1];
      const float S1 = isr ? misc[16 + b] : misc[48 + b], S2 = isr ? misc[32 + b] : misc[64 + b];
      (isr ? q.d_rp_pre : q.d_tr_pre)[(row0 + b) * Dh + col] = rstd * (DXH[b * ld2 + c] - S1 - XH[b * ld2 + c] * S2);
    }
    grid_barrier(ws, bar_target);  // b3

    // ============ BS3: dh = d_latent_h + carry + d_rp_pre W_r1h + d_tr_pre W_t1 ; GRU gate backward ; LN(3R) partials
    for (int b = wid; b < MAXB; b += SCAN_NW) {
      warp_load_row(X + b * KS, q.d_rp_pre + (row0 + b) * Dr, Dr, r4(Dr), b < B, lane);
      warp_load_row(X + b * KS + r4(Dr), q.d_tr_pre + (row0 + b) * Dt, Dt, r4(Dt), b < B, lane);
    }
    run_stage(X, KS, W1, g.sH1, g.ngh, g.sH1, OUT, ld3, 0, true, tid);
    for (int b = wid; b < B; b += SCAN_NW) {
      float s1 = 0.f, s2 = 0.f;
      const float mu = ws.ln_stats[((size_t)1 * NB + row0 + b) * 2], rstd = ws.ln_stats[((size_t)1 * NB + row0 + b) * 2 + 1];
      for (int c = lane; c < nh4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        float dgl[3] = {0.f, 0.f, 0.f};
        float dhin = 0.f;
        if (col < R) {
          float dh = q.d_latent[(row0 + b) * a.ld_lat + Z + col] + OUT[b * ld3 + c];
          if (!last) dh += DHC[b * ld3 + c];
          const float* gl = a.g_ln + (row0 + b) * 3 * R;
          const float gr = gl[col], gc = gl[R + col], gu = gl[2 * R + col];
          const float r = sigmoidf_(gr), cnd = tanhf(r * gc), u = sigmoidf_(gu - 1.f);
          const float hin = a.h_in[(row0 + b) * R + col];
          const float du = dh * (cnd - hin);
          const float drc = dh * u * (1.f - cnd * cnd);
          dgl[0] = drc * gc * r * (1.f - r);
          dgl[1] = drc * r;
          dgl[2] = du * u * (1.f - u);
          dhin = dh * (1.f - u);
        }
        DHIN[b * ld3 + c] = dhin;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          float dxh = 0.f, xh = 0.f;
          if (col < R) {
            q.d_g_ln[(row0 + b) * 3 * R + part * R + col] = dgl[part];
            xh = (a.g_pre[(row0 + b) * 3 * R + part * R + col] - mu) * rstd;
            dxh = dgl[part] * a.lng_g[part * R + col];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
          }
          DXH[b * ld12 + part * nh4 + c] = dxh;
          XH[b * ld12 + part * nh4 + c] = xh;
        }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) {
        float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + cta) * 4;
        st[0] = s1; st[1] = s2;
      }
    }
    grid_barrier(ws, bar_target);  // b4
    for (int b = wid; b < B; b += SCAN_NW) {
      float v0 = 0.f, v1 = 0.f;
      for (int c = lane; c < SCAN_G; c += 32) {
        const float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + c) * 4;
        v0 += __ldcg(st); v1 += __ldcg(st + 1);
      }
      v0 = warp_sum(v0); v1 = warp_sum(v1);
      if (lane == 0) { misc[16 + b] = v0 * inv3R; misc[32 + b] = v1 * inv3R; }
    }
    __syncthreads();
    for (int e = tid; e < B * 3 * nh4; e += SCAN_NT) {
      const int b = e / (3 * nh4), r_ = e - b * 3 * nh4;
      const int part = r_ / nh4, c = r_ - part * nh4;
      const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
      if (col >= R) continue;
      const float rstd = ws.ln_stats[((size_t)1 * NB + row0 + b) * 2 + 1];
      q.d_g_pre[(row0 + b) * 3 * R + part * R + col] =
          rstd * (DXH[b * ld12 + part * nh4 + c] - misc[16 + b] - XH[b * ld12 + part * nh4 + c] * misc[32 + b]);
    }
    grid_barrier(ws, bar_target);  // b5

    // ============ BS4: [dh_in, d_x_act] = d_g_pre W_g (owned columns) ; x-LN/SiLU backward partials
    for (int b = wid; b < MAXB; b += SCAN_NW)
      warp_load_row(X + b * KS, q.d_g_pre + (row0 + b) * 3 * R, 3 * R, g.s3R, b < B, lane);
    run_stage(X, KS, WgT, g.s3R, g.ngh + g.ngx, 3 * R, OUT, ld4, 0, true, tid);
    for (int b = wid; b < B; b += SCAN_NW) {
      const float f = fl[b];
      for (int c = lane; c < nh4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        if (col < R) {
          const float dhin = DHIN[b * ld3 + c] + OUT[b * ld4 + c];
          DHC[b * ld3 + c] = (1.f - f) * dhin;          // carried to step t-1 (agent.py:428 mask)
          atomicAdd(&DH0[c], f * dhin);                 // grad of tanh(initial_recurrent_state)
        }
      }
      float s1 = 0.f, s2 = 0.f;
      const float mu = ws.ln_stats[((size_t)0 * NB + row0 + b) * 2], rstd = ws.ln_stats[((size_t)0 * NB + row0 + b) * 2 + 1];
      for (int c = lane; c < nx4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        float dxh = 0.f, xh = 0.f;
        if (col < Dx) {
          const float dact = OUT[b * ld4 + nh4 + c];
          q.d_x_act[(row0 + b) * Dx + col] = dact;
          xh = (a.x_pre[(row0 + b) * Dx + col] - mu) * rstd;
          const float gam = a.lnx_g[col];
          const float ln = xh * gam + a.lnx_b[col];
          const float sg = sigmoidf_(ln);
          dxh = dact * sg * (1.f + ln * (1.f - sg)) * gam;
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
        DXH[b * ldx + c] = dxh;
        XH[b * ldx + c] = xh;
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) {
        float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + cta) * 4;
        st[0] = s1; st[1] = s2;
      }
    }
    grid_barrier(ws, bar_target);  // b6
    for (int b = wid; b < B; b += SCAN_NW) {
      float v0 = 0.f, v1 = 0.f;
      for (int c = lane; c < SCAN_G; c += 32) {
        const float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + c) * 4;
        v0 += __ldcg(st); v1 += __ldcg(st + 1);
      }
      v0 = warp_sum(v0); v1 = warp_sum(v1);
      if (lane == 0) { misc[16 + b] = v0 * invDx; misc[32 + b] = v1 * invDx; }
    }
    __syncthreads();
    for (int e = tid; e < B * nx4; e += SCAN_NT) {
      const int b = e / nx4, c = e - b * nx4;
      const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
      if (col >= Dx) continue;
      const float rstd = ws.ln_stats[((size_t)0 * NB + row0 + b) * 2 + 1];
      q.d_x_pre[(row0 + b) * Dx + col] = rstd * (DXH[b * ldx + c] - misc[16 + b] - XH[b * ldx + c] * misc[32 + b]);
    }
    grid_barrier(ws, bar_target);  // b7

    // ============ BS5: dz_in = d_x_pre W_in[:, :Z] (owned columns) -> carried to step t-1
    for (int b = wid; b < MAXB; b += SCAN_NW)
      warp_load_row(X + b * KS, q.d_x_pre + (row0 + b) * Dx, Dx, g.sDx, b < B, lane);
    run_stage(X, KS, WinT, g.sDx, g.ngz, Dx, OUT, ld5, 0, true, tid);
    for (int e = tid; e < B * g.ngz * 4; e += SCAN_NT) {
      const int b = e / (g.ngz * 4), c = e - b * (g.ngz * 4);
      const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
      if (col < Z) ws.dz_carry[(size_t)b * Z + col] = (1.f - fl[b]) * OUT[b * ld5 + c];
    }
    grid_barrier(ws, bar_target);  // b8
    if ((t & 15) == 0 && scan_failed(ws, (int*)(misc + 80))) break;   // barrier time-out: bail out, never hang
  }
  for (int c = tid; c < nh4; c += SCAN_NT) {
    const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
    if (col < R) q.d_h0[col] = DH0[c];
  }
}

int scan_check(const b200rl_rssm_scan_args& a) {
  RL_CHECK_ARG(a.B >= 1 && a.B <= MAXB, "persistent scan supports batch <= 16 rows per rank");
  RL_CHECK_ARG(a.D >= 1 && a.D <= 32, "persistent scan supports <= 32 classes per categorical");
  RL_CHECK_ARG(a.T >= 1 && a.S >= 1 && a.S <= 64, "bad T / S (S <= 64)");
  RL_CHECK_ARG(a.workspace && a.workspace_bytes >= (long long)ws_bytes(a.T, a.B, a.S, a.D), "workspace too small");
  return B200RL_OK;
}

}  // namespace

extern "C" long long b200rl_rssm_scan_workspace_bytes(int T, int B, int S, int D) {
  return (long long)ws_bytes(T, B, S, D);
}

extern "C" int b200rl_rssm_scan_fwd(const b200rl_rssm_scan_args* args, cudaStream_t st) {
  RL_CHECK_ARG(args, "null args");
  const b200rl_rssm_scan_args& a = *args;
  if (int rc = scan_check(a)) return rc;
  const GeoF g = make_geo_f(a, 0);
  const size_t smem = sizeof(float) * (size_t)g.total;
  RL_CHECK_ARG(smem <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  RL_CUDA(cudaFuncSetAttribute(rssm_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RL_CUDA(cudaMemsetAsync(a.workspace, 0, 256, st));
  void* kargs[] = {(void*)args};
  RL_CUDA(cudaLaunchCooperativeKernel((void*)rssm_scan_fwd_kernel, dim3(SCAN_G), dim3(SCAN_NT), kargs, smem, st));
  return B200RL_OK;
}

extern "C" int b200rl_rssm_scan_bwd(const b200rl_rssm_scan_args* args, const b200rl_rssm_scan_grads* grads,
                                    cudaStream_t st) {
  RL_CHECK_ARG(args && grads, "null args");
  const b200rl_rssm_scan_args& a = *args;
  if (int rc = scan_check(a)) return rc;
  const GeoB g = make_geo_b(a, 0);
  const size_t smem = sizeof(float) * (size_t)g.total;
  RL_CHECK_ARG(smem <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  RL_CUDA(cudaFuncSetAttribute(rssm_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RL_CUDA(cudaMemsetAsync(a.workspace, 0, 256, st));
  void* kargs[] = {(void*)args, (void*)grads};
  RL_CUDA(cudaLaunchCooperativeKernel((void*)rssm_scan_bwd_kernel, dim3(SCAN_G), dim3(SCAN_NT), kargs, smem, st));
  return B200RL_OK;
}

extern "C" int b200rl_rssm_scan_error(const void* workspace, cudaStream_t st) {
  int flag = 0;
  RL_CUDA(cudaMemcpyAsync(&flag, (const char*)workspace + 64, sizeof(int), cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));
  return flag;
}
