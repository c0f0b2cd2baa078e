// Persistent RSSM scan (forward): all T steps of RSSM.dynamic in ONE cooperative kernel.
//
// Replaces the Python loop `for i in range(sequence_length): rssm.dynamic(...)`
// (sheeprl/algos/dreamer_v3/dreamer_v3.py:131-145 -> agent.py:396-435: is_first masking, RecurrentModel +
// LayerNormGRUCell models.py:396-403, transition / representation MLPs, unimix, straight-through sampling).
//
// Design (B200): the batch is tiny (B <= 16 rows) and the steps are strictly sequential, so the scan is
// latency-bound.  One CTA per SM (G CTAs, cooperative launch) owns a fixed slice of OUTPUT COLUMNS of every
// weight matrix and keeps that slice resident in shared memory for the whole scan (S size: ~82 KB of weights
// per CTA, 23 MB over the grid — weights are read from HBM exactly once per scan instead of once per step).
// Per step: 4 dependent skinny GEMM stages separated by grid barriers; the recurrent/stochastic state of
// all B rows is staged in shared memory ([16][K] row block) for each stage.  Each warp takes a
// (4-column group) x (K-slice) work item: 16 rows x 4 cols accumulators per lane over its k's, then a
// 62-shuffle reduce-scatter.  The previous stochastic state is one-hot per group, so z_{t-1} W_in^T is a
// gather of S columns (32x fewer FLOPs than the dense product, bit-for-bit the same terms).  LayerNorm over
// the 3R-wide GRU pre-activation is a two-level Chan/Welford merge of per-CTA (mean, M2) partials; the 32
// classes of a categorical map onto the 32 lanes of a warp (softmax / unimix / argmax by shuffles).
#include <cooperative_groups.h>

#include "common.cuh"
#include "b200rl.h"

namespace {

constexpr int SCAN_G = 128;    // CTAs (one per SM; 148 SMs available)
constexpr int SCAN_NT = 512;   // threads per CTA (16 warps)
constexpr int SCAN_NW = SCAN_NT / 32;
constexpr int MAXB = 16;

struct Geo {          // per-CTA column ownership (groups of 4 columns, interleaved over CTAs)
  int ngx, ngh, ngt, ngr;   // number of owned 4-col groups of Dx, R, Dt, Dr
  int KS;                   // row stride of the X block in smem (odd)
  int oWin, oWg, oWt1, oWr1, oX, oOut, oMisc, total;  // smem offsets in floats
};

__host__ __device__ inline int owned_groups(int width, int cta) {
  const int groups = (width + 3) / 4;
  return (groups > cta) ? (groups - cta + SCAN_G - 1) / SCAN_G : 0;
}

__host__ __device__ inline Geo make_geo(const b200rl_rssm_scan_args& a, int cta) {
  Geo g;
  const int Z = a.S * a.D;
  g.ngx = owned_groups(a.Dx, cta);
  g.ngh = owned_groups(a.R, cta);
  g.ngt = owned_groups(a.Dt, cta);
  g.ngr = owned_groups(a.Dr, cta);
  int kmax = a.R + a.Dx;
  if (a.Dt > kmax) kmax = a.Dt;
  if (a.Dr > kmax) kmax = a.Dr;
  g.KS = kmax | 1;
  // sizes are computed for CTA 0 (the largest owner) so that every CTA uses the same layout
  const int mx = owned_groups(a.Dx, 0), mh = owned_groups(a.R, 0), mt = owned_groups(a.Dt, 0), mr = owned_groups(a.Dr, 0);
  int o = 0;
  g.oWin = o;  o += mx * 4 * (Z + a.A);
  g.oWg = o;   o += mh * 12 * (a.R + a.Dx);
  g.oWt1 = o;  o += mt * 4 * a.R;
  g.oWr1 = o;  o += mr * 4 * a.R;
  g.oX = o;    o += MAXB * g.KS;
  int outc = mh * 12;
  if (mx * 4 > outc) outc = mx * 4;
  if ((mt + mr) * 4 > outc) outc = (mt + mr) * 4;
  if (32 > outc) outc = 32;
  g.oOut = o;  o += MAXB * outc;
  g.oMisc = o; o += 4 * MAXB + 64;
  g.total = o;
  return g;
}

struct Workspace {
  unsigned* counter;   // grid barrier arrivals
  int* error;
  float* stats;        // [2][MAXB][SCAN_G][2]
  int* zidx;           // [T][B][S]
};

__device__ inline Workspace carve(void* ws, int T, int B, int S) {
  Workspace w;
  char* p = (char*)ws;
  w.counter = (unsigned*)p;
  w.error = (int*)(p + 64);
  w.stats = (float*)(p + 256);
  w.zidx = (int*)(p + 256 + sizeof(float) * 2 * MAXB * SCAN_G * 2);
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
    __threadfence();
    atomicAdd(w.counter, 1u);
    long long t0 = clock64();
    while (ld_acquire(w.counter) < target) {
      if (clock64() - t0 > 4000000000LL) {  // ~2 s: never hang the device; flag and bail out
        atomicExch(w.error, 1);
        break;
      }
      if (ld_acquire((const unsigned*)w.error) != 0u) break;  // another CTA gave up: do not wait for it
    }
    __threadfence();
  }
  __syncthreads();
}

// true (uniformly over the CTA) if any CTA flagged a barrier time-out
__device__ __forceinline__ bool scan_failed(const Workspace& w, int* flag_smem) {
  if (threadIdx.x == 0) *flag_smem = (int)ld_acquire((const unsigned*)w.error);
  __syncthreads();
  const bool f = *flag_smem != 0;
  __syncthreads();
  return f;
}

// v[64] = acc[16 rows][4 cols] per lane; sum across the 32 lanes; lane ends up owning 2 consecutive values
// starting at `base` (returned): 62 shuffles instead of 320.
__device__ __forceinline__ int reduce_scatter64(float (&v)[64], int lane) {
#pragma unroll
  for (int off = 16, n = 64; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float lo = v[i], hi = v[i + n / 2];
      const float send = upper ? lo : hi;
      const float keep = upper ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
}

// out[b][c0 + j] += sum_{k in [k0,k1)} X[b][k] * Wrow_j[k]   for the 4 columns of one group.
// X: smem [MAXB][KS]; w0..w3: pointers to the 4 weight rows (smem or global), nullptr => column masked.
// One warp per call; lanes stride over k.  `out` is smem [MAXB][ldo] accumulated with shared atomics.
__device__ __forceinline__ void warp_item(const float* __restrict__ X, int KS, const float* w0, const float* w1,
                                          const float* w2, const float* w3, int k0, int k1, float* out, int ldo,
                                          int c0, int lane) {
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  for (int k = k0 + lane; k < k1; k += 32) {
    const float a0 = w0 ? w0[k] : 0.f, a1 = w1 ? w1[k] : 0.f, a2 = w2 ? w2[k] : 0.f, a3 = w3 ? w3[k] : 0.f;
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      const float x = X[b * KS + k];
      acc[b * 4 + 0] = fmaf(x, a0, acc[b * 4 + 0]);
      acc[b * 4 + 1] = fmaf(x, a1, acc[b * 4 + 1]);
      acc[b * 4 + 2] = fmaf(x, a2, acc[b * 4 + 2]);
      acc[b * 4 + 3] = fmaf(x, a3, acc[b * 4 + 3]);
    }
  }
  const int base = reduce_scatter64(acc, lane);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = base + i, b = idx >> 2, j = idx & 3;
    atomicAdd(&out[b * ldo + c0 + j], acc[i]);
  }
}

// LayerNorm(+SiLU) of row `b` held in smem (length n), in place; one warp.
__device__ __forceinline__ void warp_ln_row(float* x, int n, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float eps, bool silu, int lane) {
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s += x[k];
  const float mu = warp_sum(s) / (float)n;
  float v = 0.f;
  for (int k = lane; k < n; k += 32) { const float d = x[k] - mu; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / (float)n + eps);
  for (int k = lane; k < n; k += 32) {
    float o = (x[k] - mu) * rstd * gamma[k] + beta[k];
    if (silu) o = siluf_(o);
    x[k] = o;
  }
}

__global__ void __launch_bounds__(SCAN_NT, 1) rssm_scan_fwd_kernel(const b200rl_rssm_scan_args a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const int T = a.T, B = a.B, S = a.S, D = a.D, Z = S * D, R = a.R, A = a.A, Dx = a.Dx, Dt = a.Dt, Dr = a.Dr;
  const int KIN = Z + A, KG = R + Dx;
  const Geo g = make_geo(a, cta);
  const Workspace ws = carve(a.workspace, T, B, S);
  float* Win = sm + g.oWin;     // [ngx*4][KIN]
  float* Wg = sm + g.oWg;       // [ngh*12][KG]  rows: (group, part r/c/u, col-in-group)
  float* Wt1 = sm + g.oWt1;     // [ngt*4][R]
  float* Wr1 = sm + g.oWr1;     // [ngr*4][R]
  float* X = sm + g.oX;         // [MAXB][KS]
  float* OUT = sm + g.oOut;     // [MAXB][ldo]
  float* misc = sm + g.oMisc;   // [0,16): first flags; [16,32): mean; [32,48): rstd; [64,..): z0idx (as int)
  int* z0idx = (int*)(misc + 64);
  const int KS = g.KS;
  unsigned bar_target = 0;

  // ---------------- prologue: weight slices -> shared memory (read from HBM once per scan)
  for (int gi = 0; gi < g.ngx; ++gi)
    for (int j = 0; j < 4; ++j) {
      const int col = (cta + gi * SCAN_G) * 4 + j;
      float* dst = Win + (gi * 4 + j) * KIN;
      for (int k = tid; k < KIN; k += SCAN_NT) dst[k] = (col < Dx) ? a.W_in[(size_t)col * KIN + k] : 0.f;
    }
  for (int gi = 0; gi < g.ngh; ++gi)
    for (int part = 0; part < 3; ++part)
      for (int j = 0; j < 4; ++j) {
        const int col = (cta + gi * SCAN_G) * 4 + j;
        float* dst = Wg + ((gi * 3 + part) * 4 + j) * KG;
        for (int k = tid; k < KG; k += SCAN_NT) dst[k] = (col < R) ? a.W_g[(size_t)(part * R + col) * KG + k] : 0.f;
      }
  for (int gi = 0; gi < g.ngt; ++gi)
    for (int j = 0; j < 4; ++j) {
      const int col = (cta + gi * SCAN_G) * 4 + j;
      float* dst = Wt1 + (gi * 4 + j) * R;
      for (int k = tid; k < R; k += SCAN_NT) dst[k] = (col < Dt) ? a.W_t1[(size_t)col * R + k] : 0.f;
    }
  for (int gi = 0; gi < g.ngr; ++gi)
    for (int j = 0; j < 4; ++j) {
      const int col = (cta + gi * SCAN_G) * 4 + j;
      float* dst = Wr1 + (gi * 4 + j) * R;
      for (int k = tid; k < R; k += SCAN_NT) dst[k] = (col < Dr) ? a.W_r1[(size_t)col * a.ld_wr1 + k] : 0.f;
    }
  if (wid == 0) {  // index of the learned initial posterior (one-hot `z0`)
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
          // z_in = (1-f) z_prev + f z0 : with f in {0,1} this is one column; keep the mask-multiply form
          const int i0 = z0idx[gq];
          if (t > 0) {
            const int ip = __ldcg(&ws.zidx[((size_t)(t - 1) * B + b) * S + gq]);
            acc = fmaf(1.f - f, wrow[gq * D + ip], acc);
          }
          acc = fmaf(f, wrow[gq * D + i0], acc);
        }
        for (int q = 0; q < A; ++q) acc = fmaf((1.f - f) * a.actions[(row0 + b) * A + q], wrow[Z + q], acc);
        a.x_pre[(row0 + b) * Dx + col] = acc;
      }
      // dense saves for the deferred weight-gradient GEMMs: z_in / a_in rows, spread over the CTAs
      for (int e = cta * SCAN_NT + tid; e < B * Z; e += SCAN_G * SCAN_NT) {
        const int b = e / Z, k = e - b * Z;
        const int gq = k / D, d = k - gq * D;
        const float f = fl[b];
        float zp = 0.f;
        if (t > 0) zp = (__ldcg(&ws.zidx[((size_t)(t - 1) * B + b) * S + gq]) == d) ? 1.f : 0.f;
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
        for (int k = lane; k < R; k += 32) {
          const float hp = (t > 0) ? __ldcg(&a.latent[(row0 - B + b) * a.ld_lat + Z + k]) : 0.f;
          xr[k] = (1.f - f) * hp + f * a.h0[k];
        }
        for (int k = lane; k < Dx; k += 32) xr[R + k] = __ldcg(&a.x_pre[(row0 + b) * Dx + k]);
        __syncwarp();
        warp_ln_row(xr + R, Dx, a.lnx_g, a.lnx_b, a.eps, true, lane);
        if (cta == ((t + 1) % SCAN_G)) {
          for (int k = lane; k < R; k += 32) a.h_in[(row0 + b) * R + k] = xr[k];
          for (int k = lane; k < Dx; k += 32) a.x_act[(row0 + b) * Dx + k] = xr[R + k];
        }
      } else {
        for (int k = lane; k < KG; k += 32) xr[k] = 0.f;
      }
    }
    const int ldo2 = g.ngh * 12;
    for (int e = tid; e < MAXB * ldo2; e += SCAN_NT) OUT[e] = 0.f;
    __syncthreads();
    if (g.ngh > 0) {
      const int ncg = g.ngh * 3;                      // 4-column groups to compute
      int ks = SCAN_NW / ncg;                         // K-slices per group
      if (ks < 1) ks = 1;
      const int kchunk = ((KG + ks - 1) / ks + 31) / 32 * 32;
      for (int item = wid; item < ncg * ks; item += SCAN_NW) {
        const int cg = item % ncg, sl = item / ncg;
        const int k0 = sl * kchunk, k1 = min(KG, k0 + kchunk);
        if (k0 >= k1) continue;
        const float* wr = Wg + (size_t)cg * 4 * KG;
        warp_item(X, KS, wr, wr + KG, wr + 2 * KG, wr + 3 * KG, k0, k1, OUT, ldo2, cg * 4, lane);
      }
    }
    __syncthreads();
    // save g_pre columns; per-row partial statistics (mean, M2) over the owned valid columns
    {
      const int par = t & 1;
      for (int b = wid; b < B; b += SCAN_NW) {
        float s = 0.f;
        int cnt = 0;
        for (int c = lane; c < ldo2; c += 32) {
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
        for (int c = lane; c < ldo2; c += 32) {
          const int gi = c / 12, j = c & 3;
          const int col = (cta + gi * SCAN_G) * 4 + j;
          if (col < R) { const float d = OUT[b * ldo2 + c] - mean; m2 = fmaf(d, d, m2); }
        }
        m2 = warp_sum(m2);
        if (lane == 0) {
          float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + cta) * 2;
          st[0] = mean;
          st[1] = m2;
        }
      }
    }
    grid_barrier(ws, bar_target);  // B2: partial LN statistics complete

    // ============ stage 2b: merge statistics, LayerNorm, GRU gate -> h_t for the owned columns
    {
      const int par = t & 1;
      const int groupsR = (R + 3) / 4;
      for (int b = wid; b < B; b += SCAN_NW) {
        float sm_ = 0.f;
        for (int c = lane; c < SCAN_G; c += 32) {
          int nc = 0;
          for (int gi = c; gi < groupsR; gi += SCAN_G) nc += min(4, R - gi * 4);
          const float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + c) * 2;
          sm_ += (float)(3 * nc) * __ldcg(st);
        }
        const float mean = warp_sum(sm_) / (float)(3 * R);
        float m2 = 0.f;
        for (int c = lane; c < SCAN_G; c += 32) {
          int nc = 0;
          for (int gi = c; gi < groupsR; gi += SCAN_G) nc += min(4, R - gi * 4);
          const float* st = ws.stats + (((size_t)par * MAXB + b) * SCAN_G + c) * 2;
          const float d = __ldcg(st) - mean;
          m2 += __ldcg(st + 1) + (float)(3 * nc) * d * d;
        }
        m2 = warp_sum(m2);
        if (lane == 0) {
          misc[16 + b] = mean;
          misc[32 + b] = rsqrtf(m2 / (float)(3 * R) + a.eps);
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
    for (int b = wid; b < MAXB; b += SCAN_NW) {
      float* xr = X + b * KS;
      for (int k = lane; k < R; k += 32) xr[k] = (b < B) ? __ldcg(&a.latent[(row0 + b) * a.ld_lat + Z + k]) : 0.f;
    }
    const int ldo3 = (g.ngt + g.ngr) * 4;
    for (int e = tid; e < MAXB * ldo3; e += SCAN_NT) OUT[e] = 0.f;
    __syncthreads();
    {
      const int ncg = g.ngt + g.ngr;
      if (ncg > 0) {
        int ks = SCAN_NW / ncg;
        if (ks < 1) ks = 1;
        const int kchunk = ((R + ks - 1) / ks + 31) / 32 * 32;
        for (int item = wid; item < ncg * ks; item += SCAN_NW) {
          const int cg = item % ncg, sl = item / ncg;
          const int k0 = sl * kchunk, k1 = min(R, k0 + kchunk);
          if (k0 >= k1) continue;
          const float* wr = (cg < g.ngt) ? (Wt1 + (size_t)cg * 4 * R) : (Wr1 + (size_t)(cg - g.ngt) * 4 * R);
          warp_item(X, KS, wr, wr + R, wr + 2 * R, wr + 3 * R, k0, k1, OUT, ldo3, cg * 4, lane);
        }
      }
      __syncthreads();
      for (int e = tid; e < B * ldo3; e += SCAN_NT) {
        const int b = e / ldo3, c = e - b * ldo3;
        const int cg = c >> 2, j = c & 3;
        if (cg < g.ngt) {
          const int col = (cta + cg * SCAN_G) * 4 + j;
          if (col < Dt) a.tr_pre[(row0 + b) * Dt + col] = OUT[b * ldo3 + c];
        } else {
          const int col = (cta + (cg - g.ngt) * SCAN_G) * 4 + j;
          if (col < Dr) a.rp_pre[(row0 + b) * Dr + col] = OUT[b * ldo3 + c] + a.pe[(row0 + b) * Dr + col];
        }
      }
    }
    grid_barrier(ws, bar_target);  // B4: tr_pre / rp_pre complete

    // ============ stage 4: logits of one categorical group per unit, unimix, sample (posterior only)
    for (int u = cta; u < n_units; u += SCAN_G) {
      const bool post = u < S;
      const int gq = post ? u : u - S;
      const int Dh = post ? Dr : Dt;
      const float* pre = post ? a.rp_pre : a.tr_pre;
      float* act_save = post ? a.rp_act : a.tr_act;
      const float* lg_ = post ? a.lnr_g : a.lnt_g;
      const float* lb_ = post ? a.lnr_b : a.lnt_b;
      const float* W2 = post ? a.W_r2 : a.W_t2;   // [Z][Dh], rows gq*D .. gq*D+D-1 (read from L2 every step)
      const float* b2 = post ? a.b_r2 : a.b_t2;
      __syncthreads();
      for (int b = wid; b < MAXB; b += SCAN_NW) {
        float* xr = X + b * KS;
        if (b < B) {
          for (int k = lane; k < Dh; k += 32) xr[k] = __ldcg(&pre[(row0 + b) * Dh + k]);
          __syncwarp();
          warp_ln_row(xr, Dh, lg_, lb_, a.eps, true, lane);
          if (gq == 0)
            for (int k = lane; k < Dh; k += 32) act_save[(row0 + b) * Dh + k] = xr[k];
        } else {
          for (int k = lane; k < Dh; k += 32) xr[k] = 0.f;
        }
      }
      const int ncg = (D + 3) / 4;
      const int ldo4 = ncg * 4;
      for (int e = tid; e < MAXB * ldo4; e += SCAN_NT) OUT[e] = 0.f;
      __syncthreads();
      {
        int ks = SCAN_NW / ncg;
        if (ks < 1) ks = 1;
        const int kchunk = ((Dh + ks - 1) / ks + 31) / 32 * 32;
        for (int item = wid; item < ncg * ks; item += SCAN_NW) {
          const int cg = item % ncg, sl = item / ncg;
          const int k0 = sl * kchunk, k1 = min(Dh, k0 + kchunk);
          if (k0 >= k1) continue;
          const float* wr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) wr[j] = (cg * 4 + j < D) ? W2 + (size_t)(gq * D + cg * 4 + j) * Dh : nullptr;
          warp_item(X, KS, wr[0], wr[1], wr[2], wr[3], k0, k1, OUT, ldo4, cg * 4, lane);
        }
      }
      __syncthreads();
      // one warp per row: the D classes of the group live on the lanes (D <= 32)
      for (int b = wid; b < B; b += SCAN_NW) {
        const bool on = lane < D;
        const float raw = on ? OUT[b * ldo4 + lane] + b2[gq * D + lane] : -INFINITY;
        const size_t o = (row0 + b) * Z + (size_t)gq * D + lane;
        float mx = warp_max(raw);
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
    if (scan_failed(ws, (int*)(misc + 48))) return;
  }
}

}  // namespace

extern "C" long long b200rl_rssm_scan_workspace_bytes(int T, int B, int S) {
  return 256 + (long long)sizeof(float) * 2 * MAXB * SCAN_G * 2 + (long long)sizeof(int) * T * B * S + 256;
}

extern "C" int b200rl_rssm_scan_fwd(const b200rl_rssm_scan_args* args, cudaStream_t st) {
  RL_CHECK_ARG(args, "null args");
  const b200rl_rssm_scan_args& a = *args;
  RL_CHECK_ARG(a.B >= 1 && a.B <= MAXB, "persistent scan supports batch <= 16 rows per rank");
  RL_CHECK_ARG(a.D >= 1 && a.D <= 32, "persistent scan supports <= 32 classes per categorical");
  RL_CHECK_ARG(a.T >= 1 && a.S >= 1 && a.S <= 64, "bad T / S (S <= 64)");
  RL_CHECK_ARG(a.workspace && a.workspace_bytes >= b200rl_rssm_scan_workspace_bytes(a.T, a.B, a.S), "workspace too small");
  const Geo g = make_geo(a, 0);
  const size_t smem = sizeof(float) * (size_t)g.total;
  RL_CHECK_ARG(smem <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  RL_CUDA(cudaFuncSetAttribute(rssm_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RL_CUDA(cudaMemsetAsync(a.workspace, 0, 256, st));
  void* kargs[] = {(void*)args};
  RL_CUDA(cudaLaunchCooperativeKernel((void*)rssm_scan_fwd_kernel, dim3(SCAN_G), dim3(SCAN_NT), kargs, smem, st));
  return B200RL_OK;
}

extern "C" int b200rl_rssm_scan_error(const void* workspace, cudaStream_t st) {
  int flag = 0;
  RL_CUDA(cudaMemcpyAsync(&flag, (const char*)workspace + 64, sizeof(int), cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));
  return flag;
}
