// Persistent RSSM scan: all T steps of RSSM.dynamic (forward) and their BPTT (backward), each as ONE
// cooperative kernel.
//
// Replaces the Python loop `for i in range(sequence_length): rssm.dynamic(...)`
// (sheeprl/algos/dreamer_v3/dreamer_v3.py:131-145 -> agent.py:396-435: is_first masking, RecurrentModel +
// LayerNormGRUCell models.py:396-403, representation MLP, unimix, straight-through sampling) and the autograd
// replay of it inside `fabric.backward(rec_loss)` (dreamer_v3.py:191).
//
// What is on the recurrence and what is not.  Only the POSTERIOR feeds the next step (z_t -> x_{t+1} -> h_{t+1});
// the prior (transition model on h_t, agent.py:433) is a function of the finished h sequence, so its forward and
// its backward (its gradient comes from the KL term only) are batched tensor-core products over all T*B rows
// OUTSIDE this kernel (engine.py: _prior_forward / _prior_backward).  The kernels below carry the dependent
// chain only:
//   forward   z_{t-1} --gather W_in--> x (LN, SiLU) --W_g--> gates (LN over 3R) --> h_t --W_r1--> rp (LN, SiLU)
//             --W_r2--> logits --unimix, sample--> z_t
//   backward  the reverse chain, LayerNorm corrections applied by the CONSUMER of each gradient row.
//
// Design (B200).  The batch is tiny (B <= 16 rows), the steps strictly sequential: the scan is latency-bound, so
// the design minimises the number and the cost of cross-SM hand-offs per step.
//   * 128 co-resident CTAs (one per SM, cooperative launch), each owning a fixed slice of OUTPUT COLUMNS of every
//     weight matrix, resident in shared memory for the whole scan (weights are read from HBM once per scan).
//   * No grid barriers.  Every cross-CTA hand-off is a flag-carrying data exchange ("LL": each 8-byte element is
//     {fp32 value, 32-bit step tag}, written with one 64-bit store and polled with 64-bit loads straight from L2):
//     the consumer spins on the very data it needs, so a hand-off costs one L2 store + one L2 load latency instead of
//     release-fence + atomic counter + acquire + a dependent load.  Five hand-offs per forward step (z indices, x rows,
//     LayerNorm partial statistics, h rows, rp rows), four per backward step.
//   * Work that does not depend on the newest hand-off is issued before waiting for it (the h-part of the GRU product
//     runs while the x rows are still being built; input prefetches at step start).
//   * z_{t-1} is one-hot per group, so z W_in^T is a gather of S rows of W_in^T (read from L2 by the CTA that owns
//     the batch row: it then normalises the full row once, instead of every CTA normalising every row).
//   * Products: a warp takes a (4-column group) x (K slice) item: 16 rows x 4 columns of accumulators per lane,
//     128-bit shared loads along K, a 62-shuffle reduce-scatter; K-slice partials are summed in a fixed order
//     (bit-reproducible).  The 32 classes of a categorical sit on the 32 lanes of a warp.
#include "b200rl.h"
#include "common.cuh"

namespace {

constexpr int SCAN_G = 128;    // CTAs (one per SM; 148 SMs available)
constexpr int SCAN_NT = 256;   // threads per CTA
constexpr int SCAN_NW = SCAN_NT / 32;
constexpr int MAXB = 16;
constexpr int OWN_STRIDE = SCAN_G / MAXB;   // batch row b is owned by CTA b * OWN_STRIDE
constexpr int MAXRPU = 8;      // rows of one sampling unit (S <= 64 groups over 128 CTAs -> >= 2 row splits)
constexpr int LL_UNROLL = 8;   // 16-byte loads in flight per thread while receiving rows

typedef unsigned long long u64;

__host__ __device__ inline int r4(int x) { return (x + 3) / 4 * 4; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }

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

// ---------------------------------------------------------------------------------------------------------------
// workspace: [header 256 B | profile | LL exchange region (zeroed at every launch) | saves shared by fwd and bwd]
// ---------------------------------------------------------------------------------------------------------------
struct Workspace {
  int* error;
  long long* prof;       // [2][32] cycle counters of CTA 0 (a row owner) and CTA 1 (profiling aid)
  u64* ll;               // LL region base
  int* zidx;             // [T][B][S] sampled class per group
  float* ln_stats;       // [3][T*B][2] (mean, rstd) of the x / g / representation LayerNorms
};

struct LLGeo {           // offsets (in u64 elements) inside the LL region; every buffer is double-buffered by step parity
  size_t z, x, s, h, r, a, b, c, d, sb, sc, sd, total;
};

__host__ __device__ inline LLGeo make_ll(int S, int Dx, int R, int Dr, int Z) {
  LLGeo g;
  size_t o = 0;
  g.z = o; o += 2 * (size_t)MAXB * S;            // forward: sampled class indices
  g.x = o; o += 2 * (size_t)MAXB * Dx;           //          x_act rows
  g.s = o; o += 2 * (size_t)MAXB * SCAN_G * 2;   //          LayerNorm partial statistics (also backward row sums)
  g.h = o; o += 2 * (size_t)MAXB * R;            //          h rows
  g.r = o; o += 2 * (size_t)MAXB * Dr;           //          rp_pre rows
  g.a = o; o += 2 * (size_t)MAXB * Z;            // backward: d_post_raw rows
  g.b = o; o += 2 * (size_t)MAXB * Dr;           //           dxh of the representation LayerNorm
  g.c = o; o += 2 * (size_t)MAXB * 3 * R;        //           dxh of the GRU LayerNorm
  g.d = o; o += 2 * (size_t)MAXB * Dx;           //           dxh of the x LayerNorm
  g.sb = o; o += 2 * (size_t)MAXB * SCAN_G * 2;  //           per-CTA row sums (sum dxh, sum dxh*xh) of the three
  g.sc = o; o += 2 * (size_t)MAXB * SCAN_G * 2;
  g.sd = o; o += 2 * (size_t)MAXB * SCAN_G * 2;
  g.total = o;
  return g;
}

constexpr size_t WS_HEADER = 256, WS_PROF = 2 * 32 * sizeof(long long);

__host__ __device__ inline size_t ws_ll_bytes(int S, int D, int Dx, int R, int Dr) {
  return make_ll(S, Dx, R, Dr, S * D).total * sizeof(u64);
}
__host__ __device__ inline size_t ws_bytes(int T, int B, int S, int D, int Dx, int R, int Dr) {
  return WS_HEADER + WS_PROF + ws_ll_bytes(S, D, Dx, R, Dr) + sizeof(int) * (size_t)T * B * S +
         sizeof(float) * 3 * (size_t)T * B * 2 + 256;
}

__host__ __device__ inline Workspace carve(void* ws, int T, int B, int S, int D, int Dx, int R, int Dr) {
  Workspace w;
  char* p = (char*)ws;
  w.error = (int*)(p + 64);
  p += WS_HEADER;
  w.prof = (long long*)p;   p += WS_PROF;
  w.ll = (u64*)p;           p += ws_ll_bytes(S, D, Dx, R, Dr);
  w.zidx = (int*)p;         p += sizeof(int) * (size_t)T * B * S;
  w.ln_stats = (float*)p;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// LL exchange primitives
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ll_pack(float v, unsigned tag) { return ((u64)tag << 32) | (u64)__float_as_uint(v); }
__device__ __forceinline__ void ll_store(u64* p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(ll_pack(v, tag)) : "memory");
}
__device__ __forceinline__ void ll_store2(u64* p, float v0, float v1, unsigned tag) {   // p 16-byte aligned
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(ll_pack(v0, tag)), "l"(ll_pack(v1, tag)) : "memory");
}
__device__ __forceinline__ u64 ll_load(const u64* p) {
  u64 x;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void ll_load2(const u64* p, u64& a, u64& b) {                // p 16-byte aligned
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

struct Spin {            // bounded polling: a lost hand-off must never hang the device
  int* error;
  unsigned polls;
  long long t0;
  bool dead;
  __device__ __forceinline__ void init(int* e) { error = e; polls = 0; t0 = 0; dead = false; }
  // called on every failed poll; returns true when the wait must be abandoned
  __device__ __forceinline__ bool fail() {
    if (dead) return true;
    if ((++polls & 255u) == 0u) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > 4000000000LL) { atomicExch(error, 1); dead = true; }               // ~2 s
      else if (*(volatile int*)error != 0) dead = true;                                        // another CTA gave up
    }
    return dead;
  }
};

__device__ __forceinline__ float ll_wait(const u64* p, unsigned tag, Spin& sp) {
  u64 v = ll_load(p);
  while ((unsigned)(v >> 32) != tag) {
    if (sp.fail()) break;
    v = ll_load(p);
  }
  return __uint_as_float((unsigned)v);
}

// Receives rows x n values (n even; LL rows of stride ss elements, 16-byte aligned) into shared memory rows of stride ds.
// LL_UNROLL 16-byte loads are in flight per thread before the first tag is checked.
__device__ __forceinline__ void ll_recv(float* dst, int ds, const u64* src, int ss, int rows, int n, unsigned tag, int tid, Spin& sp) {
  const int half = n >> 1, total = rows * half;
  for (int e0 = tid; e0 < total; e0 += SCAN_NT * LL_UNROLL) {
    u64 a[LL_UNROLL], b[LL_UNROLL];
#pragma unroll
    for (int u = 0; u < LL_UNROLL; ++u) {
      const int e = e0 + u * SCAN_NT;
      if (e < total) ll_load2(src + (size_t)(e / half) * ss + 2 * (size_t)(e % half), a[u], b[u]);
    }
#pragma unroll
    for (int u = 0; u < LL_UNROLL; ++u) {
      const int e = e0 + u * SCAN_NT;
      if (e < total) {
        const int row = e / half, k = (e - row * half) * 2;
        while ((unsigned)(a[u] >> 32) != tag || (unsigned)(b[u] >> 32) != tag) {
          if (sp.fail()) break;
          ll_load2(src + (size_t)row * ss + k, a[u], b[u]);
        }
        *reinterpret_cast<float2*>(dst + row * ds + k) =
            make_float2(__uint_as_float((unsigned)a[u]), __uint_as_float((unsigned)b[u]));
      }
    }
  }
}

// fast transcendental form for SiLU; rel. error ~1e-6
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

// part[b][c0 + j] = sum_{k in [k0,k1)} X[b][k] * W_j[k] for NR rows and the 4 columns (weight rows w + j*wst).
// X: smem [NR][xs] (xs % 4 == 0, zero padded); weight rows in smem (stride wst % 4 == 0, zero padded);
// k0, k1 multiples of 4.  One warp; a lane handles 4 consecutive k per 128-wide sweep (LDS.128).
template <int NR>
__device__ __forceinline__ void warp_item(const float* __restrict__ X, int xs, const float* __restrict__ w, int wst,
                                          int k0, int k1, float* part, int ldp, int c0, int lane) {
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
      const float4 x = *reinterpret_cast<const float4*>(X + b * xs + k);
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
    part[b * ldp + c0 + j] = acc[i];
  }
}

// number of K slices a product of `ngroups` column groups over K is cut into (all SCAN_NW warps busy when possible)
__device__ __forceinline__ int k_slices(int ngroups, int K, int& kchunk) {
  const int Kp = r4(K);
  int ks = imax(1, SCAN_NW / imax(ngroups, 1));
  kchunk = ((Kp + ks - 1) / ks + 127) / 128 * 128;
  return (Kp + kchunk - 1) / kchunk;
}

// Runs all (column group, K slice) items of one product over the CTA's warps.  Slice s writes its partial sums to
// PART[s][b][cbase + 4*group + j] (every element written by exactly one lane); the caller sums the slices in order.
// `wbase`: first weight row of group 0; group g starts at wbase + g*4*wst.  Returns the number of slices.
// No barrier inside: the caller synchronises before (X complete) and after (PART complete).
__device__ __forceinline__ int product(const float* X, int xs, const float* wbase, int wst, int ngroups, int K,
                                       float* PART, int ldp, int cbase, int tid) {
  if (ngroups <= 0) return 0;
  const int lane = tid & 31, wid = tid >> 5;
  int kchunk;
  const int ks = k_slices(ngroups, K, kchunk), Kp = r4(K);
  for (int item = wid; item < ngroups * ks; item += SCAN_NW) {
    const int cg = item % ngroups, sl = item / ngroups;
    const int k0 = sl * kchunk, k1 = imin(Kp, k0 + kchunk);
    warp_item<MAXB>(X, xs, wbase + (size_t)cg * 4 * wst, wst, k0, k1, PART + (size_t)sl * MAXB * ldp, ldp, cbase + cg * 4, lane);
  }
  return ks;
}

__device__ __forceinline__ float part_sum(const float* PART, int ldp, int ks, int b, int c) {
  float v = 0.f;
  for (int s = 0; s < ks; ++s) v += PART[((size_t)s * MAXB + b) * ldp + c];
  return v;
}

// Copies 4 rows (row0..row0+3) of a row-major [rows][ld] weight, columns [c0, c0+K), into smem rows of stride `wst` at
// offset koff, zero padded to Kp.
__device__ __forceinline__ void load_rows4(float* dst, int wst, int koff, const float* W, size_t ld, int row0, int nrows_valid,
                                           int c0, int K, int Kp, int tid) {
  for (int j = 0; j < 4; ++j)
    for (int k = tid; k < Kp; k += SCAN_NT)
      dst[j * wst + koff + k] = (row0 + j < nrows_valid && k < K) ? W[(size_t)(row0 + j) * ld + c0 + k] : 0.f;
}
// Copies 4 COLUMNS (col0..col0+3) of a row-major [K][ld] weight into 4 smem rows (transposed slice) at offset koff.
__device__ __forceinline__ void load_cols4(float* dst, int wst, const float* W, size_t ld, int col0, int ncols_valid,
                                           int K, int koff, int tid) {
  for (int e = tid; e < K * 4; e += SCAN_NT) {
    const int k = e >> 2, j = e & 3;
    dst[j * wst + koff + k] = (col0 + j < ncols_valid) ? W[(size_t)k * ld + col0 + j] : 0.f;
  }
}

__device__ __forceinline__ void prof_mark(long long* prof, int slot, long long& last, bool on) {
  if (on) {
    const long long now = clock64();
    prof[slot] += now - last;
    last = now;
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
struct GeoF {
  int ngh, ngr;                 // owned 4-column groups of R (GRU output columns) and Dr (representation layer 1)
  int sR, sDx, sDr, wgst, w2st; // padded widths; smem row strides of the W_g / W_r2 slices
  int nsplit, rpu;              // sampling units: S groups x nsplit row blocks of rpu rows
  int unit_g, unit_r0, unit_nr; // this CTA's unit: group (-1: none), first row, row count
  int owner_row;                // batch row whose x this CTA builds (-1: none)
  int ldp;                      // row length of the product partial buffers
  int oWg, oWr1, oW2, oXh, oXx, oPart, oAcc, oX0, oMisc, oInt, total;
};

__host__ __device__ inline GeoF make_geo_f(const b200rl_rssm_scan_args& a, int cta) {
  GeoF g;
  g.ngh = owned_groups(a.R, cta);
  g.ngr = owned_groups(a.Dr, cta);
  g.sR = r4(a.R); g.sDx = r4(a.Dx); g.sDr = r4(a.Dr);
  g.wgst = g.sR + g.sDx;
  g.w2st = g.sDr + 4;           // +4: the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
  g.nsplit = imax(1, imin(a.B, SCAN_G / a.S));
  g.rpu = (a.B + g.nsplit - 1) / g.nsplit;
  g.unit_g = -1; g.unit_r0 = 0; g.unit_nr = 0;
  if (cta < a.S * g.nsplit) {
    const int sp = cta / a.S;
    g.unit_r0 = sp * g.rpu;
    g.unit_nr = imin(g.rpu, a.B - g.unit_r0);
    if (g.unit_nr > 0) g.unit_g = cta % a.S; else g.unit_nr = 0;
  }
  g.owner_row = (cta % OWN_STRIDE == 0 && cta / OWN_STRIDE < a.B) ? cta / OWN_STRIDE : -1;
  const int mh = owned_groups(a.R, 0), mr = owned_groups(a.Dr, 0);
  g.ldp = imax(imax(mh * 12, mr * 4), 4);
  int o = 0;
  g.oWg = o;   o += mh * 12 * g.wgst;
  g.oWr1 = o;  o += mr * 4 * g.sR;
  g.oW2 = o;   o += a.D * g.w2st;
  g.oXh = o;   o += MAXB * g.sR;
  g.oXx = o;   o += MAXB * imax(g.sDx, g.sDr);
  g.oPart = o; o += imax(SCAN_NW * MAXB * g.ldp, SCAN_NW * MAXRPU * 32);
  g.oAcc = o;  o += MAXB * g.ldp;
  g.oX0 = o;   o += g.sDx;
  g.oMisc = o; o += 8 * MAXB + 64;
  g.oInt = o;  o += r4(64 + 64 + SCAN_G);      // z0 class indices, z_{t-1} indices of the owned row, column counts
  g.total = o;
  return g;
}

__global__ void __launch_bounds__(SCAN_NT, 1) rssm_scan_fwd_kernel(const b200rl_rssm_scan_args a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const int T = a.T, B = a.B, S = a.S, D = a.D, Z = S * D, R = a.R, A = a.A, Dx = a.Dx, Dr = a.Dr;
  const int NB = T * B;
  const GeoF g = make_geo_f(a, cta);
  const Workspace ws = carve(a.workspace, T, B, S, D, Dx, R, Dr);
  const LLGeo L = make_ll(S, Dx, R, Dr, Z);
  float* Wg = sm + g.oWg;       // [ngh*12][wgst]  rows: (group, part r/c/u, col-in-group); cols [h (sR) | x (sDx)]
  float* Wr1 = sm + g.oWr1;     // [ngr*4][sR]
  float* W2 = sm + g.oW2;       // [D][w2st] rows of W_r2 of this CTA's categorical group
  float* Xh = sm + g.oXh;       // [MAXB][sR]  h rows (carried from one step to the next)
  float* Xx = sm + g.oXx;       // [MAXB][sDx] x rows; reused for the unit's rp rows
  float* PART = sm + g.oPart;
  float* ACC = sm + g.oAcc;     // [MAXB][ldp] h-part of the GRU product, then the finished g_pre columns
  float* X0 = sm + g.oX0;       // [Dx] x_pre contribution of the learned initial posterior z0
  float* misc = sm + g.oMisc;   // [0,16) first flags; [16,32) mean; [32,48) rstd; [48,80) reduction scratch
  int* z0idx = (int*)(sm + g.oInt);
  int* zrow = z0idx + 64;
  int* nctab = zrow + 64;
  const int ldp = g.ldp, xxs = imax(g.sDx, g.sDr);
  const bool owner = g.owner_row >= 0, sampler = g.unit_g >= 0;
  const bool prof_on = (cta < 2) && tid == 0;
  long long* prof = ws.prof + cta * 32;
  long long tlast = prof_on ? clock64() : 0;
  Spin sp;
  sp.init(ws.error);

  // ---------------- prologue: weight slices -> shared memory (read from HBM once per scan)
  for (int e = tid; e < g.total; e += SCAN_NT) sm[e] = 0.f;
  __syncthreads();
  for (int gi = 0; gi < g.ngh; ++gi)
    for (int part = 0; part < 3; ++part) {
      float* dst = Wg + (size_t)(gi * 3 + part) * 4 * g.wgst;
      const float* Wp = a.W_g + (size_t)part * R * (R + Dx);
      load_rows4(dst, g.wgst, 0, Wp, R + Dx, (cta + gi * SCAN_G) * 4, R, 0, R, g.sR, tid);
      load_rows4(dst, g.wgst, g.sR, Wp, R + Dx, (cta + gi * SCAN_G) * 4, R, R, Dx, g.sDx, tid);
    }
  for (int gi = 0; gi < g.ngr; ++gi)
    load_rows4(Wr1 + (size_t)gi * 4 * g.sR, g.sR, 0, a.W_r1, a.ld_wr1, (cta + gi * SCAN_G) * 4, Dr, 0, R, g.sR, tid);
  if (sampler)
    for (int e = tid; e < D * Dr; e += SCAN_NT) {
      const int d = e / Dr, k = e - d * Dr;
      W2[d * g.w2st + k] = a.W_r2[((size_t)g.unit_g * D + d) * Dr + k];
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
  if (owner)       // x_pre row of z0 (used wherever is_first is set): sum of the gathered rows of W_in^T
    for (int c = tid; c < Dx; c += SCAN_NT) {
      float acc = 0.f;
      for (int gq = 0; gq < S; ++gq) acc += a.W_in_t[(size_t)(gq * D + z0idx[gq]) * Dx + c];
      X0[c] = acc;
    }
  // z_in of step 0 (z_{-1} = 0): f * z0, written by the sampling units for their (rows, group) block
  if (sampler)
    for (int e = tid; e < g.unit_nr * D; e += SCAN_NT) {
      const int b = g.unit_r0 + e / D, d = e % D;
      a.z_in[(size_t)b * Z + g.unit_g * D + d] = a.first[b] * ((z0idx[g.unit_g] == d) ? 1.f : 0.f);
    }
  __syncthreads();
  prof_mark(prof, 0, tlast, prof_on);

  for (int t = 0; t < T; ++t) {
    const size_t row0 = (size_t)t * B;
    const int par = t & 1;
    const unsigned tag = (unsigned)t + 1u;
    if (tid < MAXB) misc[tid] = (tid < B) ? a.first[row0 + tid] : 0.f;
    // prefetches that do not depend on the chain
    float pe_pref[4] = {0.f, 0.f, 0.f, 0.f};      // pe[t][b][own rp columns]: ngr*4*B values over 256 threads
    {
      const int nv = B * g.ngr * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + u * SCAN_NT;
        if (e < nv) {
          const int b = e / (g.ngr * 4), cj = e - b * (g.ngr * 4);
          const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
          if (col < Dr) pe_pref[u] = a.pe[(row0 + b) * Dr + col];
        }
      }
    }
    float noise_pref = 1.f;
    if (sampler && wid < g.unit_nr && lane < D)
      noise_pref = a.noise[(row0 + g.unit_r0 + wid) * Z + (size_t)g.unit_g * D + lane];
    __syncthreads();
    const float* fl = misc;

    // ============ A (row owner): x = SiLU(LN(W_in [z_in, a_in])) for the owned row; z_in one-hot -> row gather
    if (owner) {
      const int b = g.owner_row;
      const float f = fl[b];
      if (t > 0 && tid < S) zrow[tid] = (int)__float_as_uint(ll_wait(ws.ll + L.z + ((size_t)(par ^ 1) * MAXB + b) * S + tid, (unsigned)t, sp));
      __syncthreads();
      prof_mark(prof, 2, tlast, prof_on);
      float xv[4];
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = tid + u * SCAN_NT;
        xv[u] = 0.f;
        if (c < Dx) {
          float acc = 0.f;
          if (t > 0 && f != 1.f)
            for (int gq = 0; gq < S; ++gq) acc += a.W_in_t[(size_t)(gq * D + zrow[gq]) * Dx + c];
          float aa = 0.f;
          for (int qq = 0; qq < A; ++qq) aa = fmaf(a.actions[(row0 + b) * A + qq], a.W_in_t[(size_t)(Z + qq) * Dx + c], aa);
          xv[u] = (1.f - f) * (acc + aa) + f * X0[c];
          a.x_pre[(row0 + b) * Dx + c] = xv[u];
          s += xv[u];
        }
      }
      const float mu = block_sum(s, misc + 48) / (float)Dx;
      float v = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = tid + u * SCAN_NT;
        if (c < Dx) { const float d = xv[u] - mu; v = fmaf(d, d, v); }
      }
      const float rstd = rsqrtf(block_sum(v, misc + 48) / (float)Dx + a.eps);
      u64* dst = ws.ll + L.x + ((size_t)par * MAXB + b) * Dx;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = tid + u * SCAN_NT;
        if (c < Dx) {
          const float o = fsilu((xv[u] - mu) * rstd * a.lnx_g[c] + a.lnx_b[c]);
          ll_store(dst + c, o, tag);
          a.x_act[(row0 + b) * Dx + c] = o;
        }
      }
      if (tid == 0) { ws.ln_stats[((size_t)0 * NB + row0 + b) * 2] = mu; ws.ln_stats[((size_t)0 * NB + row0 + b) * 2 + 1] = rstd; }
      for (int e = tid; e < A; e += SCAN_NT) a.a_in[(row0 + b) * A + e] = (1.f - f) * a.actions[(row0 + b) * A + e];
      prof_mark(prof, 3, tlast, prof_on);
    }

    // ============ B1: h_in = (1-f) h_{t-1} + f h0 (agent.py:428) in place; h-part of the GRU product
    for (int e = tid; e < B * R; e += SCAN_NT) {
      const int b = e / R, k = e - b * R;
      const float f = fl[b];
      const float hp = (t > 0) ? Xh[b * g.sR + k] : 0.f;
      Xh[b * g.sR + k] = (1.f - f) * hp + f * a.h0[k];
    }
    __syncthreads();
    for (int e = tid; e < B * g.ngh * 4; e += SCAN_NT) {     // save h_in for the owned columns
      const int b = e / (g.ngh * 4), cj = e - b * (g.ngh * 4);
      const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
      if (col < R) a.h_in[(row0 + b) * R + col] = Xh[b * g.sR + col];
    }
    const int ksh = product(Xh, g.sR, Wg, g.wgst, g.ngh * 3, R, PART, ldp, 0, tid);
    __syncthreads();
    for (int e = tid; e < B * g.ngh * 12; e += SCAN_NT) {
      const int b = e / (g.ngh * 12), c = e - b * (g.ngh * 12);
      ACC[b * ldp + c] = part_sum(PART, ldp, ksh, b, c);
    }
    __syncthreads();
    prof_mark(prof, 1, tlast, prof_on);

    // ============ B2: x-part of the GRU product; g_pre columns; partial LayerNorm statistics
    ll_recv(Xx, xxs, ws.ll + L.x + (size_t)par * MAXB * Dx, Dx, B, Dx, tag, tid, sp);
    __syncthreads();
    prof_mark(prof, 4, tlast, prof_on);
    const int ksx = product(Xx, xxs, Wg + g.sR, g.wgst, g.ngh * 3, Dx, PART, ldp, 0, tid);
    __syncthreads();
    for (int e = tid; e < B * g.ngh * 12; e += SCAN_NT) {
      const int b = e / (g.ngh * 12), c = e - b * (g.ngh * 12);
      const float v = ACC[b * ldp + c] + part_sum(PART, ldp, ksx, b, c);
      ACC[b * ldp + c] = v;
      const int gi = c / 12, part = (c % 12) / 4, j = c & 3;
      const int col = (cta + gi * SCAN_G) * 4 + j;
      if (col < R) a.g_pre[(row0 + b) * 3 * R + part * R + col] = v;
    }
    __syncthreads();
    for (int b = wid; b < B; b += SCAN_NW) {   // per-row partial statistics (mean, M2) over the owned valid columns
      float s = 0.f;
      int cnt = 0;
      for (int c = lane; c < g.ngh * 12; c += 32) {
        const int col = (cta + (c / 12) * SCAN_G) * 4 + (c & 3);
        if (col < R) { s += ACC[b * ldp + c]; ++cnt; }
      }
      s = warp_sum(s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      const float mean = cnt > 0 ? s / (float)cnt : 0.f;
      float m2 = 0.f;
      for (int c = lane; c < g.ngh * 12; c += 32) {
        const int col = (cta + (c / 12) * SCAN_G) * 4 + (c & 3);
        if (col < R) { const float d = ACC[b * ldp + c] - mean; m2 = fmaf(d, d, m2); }
      }
      m2 = warp_sum(m2);
      if (lane == 0) ll_store2(ws.ll + L.s + (((size_t)par * MAXB + b) * SCAN_G + cta) * 2, mean, m2, tag);
    }
    prof_mark(prof, 5, tlast, prof_on);

    // ============ C: merge statistics (Chan), LayerNorm, GRU gate -> h_t for the owned columns
    for (int b = wid; b < B; b += SCAN_NW) {
      float pm[SCAN_G / 32], pq[SCAN_G / 32];
#pragma unroll
      for (int i = 0; i < SCAN_G / 32; ++i) {
        const u64* p = ws.ll + L.s + (((size_t)par * MAXB + b) * SCAN_G + lane + 32 * i) * 2;
        u64 x, y;
        ll_load2(p, x, y);
        while ((unsigned)(x >> 32) != tag || (unsigned)(y >> 32) != tag) {
          if (sp.fail()) break;
          ll_load2(p, x, y);
        }
        pm[i] = __uint_as_float((unsigned)x);
        pq[i] = __uint_as_float((unsigned)y);
      }
      float sm_ = 0.f;
#pragma unroll
      for (int i = 0; i < SCAN_G / 32; ++i) sm_ += (float)(3 * nctab[lane + 32 * i]) * pm[i];
      const float mean = warp_sum(sm_) / (float)(3 * R);
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < SCAN_G / 32; ++i) {
        const float d = pm[i] - mean;
        m2 += pq[i] + (float)(3 * nctab[lane + 32 * i]) * d * d;
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
    prof_mark(prof, 6, tlast, prof_on);
    for (int e = tid; e < B * g.ngh * 4; e += SCAN_NT) {
      const int b = e / (g.ngh * 4), cj = e - b * (g.ngh * 4);
      const int gi = cj >> 2, j = cj & 3;
      const int col = (cta + gi * SCAN_G) * 4 + j;
      if (col >= R) continue;
      const float mu = misc[16 + b], rstd = misc[32 + b];
      float gl[3];
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const float v = ACC[b * ldp + gi * 12 + part * 4 + j];
        gl[part] = (v - mu) * rstd * a.lng_g[part * R + col] + a.lng_b[part * R + col];
        a.g_ln[(row0 + b) * 3 * R + part * R + col] = gl[part];
      }
      const float r = sigmoidf_(gl[0]);
      const float c = tanhf(r * gl[1]);
      const float u = sigmoidf_(gl[2] - 1.f);
      const float h = u * c + (1.f - u) * Xh[b * g.sR + col];
      ll_store(ws.ll + L.h + ((size_t)par * MAXB + b) * R + col, h, tag);
      a.latent[(row0 + b) * a.ld_lat + Z + col] = h;
    }
    prof_mark(prof, 7, tlast, prof_on);

    // ============ D: rp_pre = h W_r1[:, :R]^T + pe for the owned columns (h rows stay in Xh for the next step)
    __syncthreads();                         // every thread is done reading the old Xh
    ll_recv(Xh, g.sR, ws.ll + L.h + (size_t)par * MAXB * R, R, B, R, tag, tid, sp);
    __syncthreads();
    prof_mark(prof, 8, tlast, prof_on);
    const int ksr = product(Xh, g.sR, Wr1, g.sR, g.ngr, R, PART, ldp, 0, tid);
    __syncthreads();
    {
      const int nv = B * g.ngr * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + u * SCAN_NT;
        if (e < nv) {
          const int b = e / (g.ngr * 4), cj = e - b * (g.ngr * 4);
          const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
          if (col < Dr) {
            const float v = part_sum(PART, ldp, ksr, b, cj) + pe_pref[u];
            ll_store(ws.ll + L.r + ((size_t)par * MAXB + b) * Dr + col, v, tag);
            a.rp_pre[(row0 + b) * Dr + col] = v;
          }
        }
      }
      for (int e = tid + 4 * SCAN_NT; e < nv; e += SCAN_NT) {   // (only when a CTA owns more than 16 columns)
        const int b = e / (g.ngr * 4), cj = e - b * (g.ngr * 4);
        const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
        if (col < Dr) {
          const float v = part_sum(PART, ldp, ksr, b, cj) + a.pe[(row0 + b) * Dr + col];
          ll_store(ws.ll + L.r + ((size_t)par * MAXB + b) * Dr + col, v, tag);
          a.rp_pre[(row0 + b) * Dr + col] = v;
        }
      }
    }
    prof_mark(prof, 9, tlast, prof_on);

    // ============ E (sampling unit): LN + SiLU of the unit's rows, logits of its group, unimix, sample
    if (sampler) {
      const int gq = g.unit_g, nr = g.unit_nr, rb = g.unit_r0;
      __syncthreads();                       // PART / Xx free
      ll_recv(Xx, xxs, ws.ll + L.r + ((size_t)par * MAXB + rb) * Dr, Dr, nr, Dr, tag, tid, sp);
      __syncthreads();
      prof_mark(prof, 10, tlast, prof_on);
      for (int bb = wid; bb < nr; bb += SCAN_NW) {
        float* xr = Xx + bb * xxs;
        float s = 0.f;
        for (int k = lane; k < Dr; k += 32) s += xr[k];
        const float mu = warp_sum(s) / (float)Dr;
        float v = 0.f;
        for (int k = lane; k < Dr; k += 32) { const float d = xr[k] - mu; v = fmaf(d, d, v); }
        const float rstd = rsqrtf(warp_sum(v) / (float)Dr + a.eps);
        for (int k = lane; k < Dr; k += 32) {
          const float o = fsilu((xr[k] - mu) * rstd * a.lnr_g[k] + a.lnr_b[k]);
          xr[k] = o;
          if (gq == 0) a.rp_act[(row0 + rb + bb) * Dr + k] = o;
        }
        if (gq == 0 && lane == 0) {
          ws.ln_stats[((size_t)2 * NB + row0 + rb + bb) * 2] = mu;
          ws.ln_stats[((size_t)2 * NB + row0 + rb + bb) * 2 + 1] = rstd;
        }
      }
      __syncthreads();
      // logits: lane = class, warp = K slice, all rows of the unit
      {
        const int kc = ((g.sDr + SCAN_NW - 1) / SCAN_NW + 3) / 4 * 4;
        const int k0 = wid * kc, k1 = imin(g.sDr, k0 + kc);
        float acc[MAXRPU];
#pragma unroll
        for (int i = 0; i < MAXRPU; ++i) acc[i] = 0.f;
        const float* wrow = W2 + (size_t)imin(lane, D - 1) * g.w2st;
        for (int k = k0; k < k1; k += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
          for (int i = 0; i < MAXRPU; ++i)
            if (i < nr) {
              const float4 x = *reinterpret_cast<const float4*>(Xx + i * xxs + k);
              acc[i] = fmaf(x.w, w4.w, fmaf(x.z, w4.z, fmaf(x.y, w4.y, fmaf(x.x, w4.x, acc[i]))));
            }
        }
#pragma unroll
        for (int i = 0; i < MAXRPU; ++i)
          if (i < nr) PART[((size_t)wid * MAXRPU + i) * 32 + lane] = acc[i];
      }
      __syncthreads();
      // one warp per row: the D classes of the group live on the lanes (D <= 32)
      for (int bb = wid; bb < nr; bb += SCAN_NW) {
        const int b = rb + bb;
        const bool on = lane < D;
        float lg = 0.f;
        for (int s2 = 0; s2 < SCAN_NW; ++s2) lg += PART[((size_t)s2 * MAXRPU + bb) * 32 + lane];
        const float raw = on ? lg + a.b_r2[gq * D + lane] : -INFINITY;
        const size_t o = (row0 + b) * Z + (size_t)gq * D + lane;
        const float mx = warp_max(raw);
        const float ex = on ? expf(raw - mx) : 0.f;
        const float se = warp_sum(ex);
        float l = raw;
        if (a.unimix > 0.f && on) {
          const float pmx = (1.f - a.unimix) * (ex / se) + a.unimix / (float)D;
          l = logf(fminf(fmaxf(pmx, kFp32Eps), 1.f - kFp32Eps));
        }
        if (on) {
          a.post_raw[o] = raw;
          a.post_mix[o] = l;
        }
        // torch Categorical: lg = l - logsumexp(l); probs = softmax(lg); sample = argmax(probs / q)
        const float lmx = warp_max(on ? l : -INFINITY);
        const float lse = lmx + logf(warp_sum(on ? expf(l - lmx) : 0.f));
        const float lgmax = warp_max(on ? l - lse : -INFINITY);
        const float pe_ = on ? expf(l - lse - lgmax) : 0.f;
        const float psum = warp_sum(pe_);
        const float q = (bb == wid) ? noise_pref : (on ? a.noise[o] : 1.f);
        float best = on ? (pe_ / psum) / q : -INFINITY;
        int besti = on ? lane : 0x7fffffff;
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, s2);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, s2);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) {
          ll_store(ws.ll + L.z + ((size_t)par * MAXB + b) * S + gq, __uint_as_float((unsigned)besti), tag);
          ws.zidx[(row0 + b) * S + gq] = besti;
        }
        if (on) {
          a.latent[(row0 + b) * a.ld_lat + (size_t)gq * D + lane] = (lane == besti) ? 1.f : 0.f;
          if (t + 1 < T) {   // z_in of the next step: (1-f) z_t + f z0 (agent.py:430), both one-hot
            const float fn = a.first[row0 + B + b];
            a.z_in[(row0 + B + b) * Z + (size_t)gq * D + lane] =
                (1.f - fn) * ((lane == besti) ? 1.f : 0.f) + fn * ((z0idx[gq] == lane) ? 1.f : 0.f);
          }
        }
      }
      prof_mark(prof, 11, tlast, prof_on);
    }
    __syncthreads();
    if (sp.dead) return;   // a hand-off timed out somewhere: bail out, never hang
  }
}

// =====================================================================================================
// backward (BPTT); consumes the activations, LayerNorm statistics and class indices saved by the forward kernel.
//
// Per step (t = T-1 .. 0) the gradient walks  z_t -> post logits -> rp -> h_t -> gates -> (h_{t-1}, x_t) -> z_{t-1}.
// LayerNorm backward  d_pre = rstd (dxh - mean(dxh) - xh mean(dxh xh))  is NOT a hand-off of its own: the producer
// sends dxh = d_act * act'(ln) * gamma for its columns together with its partial row sums, and the CONSUMER of d_pre
// (always a product d_pre W) applies the correction through linearity:
//     d_pre W = rstd (dxh W  -  S1 colsum(W)  -  S2 (xh W)),      xh W = rstd (pre W - mu colsum(W))
// where `pre W` does not depend on the backward chain: it is one batched tensor-core product per LayerNorm over all
// T*B rows, done before this kernel (q.q_r / q.q_g / q.q_x).  The d_pre rows themselves (operands of the deferred
// weight-gradient products) are produced afterwards by the batched LayerNorm-backward kernels from the d_act rows
// saved here.  That leaves four hand-offs per step: d_post_raw rows, dxh of the representation LayerNorm, dxh of the
// GRU LayerNorm, dxh of the x LayerNorm; the last one is consumed by the CTA that owns the (rows, categorical group)
// block, which immediately turns it into the next step's d_post_raw (straight-through + unimix + softmax backward).
// =====================================================================================================
struct GeoB {
  int ngh, ngx, ngr;            // owned 4-column groups of R (dh), Dx (d_x_act), Dr (d_rp_act)
  int sZ, sR, sDx, sDr, xw;     // padded widths; width of the row staging buffer
  int wgst, winst;              // smem row strides of the W_g^T slice (3 parts of sR) and the unit's W_in slice
  int nsplit, rpu, unit_g, unit_r0, unit_nr;
  int ldp;
  int oW2, oW1, oWg, oWin, oX, oPart, oAcc, oDhin, oDhc, oDh0, oWs, oPre, oMisc, total;
};

// per-step prefetch area (floats): small chain-independent inputs of the owned columns
struct PreB {
  int rp, gl, gp, hin, xp, dlh, qr, qg, n;
};
__host__ __device__ inline PreB make_pre(int mh, int mx, int mr) {
  PreB p;
  int o = 0;
  p.rp = o;  o += MAXB * mr * 4;        // rp_pre
  p.gl = o;  o += MAXB * mh * 12;       // g_ln
  p.gp = o;  o += MAXB * mh * 12;       // g_pre
  p.hin = o; o += MAXB * mh * 4;        // h_in
  p.xp = o;  o += MAXB * mx * 4;        // x_pre
  p.dlh = o; o += MAXB * mh * 4;        // d_latent (h part)
  p.qr = o;  o += MAXB * mh * 4;        // q_r
  p.qg = o;  o += MAXB * (mh + mx) * 4; // q_g
  p.n = o;
  return p;
}

__host__ __device__ inline GeoB make_geo_b(const b200rl_rssm_scan_args& a, int cta) {
  GeoB g;
  const int Z = a.S * a.D;
  g.ngh = owned_groups(a.R, cta);
  g.ngx = owned_groups(a.Dx, cta);
  g.ngr = owned_groups(a.Dr, cta);
  g.sZ = r4(Z); g.sR = r4(a.R); g.sDx = r4(a.Dx); g.sDr = r4(a.Dr);
  g.xw = imax(imax(g.sZ, g.sR), imax(g.sDx, g.sDr));
  g.wgst = 3 * g.sR;
  g.winst = g.sDx + 4;
  g.nsplit = imax(1, imin(a.B, SCAN_G / a.S));
  g.rpu = (a.B + g.nsplit - 1) / g.nsplit;
  g.unit_g = -1; g.unit_r0 = 0; g.unit_nr = 0;
  if (cta < a.S * g.nsplit) {
    const int sp = cta / a.S;
    g.unit_r0 = sp * g.rpu;
    g.unit_nr = imin(g.rpu, a.B - g.unit_r0);
    if (g.unit_nr > 0) g.unit_g = cta % a.S; else g.unit_nr = 0;
  }
  const int mh = owned_groups(a.R, 0), mx = owned_groups(a.Dx, 0), mr = owned_groups(a.Dr, 0);
  g.ldp = imax(imax((mh + mx) * 4, mr * 4), 4);
  int o = 0;
  g.oW2 = o;   o += mr * 4 * g.sZ;             // W_r2[:, c]  (own d_rp_act columns)
  g.oW1 = o;   o += mh * 4 * g.sDr;            // W_r1[:, j]  (own dh columns)
  g.oWg = o;   o += (mh + mx) * 4 * g.wgst;    // W_g[:, j] (h part) then W_g[:, R + c] (x part), 3 parts of sR each
  g.oWin = o;  o += a.D * g.winst;             // W_in[:, gq*D + d] of the unit's group
  g.oX = o;    o += MAXB * g.xw;
  g.oPart = o; o += imax(SCAN_NW * MAXB * g.ldp, SCAN_NW * MAXRPU * 32);
  g.oAcc = o;  o += MAXB * g.ldp;
  g.oDhin = o; o += MAXB * mh * 4;
  g.oDhc = o;  o += MAXB * mh * 4;
  g.oDh0 = o;  o += mh * 4;
  g.oWs = o;   o += mh * 4 + (mh + mx) * 4 + 32;   // column sums of the weight slices
  g.oPre = o;  o += make_pre(mh, mx, mr).n;
  g.oMisc = o; o += 16 * MAXB + 64;
  g.total = o;
  return g;
}

// (sum dxh, sum dxh*xh) / n of the rows [r0, r0+nr) from the per-CTA partials; one warp per row, fixed summation order
__device__ __forceinline__ void recv_row_sums(const u64* base, int par, int r0, int nr, unsigned tag, float inv, float* out1,
                                              float* out2, int tid, Spin& sp) {
  const int lane = tid & 31, wid = tid >> 5;
  for (int bb = wid; bb < nr; bb += SCAN_NW) {
    const int b = r0 + bb;
    float v0 = 0.f, v1 = 0.f;
    u64 x[SCAN_G / 32], y[SCAN_G / 32];
#pragma unroll
    for (int i = 0; i < SCAN_G / 32; ++i)
      ll_load2(base + (((size_t)par * MAXB + b) * SCAN_G + lane + 32 * i) * 2, x[i], y[i]);
#pragma unroll
    for (int i = 0; i < SCAN_G / 32; ++i) {
      const u64* p = base + (((size_t)par * MAXB + b) * SCAN_G + lane + 32 * i) * 2;
      while ((unsigned)(x[i] >> 32) != tag || (unsigned)(y[i] >> 32) != tag) {
        if (sp.fail()) break;
        ll_load2(p, x[i], y[i]);
      }
      v0 += __uint_as_float((unsigned)x[i]);
      v1 += __uint_as_float((unsigned)y[i]);
    }
    v0 = warp_sum(v0); v1 = warp_sum(v1);
    if (lane == 0) { out1[b] = v0 * inv; out2[b] = v1 * inv; }
  }
}

__global__ void __launch_bounds__(SCAN_NT, 1)
rssm_scan_bwd_kernel(const b200rl_rssm_scan_args a, const b200rl_rssm_scan_grads q) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const int T = a.T, B = a.B, S = a.S, D = a.D, Z = S * D, R = a.R, Dx = a.Dx, Dr = a.Dr;
  const int KG = R + Dx, KIN = Z + a.A, NB = T * B;
  const GeoB g = make_geo_b(a, cta);
  const Workspace ws = carve(a.workspace, T, B, S, D, Dx, R, Dr);
  const LLGeo L = make_ll(S, Dx, R, Dr, Z);
  const int mh = owned_groups(R, 0), mx = owned_groups(Dx, 0), mr = owned_groups(Dr, 0);
  const PreB P = make_pre(mh, mx, mr);
  float* W2T = sm + g.oW2;      // [ngr*4][sZ]
  float* W1T = sm + g.oW1;      // [ngh*4][sDr]
  float* WgT = sm + g.oWg;      // [(ngh+ngx)*4][3*sR]
  float* WinU = sm + g.oWin;    // [D][winst]
  float* X = sm + g.oX;         // [MAXB][xw]
  float* PART = sm + g.oPart;
  float* ACC = sm + g.oAcc;     // [MAXB][ldp]
  float* DHIN = sm + g.oDhin;   // [MAXB][nh4] dh wrt h_in through the gate
  float* DHC = sm + g.oDhc;     // [MAXB][nh4] dh carried to step t-1
  float* DH0 = sm + g.oDh0;     // [nh4] accumulated grad of tanh(initial_recurrent_state)
  float* WS1 = sm + g.oWs;      // [nh4] column sums of the W_r1 slice
  float* WSG = WS1 + mh * 4;    // [(ngh+ngx)*4] column sums of the W_g slice
  float* WSI = WSG + (mh + mx) * 4;   // [D] column sums of the unit's W_in slice
  float* PRE = sm + g.oPre;
  float* misc = sm + g.oMisc;   // [0,16) first(t); [16,32) first(t+1); [32,48) S1; [48,64) S2; [64,..) per-row (mu, rstd) x3
  float* S1 = misc + 32;
  float* S2 = misc + 48;
  float* LNS = misc + 64;       // [3][MAXB][2]: x, g, rp LayerNorm statistics of step t
  float* LNX1 = misc + 64 + 6 * MAXB;   // [MAXB][2]: x LayerNorm statistics of step t+1 (units)
  const int xw = g.xw, ldp = g.ldp;
  const int nh4 = g.ngh * 4, nx4 = g.ngx * 4, nr4 = g.ngr * 4;
  const bool unit = g.unit_g >= 0;
  const bool prof_on = (cta < 2) && tid == 0;
  long long* prof = ws.prof + cta * 32;
  long long tlast = prof_on ? clock64() : 0;
  Spin sp;
  sp.init(ws.error);

  // ---------------- prologue: zero everything, then transposed weight slices -> shared memory, column sums
  for (int e = tid; e < g.total; e += SCAN_NT) sm[e] = 0.f;
  __syncthreads();
  for (int gi = 0; gi < g.ngr; ++gi)
    load_cols4(W2T + (size_t)gi * 4 * g.sZ, g.sZ, a.W_r2, Dr, (cta + gi * SCAN_G) * 4, Dr, Z, 0, tid);
  for (int gi = 0; gi < g.ngh; ++gi) {
    const int c0 = (cta + gi * SCAN_G) * 4;
    load_cols4(W1T + (size_t)gi * 4 * g.sDr, g.sDr, a.W_r1, a.ld_wr1, c0, R, Dr, 0, tid);
    for (int part = 0; part < 3; ++part)
      load_cols4(WgT + (size_t)gi * 4 * g.wgst, g.wgst, a.W_g + (size_t)part * R * KG, KG, c0, R, R, part * g.sR, tid);
  }
  for (int gi = 0; gi < g.ngx; ++gi)
    for (int part = 0; part < 3; ++part)
      load_cols4(WgT + (size_t)(g.ngh + gi) * 4 * g.wgst, g.wgst, a.W_g + (size_t)part * R * KG + R, KG,
                 (cta + gi * SCAN_G) * 4, Dx, R, part * g.sR, tid);
  if (unit)
    for (int e = tid; e < D * Dx; e += SCAN_NT) {
      const int c = e / D, d = e - c * D;
      WinU[d * g.winst + c] = a.W_in[(size_t)c * KIN + g.unit_g * D + d];
    }
  __syncthreads();
  for (int j = wid; j < nh4; j += SCAN_NW) {
    float s = 0.f;
    for (int k = lane; k < g.sDr; k += 32) s += W1T[(size_t)j * g.sDr + k];
    s = warp_sum(s);
    if (lane == 0) WS1[j] = s;
  }
  for (int j = wid; j < nh4 + nx4; j += SCAN_NW) {
    float s = 0.f;
    for (int k = lane; k < g.wgst; k += 32) s += WgT[(size_t)j * g.wgst + k];
    s = warp_sum(s);
    if (lane == 0) WSG[j] = s;
  }
  if (unit)
    for (int d = wid; d < D; d += SCAN_NW) {
      float s = 0.f;
      for (int k = lane; k < g.sDx; k += 32) s += WinU[(size_t)d * g.winst + k];
      s = warp_sum(s);
      if (lane == 0) WSI[d] = s;
    }
  __syncthreads();
  prof_mark(prof, 16, tlast, prof_on);

  for (int t = T - 1; t >= 0; --t) {
    const size_t row0 = (size_t)t * B;
    const int bt = T - 1 - t, par = bt & 1;
    const unsigned tag = (unsigned)bt + 1u;
    const bool last = (t == T - 1);

    // ---------------- step-start prefetch of chain-independent inputs (owned columns only)
    if (tid < MAXB) {
      misc[tid] = (tid < B) ? a.first[row0 + tid] : 0.f;
      misc[16 + tid] = (tid < B && !last) ? a.first[row0 + B + tid] : 0.f;
    }
    for (int e = tid; e < 3 * B * 2; e += SCAN_NT) {
      const int which = e / (B * 2), r_ = e - which * (B * 2);
      LNS[which * MAXB * 2 + r_] = ws.ln_stats[((size_t)which * NB + row0) * 2 + r_];
    }
    if (!last)
      for (int e = tid; e < B * 2; e += SCAN_NT) LNX1[e] = ws.ln_stats[((size_t)0 * NB + row0 + B) * 2 + e];
    for (int e = tid; e < B * nr4; e += SCAN_NT) {
      const int b = e / nr4, cj = e - b * nr4;
      const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
      PRE[P.rp + e] = (col < Dr) ? a.rp_pre[(row0 + b) * Dr + col] : 0.f;
    }
    for (int e = tid; e < B * nh4; e += SCAN_NT) {
      const int b = e / nh4, cj = e - b * nh4;
      const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
      const bool ok = col < R;
      PRE[P.hin + e] = ok ? a.h_in[(row0 + b) * R + col] : 0.f;
      PRE[P.dlh + e] = ok ? q.d_latent[(row0 + b) * a.ld_lat + Z + col] : 0.f;
      PRE[P.qr + e] = ok ? q.q_r[(row0 + b) * R + col] : 0.f;
      PRE[P.qg + b * (nh4 + nx4) + cj] = ok ? q.q_g[(row0 + b) * KG + col] : 0.f;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        PRE[P.gl + (b * 3 + part) * nh4 + cj] = ok ? a.g_ln[(row0 + b) * 3 * R + part * R + col] : 0.f;
        PRE[P.gp + (b * 3 + part) * nh4 + cj] = ok ? a.g_pre[(row0 + b) * 3 * R + part * R + col] : 0.f;
      }
    }
    for (int e = tid; e < B * nx4; e += SCAN_NT) {
      const int b = e / nx4, cj = e - b * nx4;
      const int col = (cta + (cj >> 2) * SCAN_G) * 4 + (cj & 3);
      PRE[P.xp + e] = (col < Dx) ? a.x_pre[(row0 + b) * Dx + col] : 0.f;
      PRE[P.qg + b * (nh4 + nx4) + nh4 + cj] = (col < Dx) ? q.q_g[(row0 + b) * KG + R + col] : 0.f;
    }
    __syncthreads();
    const float* fl = misc;

    // ============ P (unit): dz carried from step t+1 through W_in, then the raw-logit gradients of the group
    if (unit) {
      const int gq = g.unit_g, nr = g.unit_nr, rb = g.unit_r0;
      if (!last) {
        ll_recv(X, xw, ws.ll + L.d + ((size_t)(par ^ 1) * MAXB + rb) * Dx, Dx, nr, Dx, (unsigned)bt, tid, sp);
        recv_row_sums(ws.ll + L.sd, par ^ 1, rb, nr, (unsigned)bt, 1.f / (float)Dx, S1, S2, tid, sp);
        __syncthreads();
        prof_mark(prof, 17, tlast, prof_on);
        const int kc = ((g.sDx + SCAN_NW - 1) / SCAN_NW + 3) / 4 * 4;
        const int k0 = wid * kc, k1 = imin(g.sDx, k0 + kc);
        float acc[MAXRPU];
#pragma unroll
        for (int i = 0; i < MAXRPU; ++i) acc[i] = 0.f;
        const float* wrow = WinU + (size_t)imin(lane, D - 1) * g.winst;
        for (int k = k0; k < k1; k += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
          for (int i = 0; i < MAXRPU; ++i)
            if (i < nr) {
              const float4 x = *reinterpret_cast<const float4*>(X + i * xw + k);
              acc[i] = fmaf(x.w, w4.w, fmaf(x.z, w4.z, fmaf(x.y, w4.y, fmaf(x.x, w4.x, acc[i]))));
            }
        }
#pragma unroll
        for (int i = 0; i < MAXRPU; ++i)
          if (i < nr) PART[((size_t)wid * MAXRPU + i) * 32 + lane] = acc[i];
        __syncthreads();
      }
      for (int bb = wid; bb < nr; bb += SCAN_NW) {
        const int b = rb + bb;
        const bool on = lane < D;
        const size_t o = (row0 + b) * Z + (size_t)gq * D + lane;
        float dz = on ? q.d_latent[(row0 + b) * a.ld_lat + (size_t)gq * D + lane] : 0.f;
        if (!last) {
          float p1 = 0.f;
          for (int s2 = 0; s2 < SCAN_NW; ++s2) p1 += PART[((size_t)s2 * MAXRPU + bb) * 32 + lane];
          if (on) {
            const float mu = LNX1[b * 2], rstd = LNX1[b * 2 + 1];
            const float wsum = WSI[lane];
            const float p2 = rstd * (q.q_x[(row0 + B + b) * Z + (size_t)gq * D + lane] - mu * wsum);
            const float dzin = rstd * (p1 - S1[b] * wsum - S2[b] * p2);
            dz += (1.f - fl[16 + b]) * dzin;
          }
        }
        const float raw = on ? a.post_raw[o] : -INFINITY;
        const float mx_ = warp_max(raw);
        const float ex = on ? expf(raw - mx_) : 0.f;
        const float se = warp_sum(ex);
        const float sft = ex / se;
        float pmx = 0.f, l = raw;
        if (a.unimix > 0.f && on) {
          pmx = (1.f - a.unimix) * sft + a.unimix / (float)D;
          l = logf(fminf(fmaxf(pmx, kFp32Eps), 1.f - kFp32Eps));
        }
        float gg = on ? q.d_post_mix[o] : 0.f;
        {
          const float lmx = warp_max(on ? l : -INFINITY);
          const float lse = lmx + logf(warp_sum(on ? expf(l - lmx) : 0.f));
          const float p = on ? expf(l - lse) : 0.f;
          const float pdz = warp_sum(p * dz);
          gg += p * (dz - pdz);
        }
        if (a.unimix > 0.f) {
          const bool inside = on && pmx >= kFp32Eps && pmx <= 1.f - kFp32Eps;
          const float ds = inside ? gg * (1.f - a.unimix) / pmx : 0.f;
          const float sds = warp_sum(sft * ds);
          gg = sft * (ds - sds);
        }
        if (on) {
          q.d_post_raw[o] = gg;
          ll_store(ws.ll + L.a + ((size_t)par * MAXB + b) * Z + (size_t)gq * D + lane, gg, tag);
        }
      }
      prof_mark(prof, 18, tlast, prof_on);
    }

    // ============ Q: d_rp_act = d_post_raw W_r2 for the owned columns; dxh of the representation LayerNorm
    __syncthreads();
    ll_recv(X, xw, ws.ll + L.a + (size_t)par * MAXB * Z, Z, B, Z, tag, tid, sp);
    __syncthreads();
    prof_mark(prof, 19, tlast, prof_on);
    const int ks2 = product(X, xw, W2T, g.sZ, g.ngr, Z, PART, ldp, 0, tid);
    __syncthreads();
    for (int b = wid; b < B; b += SCAN_NW) {
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < nr4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        if (col < Dr) {
          const float dact = part_sum(PART, ldp, ks2, b, c);
          q.d_rp_act[(row0 + b) * Dr + col] = dact;
          const float xh = (PRE[P.rp + b * nr4 + c] - LNS[(2 * MAXB + b) * 2]) * LNS[(2 * MAXB + b) * 2 + 1];
          const float gam = a.lnr_g[col];
          const float ln = xh * gam + a.lnr_b[col];
          const float sg = sigmoidf_(ln);
          const float dxh = dact * sg * (1.f + ln * (1.f - sg)) * gam;
          ll_store(ws.ll + L.b + ((size_t)par * MAXB + b) * Dr + col, dxh, tag);
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) ll_store2(ws.ll + L.sb + (((size_t)par * MAXB + b) * SCAN_G + cta) * 2, s1, s2, tag);
    }
    prof_mark(prof, 20, tlast, prof_on);

    // ============ R: dh = d_latent_h + carry + d_rp_pre W_r1h ; GRU gate backward ; dxh of the GRU LayerNorm
    __syncthreads();
    ll_recv(X, xw, ws.ll + L.b + (size_t)par * MAXB * Dr, Dr, B, Dr, tag, tid, sp);
    recv_row_sums(ws.ll + L.sb, par, 0, B, tag, 1.f / (float)Dr, S1, S2, tid, sp);
    __syncthreads();
    prof_mark(prof, 21, tlast, prof_on);
    const int ks1 = product(X, xw, W1T, g.sDr, g.ngh, Dr, PART, ldp, 0, tid);
    __syncthreads();
    for (int b = wid; b < B; b += SCAN_NW) {
      float s1 = 0.f, s2 = 0.f;
      const float mur = LNS[(2 * MAXB + b) * 2], rstdr = LNS[(2 * MAXB + b) * 2 + 1];
      const float mug = LNS[(1 * MAXB + b) * 2], rstdg = LNS[(1 * MAXB + b) * 2 + 1];
      for (int c = lane; c < nh4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        float dhin = 0.f;
        if (col < R) {
          const float p1 = part_sum(PART, ldp, ks1, b, c);
          const float wsum = WS1[c];
          const float p2 = rstdr * (PRE[P.qr + b * nh4 + c] - mur * wsum);
          float dh = PRE[P.dlh + b * nh4 + c] + rstdr * (p1 - S1[b] * wsum - S2[b] * p2);
          if (!last) dh += DHC[b * nh4 + c];
          const float gr = PRE[P.gl + (b * 3 + 0) * nh4 + c], gc = PRE[P.gl + (b * 3 + 1) * nh4 + c],
                      gu = PRE[P.gl + (b * 3 + 2) * nh4 + c];
          const float r = sigmoidf_(gr), cnd = tanhf(r * gc), u = sigmoidf_(gu - 1.f);
          const float hin = PRE[P.hin + b * nh4 + c];
          const float du = dh * (cnd - hin);
          const float drc = dh * u * (1.f - cnd * cnd);
          float dgl[3];
          dgl[0] = drc * gc * r * (1.f - r);
          dgl[1] = drc * r;
          dgl[2] = du * u * (1.f - u);
          dhin = dh * (1.f - u);
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            q.d_g_ln[(row0 + b) * 3 * R + part * R + col] = dgl[part];
            const float xh = (PRE[P.gp + (b * 3 + part) * nh4 + c] - mug) * rstdg;
            const float dxh = dgl[part] * a.lng_g[part * R + col];
            ll_store(ws.ll + L.c + ((size_t)par * MAXB + b) * 3 * R + (size_t)part * R + col, dxh, tag);
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
          }
        }
        DHIN[b * nh4 + c] = dhin;
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) ll_store2(ws.ll + L.sc + (((size_t)par * MAXB + b) * SCAN_G + cta) * 2, s1, s2, tag);
    }
    prof_mark(prof, 22, tlast, prof_on);

    // ============ S: [dh_in, d_x_act] = d_g_pre W_g for the owned columns (K = 3R in three parts); x-LN dxh
    for (int e = tid; e < MAXB * ldp; e += SCAN_NT) ACC[e] = 0.f;
    for (int part = 0; part < 3; ++part) {
      __syncthreads();
      ll_recv(X, xw, ws.ll + L.c + (size_t)par * MAXB * 3 * R + (size_t)part * R, 3 * R, B, R, tag, tid, sp);
      __syncthreads();
      const int ksg = product(X, xw, WgT + part * g.sR, g.wgst, g.ngh + g.ngx, R, PART, ldp, 0, tid);
      __syncthreads();
      for (int e = tid; e < B * (nh4 + nx4); e += SCAN_NT) {
        const int b = e / (nh4 + nx4), c = e - b * (nh4 + nx4);
        ACC[b * ldp + c] += part_sum(PART, ldp, ksg, b, c);
      }
    }
    recv_row_sums(ws.ll + L.sc, par, 0, B, tag, 1.f / (float)(3 * R), S1, S2, tid, sp);
    __syncthreads();
    prof_mark(prof, 23, tlast, prof_on);
    for (int b = wid; b < B; b += SCAN_NW) {
      const float f = fl[b];
      const float mug = LNS[(1 * MAXB + b) * 2], rstdg = LNS[(1 * MAXB + b) * 2 + 1];
      const float mux = LNS[(0 * MAXB + b) * 2], rstdx = LNS[(0 * MAXB + b) * 2 + 1];
      for (int c = lane; c < nh4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        if (col < R) {
          const float wsum = WSG[c];
          const float p2 = rstdg * (PRE[P.qg + b * (nh4 + nx4) + c] - mug * wsum);
          const float dhin = DHIN[b * nh4 + c] + rstdg * (ACC[b * ldp + c] - S1[b] * wsum - S2[b] * p2);
          DHC[b * nh4 + c] = (1.f - f) * dhin;          // carried to step t-1 (agent.py:428 mask)
          DHIN[b * nh4 + c] = f * dhin;                 // grad of tanh(initial_recurrent_state), summed below
        }
      }
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < nx4; c += 32) {
        const int col = (cta + (c >> 2) * SCAN_G) * 4 + (c & 3);
        if (col < Dx) {
          const float wsum = WSG[nh4 + c];
          const float p2 = rstdg * (PRE[P.qg + b * (nh4 + nx4) + nh4 + c] - mug * wsum);
          const float dact = rstdg * (ACC[b * ldp + nh4 + c] - S1[b] * wsum - S2[b] * p2);
          q.d_x_act[(row0 + b) * Dx + col] = dact;
          const float xh = (PRE[P.xp + b * nx4 + c] - mux) * rstdx;
          const float gam = a.lnx_g[col];
          const float ln = xh * gam + a.lnx_b[col];
          const float sg = sigmoidf_(ln);
          const float dxh = dact * sg * (1.f + ln * (1.f - sg)) * gam;
          ll_store(ws.ll + L.d + ((size_t)par * MAXB + b) * Dx + col, dxh, tag);
          s1 += dxh;
          s2 = fmaf(dxh, xh, s2);
        }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) ll_store2(ws.ll + L.sd + (((size_t)par * MAXB + b) * SCAN_G + cta) * 2, s1, s2, tag);
    }
    __syncthreads();
    for (int c = tid; c < nh4; c += SCAN_NT) {          // fixed row order: bit-reproducible
      float s = DH0[c];
      for (int b = 0; b < B; ++b) s += DHIN[b * nh4 + c];
      DH0[c] = s;
    }
    prof_mark(prof, 24, tlast, prof_on);
    __syncthreads();
    if (sp.dead) break;    // a hand-off timed out somewhere: bail out, never hang
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
  RL_CHECK_ARG(a.Dx % 2 == 0 && a.R % 2 == 0 && a.Dr % 2 == 0 && (a.S * a.D) % 2 == 0, "persistent scan supports even layer widths");
  RL_CHECK_ARG(a.Dx <= 4 * SCAN_NT, "persistent scan supports recurrent dense_units <= 1024");
  RL_CHECK_ARG(a.W_in_t, "W_in_t (transposed recurrent-model input weight) is required");
  RL_CHECK_ARG(a.workspace && a.workspace_bytes >= (long long)ws_bytes(a.T, a.B, a.S, a.D, a.Dx, a.R, a.Dr), "workspace too small");
  return B200RL_OK;
}

}  // namespace

extern "C" long long b200rl_rssm_scan_workspace_bytes(int T, int B, int S, int D, int Dx, int R, int Dr) {
  return (long long)ws_bytes(T, B, S, D, Dx, R, Dr);
}

extern "C" int b200rl_rssm_scan_fwd(const b200rl_rssm_scan_args* args, cudaStream_t st) {
  RL_CHECK_ARG(args, "null args");
  const b200rl_rssm_scan_args& a = *args;
  if (int rc = scan_check(a)) return rc;
  const GeoF g = make_geo_f(a, 0);
  const size_t smem = sizeof(float) * (size_t)g.total;
  RL_CHECK_ARG(smem <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  RL_CUDA(cudaFuncSetAttribute(rssm_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RL_CUDA(cudaMemsetAsync(a.workspace, 0, WS_HEADER + WS_PROF + ws_ll_bytes(a.S, a.D, a.Dx, a.R, a.Dr), st));
  void* kargs[] = {(void*)args};
  RL_CUDA(cudaLaunchCooperativeKernel((void*)rssm_scan_fwd_kernel, dim3(SCAN_G), dim3(SCAN_NT), kargs, smem, st));
  return B200RL_OK;
}

extern "C" int b200rl_rssm_scan_bwd_check(const b200rl_rssm_scan_args* args) {
  RL_CHECK_ARG(args, "null args");
  if (int rc = scan_check(*args)) return rc;
  const GeoB g = make_geo_b(*args, 0);
  RL_CHECK_ARG(sizeof(float) * (size_t)g.total <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  return B200RL_OK;
}

extern "C" int b200rl_rssm_scan_bwd(const b200rl_rssm_scan_args* args, const b200rl_rssm_scan_grads* grads,
                                    cudaStream_t st) {
  RL_CHECK_ARG(args && grads, "null args");
  const b200rl_rssm_scan_args& a = *args;
  if (int rc = scan_check(a)) return rc;
  RL_CHECK_ARG(grads->q_r && grads->q_g && grads->q_x, "q_r / q_g / q_x (pre-activation x weight products) are required");
  const GeoB g = make_geo_b(a, 0);
  const size_t smem = sizeof(float) * (size_t)g.total;
  RL_CHECK_ARG(smem <= 227 * 1024, "weight slices do not fit in shared memory for this model size");
  RL_CUDA(cudaFuncSetAttribute(rssm_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // header + LL region are reset; the forward's saves (class indices, LayerNorm statistics) behind them stay
  RL_CUDA(cudaMemsetAsync(a.workspace, 0, WS_HEADER + WS_PROF + ws_ll_bytes(a.S, a.D, a.Dx, a.R, a.Dr), st));
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

// cycle counters of CTA 0 / CTA 1 (32 slots each) accumulated by the last launch on this workspace
extern "C" int b200rl_rssm_scan_profile(const void* workspace, long long* out64, cudaStream_t st) {
  RL_CUDA(cudaMemcpyAsync(out64, (const char*)workspace + WS_HEADER, WS_PROF, cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));
  return B200RL_OK;
}
