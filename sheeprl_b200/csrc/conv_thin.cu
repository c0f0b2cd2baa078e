// Image layers whose big-image side has only a few channels (the RGB ends of the Dreamer-V3 encoder / decoder):
//   up_thin   : ConvTranspose2d(Cs -> CB, k4 s2 p1) forward  (decoder output layer, agent.py:199-222)
//   wgrad_thin: weight gradient of Conv2d(CB -> Cs) / ConvTranspose2d(Cs -> CB)  (encoder first / decoder last layer)
// With CB = 3 these are not tensor-core shapes (K or N of 3..48) and they move the two largest activations of the
// model (1024 x 32x32x32 fp32 = 134 MB and the 50 MB image), so they are written as FMA/LDS-balanced SIMT kernels
// that read every activation once with full 128-byte lines:
//   bound: max(HBM: 184 MB / launch, FMA: 1.6 GFMA / launch) ~ 50-60 us on a B200; the generic implicit-GEMM kernels
//   they replace took 1.5 ms (up) and 0.7 ms (wgrad) per launch.
// Weights keep the reference layout W[Cs][CB][ky][kx] (see conv.cu header for the index conventions).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// up_thin: one thread per small-grid position (i, j) -> the 2x2 output block (2i+a, 2j+b), all CB channels.
// Output (2i+a) takes small rows i+dy with kernel row ky = a - 2dy + 1: dy=-1 -> a=0,ky=3; dy=0 -> a,ky=a+1;
// dy=+1 -> a=1,ky=0 (same along x).  Per neighbour the (a, b, cb) weights are packed contiguously in shared memory
// (padded to a multiple of 4) so one broadcast LDS.128 feeds 4 FMAs.
// ---------------------------------------------------------------------------------------------------------
template <int D> struct Nb { static constexpr int n = (D == 0) ? 2 : 1; static constexpr int a0 = (D == 1) ? 1 : 0; };

template <int CB> __host__ __device__ constexpr int nb_slots(int p) {   // padded weight count of neighbour p = (dy+1)*3+(dx+1)
  const int ny = (p / 3 == 1) ? 2 : 1, nx = (p % 3 == 1) ? 2 : 1;
  return (ny * nx * CB + 3) / 4 * 4;
}
template <int CB> __host__ __device__ constexpr int nb_base(int p) {
  int s = 0;
  for (int q = 0; q < p; ++q) s += nb_slots<CB>(q);
  return s;
}

template <int CB, int CS, int DY, int DX>
__device__ __forceinline__ void up_neighbour(const float* __restrict__ small, const float* __restrict__ Wn, int i, int j,
                                             int h, int w, long long img_base, float (&acc)[2][2][CB]) {
  const int iy = i + DY, ix = j + DX;
  if (iy < 0 || iy >= h || ix < 0 || ix >= w) return;
  constexpr int P = (DY + 1) * 3 + (DX + 1);
  constexpr int NY = Nb<DY>::n, NX = Nb<DX>::n, A0 = Nb<DY>::a0, B0 = Nb<DX>::a0;
  constexpr int CNT = NY * NX * CB, PADDED = nb_slots<CB>(P), BASE = nb_base<CB>(P), WSTRIDE = nb_base<CB>(9);
  const float4* __restrict__ src = reinterpret_cast<const float4*>(small + (img_base + (long long)iy * w + ix) * CS);
#pragma unroll 2
  for (int c4 = 0; c4 < CS / 4; ++c4) {
    const float4 v = __ldg(src + c4);
    const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4* wp = reinterpret_cast<const float4*>(Wn + (c4 * 4 + q) * WSTRIDE + BASE);
      float wv[PADDED];
#pragma unroll
      for (int e = 0; e < PADDED / 4; ++e) {
        const float4 t = wp[e];
        wv[4 * e] = t.x; wv[4 * e + 1] = t.y; wv[4 * e + 2] = t.z; wv[4 * e + 3] = t.w;
      }
#pragma unroll
      for (int s = 0; s < CNT; ++s) {
        const int cb = s % CB, b = (s / CB) % NX, a = s / (CB * NX);
        acc[A0 + a][B0 + b][cb] = fmaf(vs[q], wv[s], acc[A0 + a][B0 + b][cb]);
      }
    }
  }
}

template <int CB, int CS>
__global__ void __launch_bounds__(128)
conv_up_thin_kernel(const float* __restrict__ small, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ big, int NB, int h, int w) {
  constexpr int WSTRIDE = nb_base<CB>(9);
  __shared__ __align__(16) float Wn[CS * WSTRIDE];
  for (int e = threadIdx.x; e < CS * WSTRIDE; e += blockDim.x) Wn[e] = 0.f;
  __syncthreads();
  // pack: Wn[cs][base(p) + (a*NX + b)*CB + cb] = W[cs][cb][ky][kx]
  for (int e = threadIdx.x; e < CS * CB * 16; e += blockDim.x) {
    const int tap = e % 16, cb = (e / 16) % CB, cs = e / (16 * CB);
    const int ky = tap >> 2, kx = tap & 3;
    // ky = a - 2dy + 1  ->  (ky=3: dy=-1,a=0) (ky=1: dy=0,a=0) (ky=2: dy=0,a=1) (ky=0: dy=1,a=1)
    const int dy = (ky == 3) ? -1 : ((ky == 0) ? 1 : 0), a = (ky == 2 || ky == 0) ? 1 : 0;
    const int dx = (kx == 3) ? -1 : ((kx == 0) ? 1 : 0), b = (kx == 2 || kx == 0) ? 1 : 0;
    const int p = (dy + 1) * 3 + (dx + 1);
    const int nx = (dx == 0) ? 2 : 1;
    const int la = (dy == 0) ? a : 0, lb = (dx == 0) ? b : 0;     // local index inside the neighbour's (a, b) set
    Wn[cs * WSTRIDE + nb_base<CB>(p) + (la * nx + lb) * CB + cb] = W[e];
  }
  __syncthreads();
  const long long total = (long long)NB * h * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % w);
    const long long r = t / w;
    const int i = (int)(r % h);
    const long long n = r / h;
    float acc[2][2][CB];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < CB; ++c) acc[a][b][c] = bias ? bias[c] : 0.f;
    const long long img = n * h * w;
    up_neighbour<CB, CS, -1, -1>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, -1, 0>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, -1, 1>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 0, -1>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 0, 0>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 0, 1>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 1, -1>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 1, 0>(small, Wn, i, j, h, w, img, acc);
    up_neighbour<CB, CS, 1, 1>(small, Wn, i, j, h, w, img, acc);
    const int Wb = 2 * w;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      float* dst = big + ((n * 2 * h + 2 * i + a) * Wb + 2 * j) * (long long)CB;   // 2*CB contiguous floats
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < CB; ++c) dst[b * CB + c] = acc[a][b][c];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// wgrad_thin: dW[cs][cb][ky][kx] += sum over small pixels of small[p][cs] * big[patch(p)][ky][kx][cb].
// CTA tile = up to 32 consecutive small pixels of one image row; the 4 big rows they touch are staged (zero-padded)
// in shared memory, lane = small channel, each warp walks 4 of the pixels: per pixel 1 LDS + 8*CB broadcast LDS.64
// feed 16*CB FMAs on register accumulators; one cross-warp + atomic reduction per CTA at the very end.
// ---------------------------------------------------------------------------------------------------------
template <int CB>
__global__ void __launch_bounds__(256, 2)
conv_wgrad_thin_kernel(const float* __restrict__ small, const float* __restrict__ big, float* __restrict__ dW, int NB,
                       int h, int w, int Cs, int tiles_per_row) {
  constexpr int TX = 32, ROWF = (2 * TX + 2) * CB;              // floats per staged big row (even)
  extern __shared__ __align__(16) float sm[];
  float* Bt = sm;                                                // [4][ROWF]
  float* St = sm + 4 * ROWF;                                     // [TX][Cs]
  float* Red = St + TX * 32;                                     // [16*CB][Cs]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = Cs / 32;
  const int Hb = 2 * h, Wb = 2 * w;
  const long long ntiles = (long long)NB * h * tiles_per_row;
  for (int e = threadIdx.x; e < 16 * CB * Cs; e += blockDim.x) Red[e] = 0.f;
  for (int g = 0; g < groups; ++g) {
    float acc[4][4 * CB];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4 * CB; ++b) acc[a][b] = 0.f;
    // software pipeline: the next tile's global loads are issued into registers before the current tile is consumed
    // (the kernel is otherwise stalled on their latency: 3 CTAs per SM cannot hide it)
    constexpr int NBR = (4 * ROWF + 255) / 256, NSR = TX * 32 / 256;
    float rb[NBR], rs[NSR];
    auto fetch = [&](long long tile) {
      const int tx = (int)(tile % tiles_per_row);
      const long long r = tile / tiles_per_row;
      const int y = (int)(r % h);
      const long long n = r / h;
      const int x0 = tx * TX, npx = min(TX, w - x0);
#pragma unroll
      for (int u = 0; u < NBR; ++u) {
        const int e = threadIdx.x + u * 256;
        float v = 0.f;
        if (e < 4 * ROWF) {
          const int row = e / ROWF, f = e - row * ROWF;
          const int col = f / CB, c = f - col * CB;
          const int yy = 2 * y - 1 + row, xx = 2 * x0 - 1 + col;
          if (yy >= 0 && yy < Hb && xx >= 0 && xx < Wb) v = __ldg(big + ((n * Hb + yy) * Wb + xx) * (long long)CB + c);
        }
        rb[u] = v;
      }
#pragma unroll
      for (int u = 0; u < NSR; ++u) {
        const int e = threadIdx.x + u * 256, p = e >> 5, c = e & 31;
        rs[u] = (p < npx) ? __ldg(small + ((n * h + y) * (long long)w + x0 + p) * Cs + g * 32 + c) : 0.f;
      }
    };
    long long tile = blockIdx.x;
    if (tile < ntiles) fetch(tile);
    for (; tile < ntiles; tile += gridDim.x) {
      __syncthreads();                                         // the previous tile has been consumed
#pragma unroll
      for (int u = 0; u < NBR; ++u) {
        const int e = threadIdx.x + u * 256;
        if (e < 4 * ROWF) Bt[e] = rb[u];
      }
#pragma unroll
      for (int u = 0; u < NSR; ++u) St[threadIdx.x + u * 256] = rs[u];
      __syncthreads();
      if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);  // in flight while this tile is multiplied
#pragma unroll
      for (int q = 0; q < TX / 8; ++q) {
        const int p = warp + 8 * q;
        const float s = St[p * 32 + lane];
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const float2* bp = reinterpret_cast<const float2*>(Bt + ky * ROWF + 2 * p * CB);
#pragma unroll
          for (int e = 0; e < 2 * CB; ++e) {
            const float2 b2 = bp[e];
            acc[ky][2 * e] = fmaf(s, b2.x, acc[ky][2 * e]);
            acc[ky][2 * e + 1] = fmaf(s, b2.y, acc[ky][2 * e + 1]);
          }
        }
      }
    }
    // cross-warp reduction in shared memory, then one atomic per output element and CTA
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int f = 0; f < 4 * CB; ++f) atomicAdd(&Red[(ky * 4 * CB + f) * Cs + g * 32 + lane], acc[ky][f]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 16 * CB * Cs; e += blockDim.x) {
    const int cs = e % Cs, q = e / Cs;                      // q = ky*(4*CB) + kx*CB + cb
    const int ky = q / (4 * CB), rem = q - ky * 4 * CB, kx = rem / CB, cb = rem - kx * CB;
    atomicAdd(&dW[((long long)cs * CB + cb) * 16 + ky * 4 + kx], Red[e]);
  }
}

}  // namespace

// internal entry points used by conv.cu's dispatchers
bool b200rl_thin_up_supported(int Cs, int Cb) { return Cb == 3 && (Cs == 32 || Cs == 48 || Cs == 64 || Cs == 96); }
bool b200rl_thin_wgrad_supported(int Cs, int Cb) { return Cb >= 1 && Cb <= 4 && Cs % 32 == 0 && Cs <= 128; }

int b200rl_conv_up_thin(const float* small, const float* W, float* big, const float* bias, int NB, int h, int w, int Cs,
                        int Cb, cudaStream_t st) {
  const long long total = (long long)NB * h * w;
  long long blocks = (total + 127) / 128;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  switch (Cs) {
    case 32: conv_up_thin_kernel<3, 32><<<(unsigned)blocks, 128, 0, st>>>(small, W, bias, big, NB, h, w); break;
    case 48: conv_up_thin_kernel<3, 48><<<(unsigned)blocks, 128, 0, st>>>(small, W, bias, big, NB, h, w); break;
    case 64: conv_up_thin_kernel<3, 64><<<(unsigned)blocks, 128, 0, st>>>(small, W, bias, big, NB, h, w); break;
    default: conv_up_thin_kernel<3, 96><<<(unsigned)blocks, 128, 0, st>>>(small, W, bias, big, NB, h, w); break;
  }
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

int b200rl_conv_wgrad_thin(const float* small, const float* big, float* dW, int NB, int h, int w, int Cs, int Cb,
                           cudaStream_t st) {
  const int tiles_per_row = (w + 31) / 32;
  const long long ntiles = (long long)NB * h * tiles_per_row;
  // ~13 KB of smem and 80 registers per thread: 6 CTAs per SM hide the stage -> compute latency of a 32-pixel tile
  long long blocks = ntiles < 6LL * kNumSMs ? ntiles : 6LL * kNumSMs;
  const size_t smem = sizeof(float) * (4 * (2 * 32 + 2) * Cb + 32 * 32 + 16 * Cb * Cs);
  switch (Cb) {
    case 1: conv_wgrad_thin_kernel<1><<<(unsigned)blocks, 256, smem, st>>>(small, big, dW, NB, h, w, Cs, tiles_per_row); break;
    case 2: conv_wgrad_thin_kernel<2><<<(unsigned)blocks, 256, smem, st>>>(small, big, dW, NB, h, w, Cs, tiles_per_row); break;
    case 3: conv_wgrad_thin_kernel<3><<<(unsigned)blocks, 256, smem, st>>>(small, big, dW, NB, h, w, Cs, tiles_per_row); break;
    default: conv_wgrad_thin_kernel<4><<<(unsigned)blocks, 256, smem, st>>>(small, big, dW, NB, h, w, Cs, tiles_per_row); break;
  }
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}
