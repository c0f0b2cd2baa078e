// Tensor-core GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T (+bias) (+C), fp32 in / fp32 out, computed as
// 3xTF32 on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged
// in shared memory by TMA with the 128-byte swizzle).
//
// Why 3xTF32: the parity bar for this path is 1e-4 against an fp32 CPU oracle through chains of ~100 dependent
// layers.  Each fp32 operand x is split in-kernel into hi = x with the 13 low mantissa bits
// cleared (exactly what the TF32 datapath keeps) and lo = x - hi (exact in fp32); the product is accumulated
// as hi*hi + hi*lo + lo*hi in fp32 (the dropped lo*lo term is ~2^-22 relative).  Effective rate is 1/3 of the
// TF32 peak, ~10x the FFMA path.
//
// Pipeline (per 128x128 output tile, K step 32 = one 128-byte swizzle atom):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor loads of the raw fp32 A / B tiles, mbarrier complete_tx
//   warps 4-7: splitters      -- write `lo` = x - trunc_tf32(x) of each landed tile to a twin buffer (same swizzled
//                                offsets, so the split is layout-agnostic), fence.proxy.async, arrive.  The raw tile
//                                itself serves as `hi`: the TF32 datapath ignores the 13 low mantissa bits (measured:
//                                identical 1.2e-6 error, 130 -> 151 TF/s from not re-writing hi: the kernel is
//                                shared-memory-bandwidth bound, 160 KB of smem traffic per 128x128x32 k-block)
//   warp 1   : MMA issuer     -- one elected lane issues 4 k-steps x 3 tcgen05.mma per stage, tcgen05.commit
//                                releases the stage back to the producer
//   warps 8-15: accumulators  -- every 4 k-blocks tcgen05.ld the finished TMEM chunk and add it into fp32
//                                registers with RN adds (the tensor core accumulates with truncation), double
//                                buffered TMEM; finally +bias / +C and store
//   warp 2   : TMEM allocator
// Replaces: every large nn.Linear forward / input-gradient product of the Dreamer-V3 step
// (sheeprl/models/models.py MLP; agent.py heads, RSSM imagination, actor, critic).
//
// The same pipeline runs the stride-2 k4 p1 convolutions as IMPLICIT GEMMs (no im2col buffer): the A tile of a
// k-block (one filter tap x 32 input channels) is a 4-D TMA box over the channel-last image
// [N][H][W][C] -- box {32 ch, bw, bh, bn} with element strides {1,2,2,1} for the strided gather of Conv2d
// forward ("down"), {1,1,1,1} for the 2x2 sub-pixel taps of ConvTranspose2d forward ("up") -- and TMA's
// out-of-bounds zero fill supplies the padding.  Replaces CNNEncoder / CNNDecoder convolutions
// (sheeprl/algos/dreamer_v3/agent.py:78-91, :199-222) and their input-gradient passes.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 32, STAGES = 4;   // 4 x 48 KB operand stages; TMEM: 2 accumulators + 4 x (A hi | lo) = 512 columns
// 18 warps: 0 TMA, 1 MMA, 4-7 split A (-> TMEM), 8-15 accumulate, {2, 3, 16, 17} split B (-> shared memory; warp 2 also
// owns the TMEM allocation).  A and B of a stage are split concurrently by different warps: the hi/lo split is a chain
// of dependent latencies (barrier -> LDS -> ALU -> tcgen05.st / STS -> fence), the longest stage of the pipeline for the
// convolutions' narrow N tiles.
constexpr int NTHREADS = 576;
constexpr int SPLIT_WARP0 = 4, NSPLIT_THREADS = 128;
constexpr int ACC_WARP0 = 8, NACC_WARPS = 8;
constexpr int BSPLIT_WARP_HI = 16;                  // B splitters: warps 2, 3, 16, 17

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch failure), never hang the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) asm volatile("trap;");
  }
}
// one lane of a converged warp (warp-uniform choice: the same lane every time)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// A operand from tensor memory (lane = tile row, one 32-bit column per k element), B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile [rows][32 fp32] with 128B swizzle: 8-row groups are 1024 B apart (SBO), one atom along K.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused: single atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// MN-major operand tile (the operand is stored [K][MN] in global memory: transposed products without transposes).
// For 32-bit (tf32) operands the only MN-major shared-memory layout the tensor core accepts is the 128-byte swizzle
// with a 32-byte base (CUTLASS: Layout_MN_SW128_32B_Atom, TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 4 k-rows x
// 128 B (32 MN floats) in which the 32-byte chunk index is XOR-ed with the row index.  A TMA box {32 MN floats, 32
// k-rows} lands as eight such atoms (4 KB); a 128-wide tile is four of those blocks.  Descriptor: LBO = distance
// between 32-wide MN blocks (4096 B), SBO = distance between 4-row k atoms (512 B); one K=8 instruction reads two atoms.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;                  // leading byte offset: next 32-element MN block
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next 4-row k atom
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}

template <int BN>
struct Smem {
  // every operand buffer is a whole number of 1024-byte swizzle groups
  float a_hi[STAGES][BM * BK];   // raw A tiles; their hi / lo halves go to TENSOR MEMORY (see kATmem below)
  float b_hi[STAGES][BN * BK];
  float b_lo[STAGES][BN * BK];
  uint64_t full[STAGES], split[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};

// Two-level accumulation.  The tensor core adds each k-step into the fp32 TMEM accumulator with truncation
// (round-toward-zero), a bias of ~2^-24 |acc| per add that grows linearly with K (measured 1.3e-5 at K=1536).
// So TMEM only ever holds a CHUNK of CH k-blocks (CH*4 k-steps); each finished chunk is drained by the
// accumulator warps into fp32 REGISTERS with round-to-nearest FADDs while the tensor core fills the other
// TMEM buffer.
constexpr int CH = 4;

constexpr int MODE_GEMM = 0, MODE_DOWN = 1, MODE_UP = 2;
// MODE_UP4: ConvTranspose2d forward with all four output parity classes in ONE 128-column tile (Cout == 32): the K loop walks
// the 9 shifted input windows (dy, dx in {-1, 0, 1}) instead of 4 parities x 4 taps, so every input tile is fetched from L2
// 9 times instead of 16 (the per-parity launch ran at the L2 -> SM bandwidth: 1.07 GB per call at 1024 x 16x16x64 inputs),
// the B tile is the parity-major weight [4 * Cout][9 * Cin] with zeros where a (parity, shift) pair does not exist, and the
// epilogue scatters column group p to output pixel (2y + py, 2x + px).
constexpr int MODE_UP4 = 3;
constexpr int GROUP_M = 16;
struct TileGeo {
  int mode;
  int h, w, NB;            // small-image grid (conv modes)
  int bw, bh, bn;          // output tile = bn images x bh rows x bw cols of the small grid (bw*bh*bn == 128)
  int tiles_x, tiles_y;    // tiles per image
  int chunks;              // input channels / 32
  int Cout;                // output channels
  int ksplits;             // GEMM mode: number of K splits (gridDim.z); > 1 => partial tiles go through `part`
  float* part;             // split-K workspace: [ksplits][mpad][ldw] partial sums (fixed-order reduction, no atomics)
  int mpad, ldw;           // split-K workspace geometry (rows per split, row stride)
  int force_part;          // write the partial tile to the workspace even with a single split (fused reduce + LayerNorm tail)
  int passes;              // 3: x = hi + lo split, three TF32 products per k-step (fp32-accurate); 1: one TF32 product
                           // (torch's float32_matmul_precision "high", the reference's default on GPUs)
  int a_mn, b_mn;          // GEMM mode: operand stored [K][M] / [K][N] (MN-major) instead of [M][K] / [N][K]
  int mtiles;              // number of M tiles
  int ntiles;              // number of N tiles.  A CTA walks the flattened (m, n) tile list blockIdx.y, blockIdx.y +
                           // gridDim.y, ... (persistent); consecutive ids cover GROUP_M m-tiles x all n-tiles column by
                           // column, so the ~148 tiles in flight share few operand panels in L2
  int wg;                  // GEMM mode, conv weight gradient: A rows = (tap, big channel), K = small-grid pixels gathered
                           // from the channel-last big image by 4-D TMA boxes of 32 pixels (Cout = big channels)
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ C,
               const float* __restrict__ bias, int M, int N, int K, int ldc, int accumulate, const TileGeo geo) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem<BN>& s = *reinterpret_cast<Smem<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = geo.mtiles * geo.ntiles, tstride = gridDim.y;   // flattened tile count
  auto tile_mn = [&](int id, int& tm, int& tn) {
    if (geo.ntiles == 1) { tm = id; tn = 0; return; }
    const int per = GROUP_M * geo.ntiles, grp = id / per, first = grp * GROUP_M;
    const int gsz = min(GROUP_M, geo.mtiles - first), r = id - grp * per;
    tn = r / gsz;
    tm = first + (r - tn * gsz);
  };
  // split-K (GEMM mode): blockIdx.z owns k-blocks [kb_base, kb_base + nkb); its partial tile goes to the workspace and
  // the LAST split to arrive for an output tile sums all partials in split order (bit-reproducible, no atomics on C)
  int kb_base = 0, nkb = (K + BK - 1) / BK;
  if (geo.mode == MODE_GEMM && geo.ksplits > 1) {
    const int per = (nkb + geo.ksplits - 1) / geo.ksplits;
    kb_base = (int)blockIdx.z * per;
    nkb = min(per, nkb - kb_base);
  }
  const int nchunks = (nkb + CH - 1) / CH;
  // conv modes: a tile's origin on the small-image grid; blockIdx.z = output parity class (up)
  const int py = (int)blockIdx.z >> 1, px = (int)blockIdx.z & 1;
  auto tile_origin = [&](int id, int& tx0, int& ty0, int& tn0) {
    tx0 = (id % geo.tiles_x) * geo.bw;
    ty0 = ((id / geo.tiles_x) % geo.tiles_y) * geo.bh;
    tn0 = (id / (geo.tiles_x * geo.tiles_y)) * geo.bn;
  };
  // Tensor memory: two accumulator buffers (2*BN columns) + per stage the A operand as hi | lo (2 x 32 columns): the
  // A halves are read by the tensor core from TMEM instead of shared memory, which removes half of the operand traffic
  // of a shared-memory-bandwidth-bound kernel (ncu: LSU + tensor-core wavefronts = 84 % of the smem pipe).
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t A_TMEM0 = 2 * BN;     // column of stage 0's A hi; stage st: + 64*st; lo: + 32
  constexpr int ACC_COLS = BN / 2;         // columns per accumulator warp (two warps share a TMEM lane quarter)
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.split[i], (PASSES == 3 ? 2 : 1) * (NSPLIT_THREADS / 32));
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tfull[i], 1);
      mbar_init(&s.tempty[i], NACC_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(&s.tmem_base, TMEM_COLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem_base;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, the elected lane issues: see the MMA issuer)
    {
      const bool leader = elect_one();
      int it = 0;                                  // k-blocks issued so far, across tiles: stage / phase bookkeeping
      for (int tile = blockIdx.y; tile < mtiles; tile += tstride) {
      int tm, tn;
      tile_mn(tile, tm, tn);
      const int m0 = tm * BM, n0 = tn * BN;
      int tx0 = 0, ty0 = 0, tn0 = 0;
      if (geo.mode != MODE_GEMM) tile_origin(tm, tx0, ty0, tn0);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int st = it % STAGES;
        if (it >= STAGES) mbar_wait(&s.empty[st], ((it / STAGES) - 1) & 1);
        if (leader) {
        mbar_expect_tx(&s.full[st], (uint32_t)((BM + BN) * BK * sizeof(float)));
        if (geo.mode == MODE_GEMM) {
          const int k0 = (kb_base + kb) * BK;
          if (geo.wg) {
            // k0 = first of 32 raster-consecutive small pixels (one box bw x bh x bn); block j of the tile = rows of one
            // (tap, 32-channel chunk): the same strided gather as the forward conv, read MN-major
            const int x0 = k0 % geo.w, y0 = (k0 / geo.w) % geo.h, i0 = k0 / (geo.w * geo.h);
            for (int j = 0; j < BM / 32; ++j) {
              const int r0 = m0 + 32 * j, tap = r0 / geo.Cout, ch = r0 - tap * geo.Cout;
              tma_load_4d(s.a_hi[st] + j * 1024, &mapA, &s.full[st], ch, 2 * x0 - 1 + (tap & 3), 2 * y0 - 1 + (tap >> 2), i0);
            }
          } else if (!geo.a_mn) tma_load_2d(s.a_hi[st], &mapA, &s.full[st], k0, m0);
          else
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(s.a_hi[st] + j * 1024, &mapA, &s.full[st], m0 + 32 * j, k0);
          if (!geo.b_mn) tma_load_2d(s.b_hi[st], &mapB, &s.full[st], k0, n0);
          else
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(s.b_hi[st] + j * 1024, &mapB, &s.full[st], n0 + 32 * j, k0);
        } else {
          const int tap = kb / geo.chunks, ch = (kb - tap * geo.chunks) * BK;
          int x, y;
          if (geo.mode == MODE_DOWN)     { x = 2 * tx0 - 1 + (tap & 3); y = 2 * ty0 - 1 + (tap >> 2); }
          else if (geo.mode == MODE_UP4) { x = tx0 + (tap % 3) - 1;     y = ty0 + (tap / 3) - 1; }      // tap = shift index
          else                           { x = tx0 + px - (tap & 1);    y = ty0 + py - (tap >> 1); }
          tma_load_4d(s.a_hi[st], &mapA, &s.full[st], ch, x, y, tn0);
          tma_load_2d(s.b_hi[st], &mapB, &s.full[st], kb * BK, n0 + (geo.mode == MODE_UP ? (int)blockIdx.z * geo.Cout : 0));
        }
        }
        __syncwarp();
      }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The WHOLE warp walks the loop with warp-uniform control flow, so stage / descriptor arithmetic
    // stays on the uniform datapath; only the tcgen05 instructions sit behind the elected lane.  (With the loop inside
    // `if (lane == 0)` ptxas cannot prove uniformity and rebuilds every uniform operand with ELECT + R2UR.BROADCAST: 225
    // dependent SASS instructions per k-block made this single thread the limiter of the kernel, ncu r2_gemm_tc.)
    const bool leader = elect_one();
    const bool b_mn = geo.b_mn != 0;
    // A comes from tensor memory (always [row][k]): only B's major bit depends on the operand layout
    const uint32_t idesc = IDESC | ((uint32_t)b_mn << 16);
    const uint64_t b_step = (b_mn ? 1024 : 32) >> 4;
    // B descriptors of stage 0; a stage advances the 14-bit start-address field (16-byte units) by the buffer size, a
    // k-step by 32 B inside the 128-byte swizzle row (K-major) or by 1024 B = two 4-row k atoms (MN-major)
    const uint64_t dbh_base = b_mn ? make_desc_mn(smem_u32(s.b_hi[0])) : make_desc(smem_u32(s.b_hi[0]));
    const uint64_t dbl_base = b_mn ? make_desc_mn(smem_u32(s.b_lo[0])) : make_desc(smem_u32(s.b_lo[0]));
    constexpr uint64_t STAGE_STEP = (uint64_t)(BN * BK * sizeof(float)) >> 4;
    int it = 0, gc0 = 0;                         // k-blocks / TMEM chunks issued so far, across tiles
    for (int tile = blockIdx.y; tile < mtiles; tile += tstride, gc0 += nchunks)
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int st = it % STAGES;
      const int c = gc0 + kb / CH, buf = c & 1;
      const bool chunk_start = (kb % CH) == 0;
      if (chunk_start && c >= 2)             // the accumulator warps must have drained this TMEM buffer
        mbar_wait(&s.tempty[buf], ((c >> 1) - 1) & 1);
      mbar_wait(&s.split[st], (it / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem + (uint32_t)(buf * BN);
      const uint64_t dbh0 = dbh_base + (uint64_t)st * STAGE_STEP, dbl0 = dbl_base + (uint64_t)st * STAGE_STEP;
      const uint32_t ta_hi = tmem + A_TMEM0 + 64u * (uint32_t)st, ta_lo = ta_hi + 32u;
      if (leader) {
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; ++k4) {
          const uint64_t ob = (uint64_t)k4 * b_step;
          if (PASSES == 3) {
            // small cross terms first, then the leading term
            umma_tf32_ts(acc, ta_hi + 8u * k4, dbl0 + ob, idesc, !(chunk_start && k4 == 0));
            umma_tf32_ts(acc, ta_lo + 8u * k4, dbh0 + ob, idesc, 1);
            umma_tf32_ts(acc, ta_hi + 8u * k4, dbh0 + ob, idesc, 1);
          } else {
            umma_tf32_ts(acc, ta_hi + 8u * k4, dbh0 + ob, idesc, !(chunk_start && k4 == 0));
          }
        }
        umma_commit(&s.empty[st]);   // stage reusable once these MMAs have read it
        if ((kb % CH) == CH - 1 || kb == nkb - 1) umma_commit(&s.tfull[buf]);   // chunk complete
      }
      __syncwarp();
    }
  } else if (warp >= SPLIT_WARP0 && warp < ACC_WARP0) {
    // ===== A splitters: hi/lo decomposition of each landed A tile into tensor memory
    const int t = threadIdx.x - SPLIT_WARP0 * 32;
    const bool a_mn_tile = geo.a_mn != 0;          // GEMM with transposed A, or the conv weight gradient
    int ntl = 0;
    for (int tile = blockIdx.y; tile < mtiles; tile += tstride) ++ntl;
    const int total_kb = ntl * nkb;
    for (int kb = 0; kb < total_kb; ++kb) {
      const int st = kb % STAGES;
      mbar_wait(&s.full[st], (kb / STAGES) & 1);
      // thread t owns tile row t = TMEM lane t (splitter warp w reads/writes lanes [32w, 32w+32)).  It gathers the
      // row's 32 k values from the swizzled tile, stores them (the raw bits are the hi operand: the datapath truncates)
      // and lo = x - trunc(x) into this stage's TMEM columns.
      const char* base = reinterpret_cast<const char*>(s.a_hi[st]);
      uint32_t hi[32], lo[32];
      if (!a_mn_tile) {
        // K-major tile: row t = 128 B, 16-byte chunk c stored at position c ^ (t & 7)
        const char* row = base + t * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ (t & 7)) << 4));
          hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
          hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
        }
      } else {
        // MN-major tile: block t/32 (4 KB), k-row k = 128 B, element t%32 inside it with the 32-byte chunk index
        // XOR-ed by k % 4 (SWIZZLE_128B_BASE32B)
        const char* blk = base + (t >> 5) * 4096 + ((t & 7) << 2);
        const int ch = (t & 31) >> 3;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          hi[k] = *reinterpret_cast<const uint32_t*>(blk + k * 128 + ((ch ^ (k & 3)) << 5));
      }
      const uint32_t ta = tmem + (((uint32_t)(t & ~31)) << 16) + A_TMEM0 + 64u * (uint32_t)st;
      tmem_st32(ta, hi);
      if (PASSES == 3) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float x = __uint_as_float(hi[k]);
          lo[k] = __float_as_uint(x - __uint_as_float(hi[k] & 0xffffe000u));
        }
        tmem_st32(ta + 32u, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // the A halves have landed in TMEM
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // TMEM stores ordered before the arrival
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.split[st]);
    }
  } else if (PASSES == 3 && (warp == 2 || warp == 3 || warp >= BSPLIT_WARP_HI)) {
    // ===== B splitters: lo = x - trunc(x) of each landed B tile to the twin buffer (same swizzled offsets, so the split
    // is layout-agnostic); the raw tile itself is the hi operand
    const int t = (warp >= BSPLIT_WARP_HI ? warp - BSPLIT_WARP_HI + 2 : warp - 2) * 32 + lane;
    int ntl = 0;
    for (int tile = blockIdx.y; tile < mtiles; tile += tstride) ++ntl;
    const int total_kb = ntl * nkb;
    for (int kb = 0; kb < total_kb; ++kb) {
      const int st = kb % STAGES;
      mbar_wait(&s.full[st], (kb / STAGES) & 1);
      const float4* bh = reinterpret_cast<const float4*>(s.b_hi[st]);
      float4* bl = reinterpret_cast<float4*>(s.b_lo[st]);
#pragma unroll
      for (int i = 0; i < BN * BK / 4 / NSPLIT_THREADS; ++i) {
        const int idx = t + i * NSPLIT_THREADS;
        const float4 v = bh[idx];
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        bl[idx] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.split[st]);
    }
  } else if (warp >= ACC_WARP0 && warp < ACC_WARP0 + NACC_WARPS) {
    // ===== accumulators + epilogue.  Warp (q, half): TMEM lanes [32q, 32q+32), columns [half*BN/2, +BN/2).
    const int q = warp & 3, half = (warp - ACC_WARP0) >> 2;
    int gc0 = 0;
    for (int tile = blockIdx.y; tile < mtiles; tile += tstride, gc0 += nchunks) {
    int tm, tn;
    tile_mn(tile, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    int tx0 = 0, ty0 = 0, tn0 = 0;
    if (geo.mode != MODE_GEMM) tile_origin(tm, tx0, ty0, tn0);
    float acc[ACC_COLS];
#pragma unroll
    for (int j = 0; j < ACC_COLS; ++j) acc[j] = 0.f;
    for (int cl = 0; cl < nchunks; ++cl) {
      const int c = gc0 + cl, buf = c & 1;
      mbar_wait(&s.tfull[buf], (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < ACC_COLS; c0 += 16) {   // 16 columns at a time: 96 registers per thread (576 threads)
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * ACC_COLS + c0), r);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tempty[buf]);
    }
    const int row = m0 + q * 32 + lane;
    bool row_ok = row < M;
    size_t row_off = (size_t)row * ldc;
    if (geo.mode != MODE_GEMM) {
      const int r = q * 32 + lane;
      const int wi = r % geo.bw, hi = (r / geo.bw) % geo.bh, ni = r / (geo.bw * geo.bh);
      const int n = tn0 + ni, y = ty0 + hi, x = tx0 + wi;
      row_ok = n < geo.NB;
      if (geo.mode == MODE_DOWN) row_off = (((size_t)n * geo.h + y) * geo.w + x) * (size_t)ldc;
      else row_off = (((size_t)n * (2 * geo.h) + (2 * y + py)) * (2 * geo.w) + (2 * x + px)) * (size_t)ldc;
    }
    bool store = true;
    if (geo.mode == MODE_UP4) {
      // columns [32p, 32p + 32) of the tile = the Cout = 32 channels of output pixel (2y + (p >> 1), 2x + (p & 1))
      if (row_ok) {
        const int r = q * 32 + lane;
        const int wi = r % geo.bw, hi = (r / geo.bw) % geo.bh, ni = r / (geo.bw * geo.bh);
        const size_t n = (size_t)(tn0 + ni);
        const int y = ty0 + hi, x = tx0 + wi;
#pragma unroll
        for (int jp = 0; jp < ACC_COLS / 32; ++jp) {
          const int p = (half * ACC_COLS) / 32 + jp;
          float* dst = C + ((n * (2 * geo.h) + (2 * y + (p >> 1))) * (2 * geo.w) + (2 * x + (p & 1))) * (size_t)32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(acc[32 * jp + j], acc[32 * jp + j + 1], acc[32 * jp + j + 2], acc[32 * jp + j + 3]);
            if (bias) { o.x += bias[j]; o.y += bias[j + 1]; o.z += bias[j + 2]; o.w += bias[j + 3]; }
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        }
      }
      store = false;
    }
    if (geo.mode == MODE_GEMM && (geo.ksplits > 1 || geo.force_part)) {
      // ---- deterministic split-K: this split's partial tile goes to the workspace; splitk_reduce_kernel (launched right
      // behind this kernel) sums the partials of every output element in split order — no atomics, bit-reproducible
      float* prow = geo.part + ((size_t)blockIdx.z * geo.mpad + (size_t)(m0 + q * 32 + lane)) * geo.ldw + n0 + half * ACC_COLS;
#pragma unroll
      for (int j = 0; j < ACC_COLS; j += 4)
        *reinterpret_cast<float4*>(prow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      store = false;
    }
    if (row_ok && store) {
      const int cb = n0 + half * ACC_COLS;
      float* crow = C + row_off + cb;
      const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (cb + ACC_COLS <= N);
      if (vec) {
#pragma unroll
        for (int j = 0; j < ACC_COLS; j += 4) {
          float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          if (bias) { o.x += bias[cb + j]; o.y += bias[cb + j + 1]; o.z += bias[cb + j + 2]; o.w += bias[cb + j + 3]; }
          if (accumulate) { const float4 p = *reinterpret_cast<const float4*>(crow + j); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
          *reinterpret_cast<float4*>(crow + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < ACC_COLS; ++j) {
          if (cb + j < N) {
            float o = acc[j];
            if (bias) o += bias[cb + j];
            if (accumulate) o += crow[j];
            crow[j] = o;
          }
        }
      }
    }
    }   // tile loop
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side: tensor-map cache
struct MapKey {
  const void* ptr; int rows, cols, ld, box_rows;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<long long>()(((long long)k.rows << 32) ^ k.cols) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h ^= std::hash<long long>()(((long long)k.ld << 8) ^ k.box_rows) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapHash> g_maps;
std::mutex g_maps_mu;

// cuTensorMapEncodeTiled is a driver-API symbol: resolve it through the runtime at first use so that the library
// has no link-time dependency on libcuda.so (it must load on a GPU-less build host).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][cols] fp32, row stride ld; box = [box_rows][32], 128-byte swizzle, zero fill out of bounds
int get_map(const float* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out, bool atom32 = false) {
  MapKey key{ptr, rows, cols, ld, box_rows | (atom32 ? 1 << 16 : 0)};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return B200RL_OK; }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { b200rl_set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200RL_ERR_CUDA; }
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200rl_set_error("cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d", (int)r, rows, cols, ld);
    return B200RL_ERR_CUDA;
  }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return B200RL_OK;
}

struct Map4Key {
  const void* ptr; int C, W, H, N, bw, bh, bn, es;
  bool operator==(const Map4Key& o) const {
    return ptr == o.ptr && C == o.C && W == o.W && H == o.H && N == o.N && bw == o.bw && bh == o.bh && bn == o.bn && es == o.es;
  }
};
struct Map4Hash {
  size_t operator()(const Map4Key& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    const long long v[4] = {((long long)k.C << 32) ^ k.W, ((long long)k.H << 32) ^ k.N, ((long long)k.bw << 32) ^ k.bh,
                            ((long long)k.bn << 32) ^ k.es};
    for (long long x : v) h ^= std::hash<long long>()(x) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};
std::unordered_map<Map4Key, CUtensorMap, Map4Hash> g_maps4;

// channel-last image [N][H][W][C]; box = {32 ch, bw px, bh px, bn images} sampled with element stride es in W and H
int get_map4(const float* ptr, int C, int W, int H, int N, int bw, int bh, int bn, int es, CUtensorMap* out,
             bool atom32 = false) {
  Map4Key key{ptr, C, W, H, N, bw, bh, bn, es | (atom32 ? 1 << 8 : 0)};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps4.find(key);
  if (it != g_maps4.end()) { *out = it->second; return B200RL_OK; }
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * W * 4, (cuuint64_t)C * W * H * 4};
  // with an element stride e the box spans (count-1)*e+1 source elements and loads `count` of them
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)((bw - 1) * es + 1), (cuuint32_t)((bh - 1) * es + 1), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { b200rl_set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200RL_ERR_CUDA; }
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200rl_set_error("cuTensorMapEncodeTiled(4D) failed (%d) for image [%d,%d,%d,%d]", (int)r, N, H, W, C);
    return B200RL_ERR_CUDA;
  }
  if (g_maps4.size() > 1024) g_maps4.clear();
  g_maps4.emplace(key, m);
  *out = m;
  return B200RL_OK;
}

// tile shape on the small grid: as wide as possible, 128 pixels in total
bool conv_tile(int h, int w, int NB, int* bw, int* bh, int* bn) {
  if (w <= 0 || h <= 0 || (w & (w - 1)) || (h & (h - 1))) return false;
  *bw = w < 128 ? w : 128;
  int rest = 128 / *bw;
  *bh = h < rest ? h : rest;
  *bn = rest / *bh;
  return (w % *bw == 0) && (h % *bh == 0) && (*bw * *bh * *bn == 128) && (NB % *bn == 0);
}

// Persistent scheduling: with one CTA per SM (the operand stages fill shared memory) a launch of many short tiles pays
// TMEM allocation, barrier setup, pipeline fill and an un-overlapped epilogue per tile.  When there are more than two
// waves of tiles, launch about one CTA per SM and let each walk its M tiles (the epilogue of a tile overlaps the main
// loop of the next through the double-buffered TMEM accumulator).
static unsigned persistent_grid_y(int tiles, unsigned gz) {
  const long long total = (long long)tiles * gz;
  if (total <= 2LL * kNumSMs) return (unsigned)tiles;
  unsigned gy = kNumSMs / gz;
  if (gy < 1) gy = 1;
  return gy < (unsigned)tiles ? gy : (unsigned)tiles;
}

__global__ void conv_pack_down_kernel(const float* __restrict__ W, float* __restrict__ P, int Cs, int Cb) {
  // P[cs][tap][cb] = W[cs][cb][tap]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Cs * Cb * 16) return;
  const int cb = (int)(idx % Cb);
  const int tap = (int)((idx / Cb) % 16);
  const int cs = (int)(idx / ((long long)Cb * 16));
  P[idx] = W[((long long)cs * Cb + cb) * 16 + tap];
}
__global__ void conv_pack_up_kernel(const float* __restrict__ W, float* __restrict__ P, int Cs, int Cb) {
  // P[parity][cb][t][cs] = W[cs][cb][ky][kx], ky = (1-py)+2j, kx = (1-px)+2i, t = 2j+i
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Cs * Cb * 16) return;
  const int cs = (int)(idx % Cs);
  const int t = (int)((idx / Cs) % 4);
  const int cb = (int)((idx / ((long long)Cs * 4)) % Cb);
  const int par = (int)(idx / ((long long)Cs * 4 * Cb));
  const int py = par >> 1, px = par & 1;
  const int ky = (1 - py) + 2 * (t >> 1), kx = (1 - px) + 2 * (t & 1);
  P[idx] = W[((long long)cs * Cb + cb) * 16 + ky * 4 + kx];
}

__global__ void conv_pack_up4_kernel(const float* __restrict__ W, float* __restrict__ P, int Cs, int Cb) {
  // P[parity * Cb + cb][shift * Cs + cs] = W[cs][cb][ky][kx] when output parity (py, px) reads shift (dy, dx) (j = py - dy,
  // i = px - dx in {0, 1}; ky = (1 - py) + 2j, kx = (1 - px) + 2i), else 0.  shift = (dy + 1) * 3 + (dx + 1).
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 36LL * Cs * Cb) return;
  const int cs = (int)(idx % Cs);
  const int shift = (int)((idx / Cs) % 9);
  const int cb = (int)((idx / (9LL * Cs)) % Cb);
  const int par = (int)(idx / (9LL * Cs * Cb));
  const int py = par >> 1, px = par & 1, dy = shift / 3 - 1, dx = shift % 3 - 1;
  const int j = py - dy, i = px - dx;
  float v = 0.f;
  if (j >= 0 && j <= 1 && i >= 0 && i <= 1) v = W[((long long)cs * Cb + cb) * 16 + ((1 - py) + 2 * j) * 4 + (1 - px) + 2 * i];
  P[idx] = v;
}
static bool conv_up_merged(int Cb) { return Cb == 32; }

// Matmul precision of every tensor-core product of the library (process-wide, like torch.set_float32_matmul_precision):
// 3 = fp32-accurate 3xTF32 (default; what the 1e-4 parity tests run), 1 = single TF32 pass.
int g_passes = 3;

// Split-K workspace (partial tiles), one per device, grown on demand OUTSIDE stream capture (launches on one stream
// serialise, so consecutive products share it; a capture replays the size it was captured with).
struct SplitWs {
  float* part = nullptr;
  size_t floats = 0;
};
SplitWs g_split[16];
std::mutex g_split_mu;

int split_workspace(size_t need_floats, cudaStream_t st, float** part) {
  int dev = 0;
  RL_CUDA(cudaGetDevice(&dev));
  RL_CHECK_ARG(dev < 16, "split-K workspace: device index out of range");
  std::lock_guard<std::mutex> lk(g_split_mu);
  SplitWs& w = g_split[dev];
  if (need_floats > w.floats) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    RL_CUDA(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone) {
      b200rl_set_error("split-K workspace must grow during stream capture: run the step once eagerly first");
      return B200RL_ERR_CUDA;
    }
    RL_CUDA(cudaDeviceSynchronize());
    if (w.part) RL_CUDA(cudaFree(w.part));
    const size_t n = need_floats + need_floats / 2 > (size_t)16 << 20 ? need_floats + need_floats / 2 : (size_t)16 << 20;
    RL_CUDA(cudaMalloc(&w.part, n * sizeof(float)));
    w.floats = n;
  }
  *part = w.part;
  return B200RL_OK;
}

// C[m][n] (+)= bias[n] + sum_z part[z][m][n], z ascending: the fixed-order tail of a split-K product
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias, int M, int N,
                     int ldc, int ks, int mpad, int ldw, int accumulate) {
  const int n4 = (N + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * n4) return;
  const int m = (int)(idx / n4), n = (int)(idx - (long long)m * n4) * 4;
  const float* p = part + (size_t)m * ldw + n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4                                    // the splits' loads in flight together; the adds keep split order
  for (int z = 0; z < ks; ++z) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (size_t)z * mpad * ldw));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float* c = C + (size_t)m * ldc + n;
  const float o[4] = {acc.x, acc.y, acc.z, acc.w};
  if (n + 4 <= N && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
    float4 r = acc;
    if (bias) { r.x += bias[n]; r.y += bias[n + 1]; r.z += bias[n + 2]; r.w += bias[n + 3]; }
    if (accumulate) { const float4 q = *reinterpret_cast<const float4*>(c); r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w; }
    *reinterpret_cast<float4*>(c) = r;
  } else {
    for (int j = 0; j < 4 && n + j < N; ++j) {
      float r = o[j];
      if (bias) r += bias[n + j];
      if (accumulate) r += c[j];
      c[j] = r;
    }
  }
}

// one launch site for the four instantiations (tile width x TF32 passes)
#define LAUNCH_GEMM_TC(BN_, PASSES_, grid_, st_, ...)                                                                \
  do {                                                                                                               \
    const size_t smem__ = sizeof(Smem<BN_>) + 1024;                                                                  \
    RL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN_, PASSES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem__)); \
    gemm_tc_kernel<BN_, PASSES_><<<grid_, NTHREADS, smem__, st_>>>(__VA_ARGS__);                                      \
  } while (0)
#define DISPATCH_GEMM_TC(BN_val, passes_val, grid_, st_, ...)                                  \
  do {                                                                                          \
    if ((BN_val) == 64) {                                                                       \
      if ((passes_val) == 3) LAUNCH_GEMM_TC(64, 3, grid_, st_, __VA_ARGS__);                    \
      else LAUNCH_GEMM_TC(64, 1, grid_, st_, __VA_ARGS__);                                      \
    } else {                                                                                    \
      if ((passes_val) == 3) LAUNCH_GEMM_TC(128, 3, grid_, st_, __VA_ARGS__);                   \
      else LAUNCH_GEMM_TC(128, 1, grid_, st_, __VA_ARGS__);                                     \
    }                                                                                           \
  } while (0)

int launch_conv(int mode, const float* img, const float* Wp, float* out, const float* bias, int NB, int h, int w, int Cin,
                int Cout, cudaStream_t st) {
  TileGeo g = {};
  g.mode = mode; g.h = h; g.w = w; g.NB = NB; g.chunks = Cin / BK; g.Cout = Cout; g.ksplits = 1;
  g.passes = g_passes;
  RL_CHECK_ARG(conv_tile(h, w, NB, &g.bw, &g.bh, &g.bn), "image grid not tileable by 128 pixels");
  g.tiles_x = w / g.bw; g.tiles_y = h / g.bh;
  if (mode == MODE_UP && conv_up_merged(Cout)) mode = g.mode = MODE_UP4;
  const int taps = mode == MODE_DOWN ? 16 : (mode == MODE_UP4 ? 9 : 4);
  const int K = taps * Cin;
  CUtensorMap ma, mb;
  if (mode == MODE_DOWN) { if (int rc = get_map4(img, Cin, 2 * w, 2 * h, NB, g.bw, g.bh, g.bn, 2, &ma)) return rc; }
  else                   { if (int rc = get_map4(img, Cin, w, h, NB, g.bw, g.bh, g.bn, 1, &ma)) return rc; }
  const int BN = (mode == MODE_UP4) ? 128 : ((Cout <= 64) ? 64 : 128);
  const int brows = (mode == MODE_UP || mode == MODE_UP4) ? 4 * Cout : Cout;
  if (int rc = get_map(Wp, brows, K, K, BN, &mb)) return rc;
  const int mtiles = g.tiles_x * g.tiles_y * (NB / g.bn);
  g.mtiles = mtiles;
  g.ntiles = (mode == MODE_UP4) ? 1 : (Cout + BN - 1) / BN;
  dim3 grid(1, 1, mode == MODE_UP ? 4 : 1);
  grid.y = persistent_grid_y(mtiles * g.ntiles, grid.z);
  const int M = NB * h * w;  // unused by conv addressing; row validity comes from geo
  DISPATCH_GEMM_TC(BN, g.passes, grid, st, ma, mb, out, bias, M, Cout, K, Cout, 0, g);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

}  // namespace

extern "C" int b200rl_set_matmul_precision(int tf32_passes) {
  RL_CHECK_ARG(tf32_passes == 1 || tf32_passes == 3, "tf32_passes must be 3 (fp32-accurate 3xTF32) or 1 (single TF32 pass)");
  g_passes = tf32_passes;
  return B200RL_OK;
}
extern "C" int b200rl_get_matmul_precision(void) { return g_passes; }

// ---- convolution entry points (tensor-core implicit GEMM).  Wpacked: caller workspace of b200rl_conv_pack_floats() floats
// (16*Cs*Cb; 36*Cs*Cb for the merged-parity ConvTranspose2d layout used when Cb == 32).
extern "C" long long b200rl_conv_pack_floats(int mode_up, int Cs, int Cb) {
  return (mode_up && conv_up_merged(Cb) ? 36LL : 16LL) * Cs * Cb;
}
extern "C" int b200rl_conv_tc_supported(int mode_up, int NB, int h, int w, int Cs, int Cb) {
  int bw, bh, bn;
  if (!conv_tile(h, w, NB, &bw, &bh, &bn)) return 0;
  const int Cin = mode_up ? Cs : Cb, Cout = mode_up ? Cb : Cs;
  if (Cin % BK != 0 || Cout < 16 || Cout % 4 != 0) return 0;
  return 1;
}
extern "C" int b200rl_conv_pack(const float* W, float* Wpacked, int mode_up, int Cs, int Cb, cudaStream_t st) {
  RL_CHECK_ARG(W && Wpacked, "null pointer");
  const long long n = (long long)Cs * Cb * 16;
  if (mode_up && conv_up_merged(Cb)) conv_pack_up4_kernel<<<ceil_div(36LL * Cs * Cb, 256), 256, 0, st>>>(W, Wpacked, Cs, Cb);
  else if (mode_up) conv_pack_up_kernel<<<ceil_div(n, 256), 256, 0, st>>>(W, Wpacked, Cs, Cb);
  else conv_pack_down_kernel<<<ceil_div(n, 256), 256, 0, st>>>(W, Wpacked, Cs, Cb);
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}
extern "C" int b200rl_conv_down_tc(const float* big, const float* Wpacked, float* small_, int NB, int h, int w, int Cs,
                                   int Cb, cudaStream_t st) {
  RL_CHECK_ARG(big && Wpacked && small_, "null pointer");
  RL_CHECK_ARG(b200rl_conv_tc_supported(0, NB, h, w, Cs, Cb), "shape not eligible for the tensor-core conv path");
  return launch_conv(MODE_DOWN, big, Wpacked, small_, nullptr, NB, h, w, Cb, Cs, st);
}
extern "C" int b200rl_conv_up_tc(const float* small_, const float* Wpacked, float* big, const float* bias, int NB, int h,
                                 int w, int Cs, int Cb, cudaStream_t st) {
  RL_CHECK_ARG(big && Wpacked && small_, "null pointer");
  RL_CHECK_ARG(b200rl_conv_tc_supported(1, NB, h, w, Cs, Cb), "shape not eligible for the tensor-core conv path");
  return launch_conv(MODE_UP, small_, Wpacked, big, bias, NB, h, w, Cs, Cb, st);
}

// Shapes the tensor-core path accepts: NT product, 16-byte aligned operands with row strides that are
// multiples of 16 bytes, and enough work to fill a tile.
extern "C" int b200rl_gemm_tc_supported(const float* A, const float* B, int M, int N, int K, int lda, int ldb,
                                        int transA, int transB) {
  (void)transA; (void)transB;      // all four layouts: a transposed operand is read MN-major, in place
  // M < 128 (the 64-row products of a per-step RSSM scan at the XL width) still runs here: TMA zero-fills the missing
  // rows of the 128-row box and the epilogue masks them; half of the tile is idle but these products are bound by
  // streaming the weight matrix, which the SIMT path does 5-10x slower
  if (M < 32 || N < 48 || K < 32) return 0;
  if ((lda & 3) || (ldb & 3)) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 0;
  return 1;
}

namespace {
// `fused`: when non-null the product always leaves its result as split-K partial tiles (one split is allowed) and the
// caller launches the reduction itself, fused with what follows (LayerNorm, activation, GRU gate): *fused receives the
// workspace geometry.
struct FusedTail { float* part; int ks, mpad, ldw; };

int gemm_tc_impl(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                 int ldb, int ldc, int transA, int transB, int accumulate, cudaStream_t st, FusedTail* fused) {
  RL_CHECK_ARG(A && B && (C || fused), "null pointer");
  RL_CHECK_ARG(b200rl_gemm_tc_supported(A, B, M, N, K, lda, ldb, transA, transB), "shape not eligible for the tensor-core path");
  const int BN = (N <= 64) ? 64 : 128;
  CUtensorMap ma, mb;
  // A: [M][K] (K-major) or, transposed, [K][M] (MN-major: boxes of 32 k-rows x 32 m);  B: [N][K] or [K][N]
  if (!transA) { if (int rc = get_map(A, M, K, lda, BM, &ma)) return rc; }
  else         { if (int rc = get_map(A, K, M, lda, 32, &ma, true)) return rc; }
  if (transB)  { if (int rc = get_map(B, N, K, ldb, BN, &mb)) return rc; }
  else         { if (int rc = get_map(B, K, N, ldb, 32, &mb, true)) return rc; }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  TileGeo g = {};
  g.mode = MODE_GEMM;
  g.a_mn = transA ? 1 : 0;
  g.b_mn = transB ? 0 : 1;
  g.ksplits = 1;
  g.passes = g_passes;
  const int tiles = grid.x * grid.y, nkb = (K + BK - 1) / BK;
  if (tiles < 2 * kNumSMs && nkb >= 8) {
    // split-K for launches that cannot fill the SMs (the M = T*B = 1024 products of the imagination rollout, weight
    // gradients): pick the split count that minimises  waves x (k-blocks per CTA x t_kb + fixed cost)  with
    // t_kb ~ 0.8 us per 128x128x32 k-block, ~4 us of pipeline fill + epilogue per CTA, and for the fixed-order reduce
    // kernel behind a split product ~3 us + 0.3 us per split
    int best = 1;
    float best_t = 1e30f;
    const int max_sp = nkb / 4 < 32 ? nkb / 4 : 32;
    for (int sp = 1; sp <= (max_sp < 1 ? 1 : max_sp); ++sp) {
      const int per = (nkb + sp - 1) / sp, real = (nkb + per - 1) / per;
      const int waves = (tiles * real + kNumSMs - 1) / kNumSMs;
      const float t = waves * (per * 0.8f + 4.0f) + (real > 1 ? 3.0f + 0.3f * real : 0.0f);
      if (t < best_t - 1e-3f) { best_t = t; best = real; }
    }
    g.ksplits = best;
  }
  if (g.ksplits > 1 || fused) {
    grid.z = g.ksplits;
    g.mpad = (int)grid.y * BM;
    g.ldw = (int)grid.x * BN;
    if (int rc = split_workspace((size_t)g.ksplits * g.mpad * g.ldw, st, &g.part)) return rc;
    if (fused) { g.force_part = 1; *fused = FusedTail{g.part, g.ksplits, g.mpad, g.ldw}; }
  }
  g.mtiles = (int)grid.y;
  g.ntiles = (int)grid.x;
  grid.x = 1;
  grid.y = persistent_grid_y(g.mtiles * g.ntiles, grid.z);
  DISPATCH_GEMM_TC(BN, g.passes, grid, st, ma, mb, C, bias, M, N, K, ldc, accumulate, g);
  RL_CHECK_LAUNCH();
  if (g.ksplits > 1 && !fused) {
    splitk_reduce_kernel<<<ceil_div((long long)M * ((N + 3) / 4), 256), 256, 0, st>>>(g.part, C, bias, M, N, ldc, g.ksplits, g.mpad,
                                                                                      g.ldw, accumulate);
    RL_CHECK_LAUNCH();
  }
  return B200RL_OK;
}

// Fixed-order sum of the split-K partial tiles of a row, fused with LayerNorm (+ activation) or LayerNorm + GRU gate.
// One warp per row, the row in registers (NV float4 per lane, N = 128 * NV or less).
//   mode 0: out = act(LN(pre));  mode 1 (N = 3R, R % 128 == 0): LayerNormGRUCell gate (models.py:396-403) on the
//   normalised (reset | cand | update) thirds with h_prev -> h_out (and h_out2).  `pre` / `ln_out` are optional saves.
template <int NV>
__global__ void __launch_bounds__(256)
splitk_ln_kernel(const float* __restrict__ part, int ks, int mpad, int ldw, int M, int N, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int act, float* __restrict__ pre, long long ldpre,
                 float* __restrict__ ln_out, long long ldln, int mode, const float* __restrict__ h_prev, long long ldh,
                 float* __restrict__ h_out, long long ldho, float* __restrict__ h_out2, long long ldho2) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const int n4 = N >> 2;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4                                    // the splits' loads in flight together; the adds keep split order
  for (int z = 0; z < ks; ++z) {
    const float4* p = reinterpret_cast<const float4*>(part + ((size_t)z * mpad + row) * ldw);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < n4) { const float4 t = __ldcg(p + c); v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < n4) {
      if (pre) reinterpret_cast<float4*>(pre + row * ldpre)[c] = v[i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mu = warp_sum(s) / (float)N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < n4) {
      const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)N + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < n4) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[c], b = reinterpret_cast<const float4*>(beta)[c];
      v[i].x = (v[i].x - mu) * rstd * g.x + b.x; v[i].y = (v[i].y - mu) * rstd * g.y + b.y;
      v[i].z = (v[i].z - mu) * rstd * g.z + b.z; v[i].w = (v[i].w - mu) * rstd * g.w + b.w;
      if (mode == 0 && act == 1) {   // SiLU
        v[i].x = v[i].x / (1.f + expf(-v[i].x)); v[i].y = v[i].y / (1.f + expf(-v[i].y));
        v[i].z = v[i].z / (1.f + expf(-v[i].z)); v[i].w = v[i].w / (1.f + expf(-v[i].w));
      }
      if (ln_out) reinterpret_cast<float4*>(ln_out + row * ldln)[c] = v[i];
    }
  }
  if (mode == 1) {
    constexpr int NR = NV / 3;      // float4 per lane of one third (R = 128 * NR)
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int c = lane + 32 * i;
      const float4 hp = reinterpret_cast<const float4*>(h_prev + row * ldh)[c];
      const float4 gr = v[i], gc = v[i + NR], gu = v[i + 2 * NR];
      float4 h;
      auto gate = [](float r_, float c_, float u_, float hprev) {
        const float r = 1.f / (1.f + expf(-r_));
        const float cand = tanhf(r * c_);
        const float u = 1.f / (1.f + expf(-(u_ - 1.f)));
        return u * cand + (1.f - u) * hprev;
      };
      h.x = gate(gr.x, gc.x, gu.x, hp.x); h.y = gate(gr.y, gc.y, gu.y, hp.y);
      h.z = gate(gr.z, gc.z, gu.z, hp.z); h.w = gate(gr.w, gc.w, gu.w, hp.w);
      reinterpret_cast<float4*>(h_out + row * ldho)[c] = h;
      if (h_out2) reinterpret_cast<float4*>(h_out2 + row * ldho2)[c] = h;
    }
  }
}
}  // namespace

extern "C" int b200rl_gemm_tc(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                              int ldb, int ldc, int transA, int transB, int accumulate, cudaStream_t st) {
  return gemm_tc_impl(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, accumulate, st, nullptr);
}

extern "C" int b200rl_gemm_ln_supported(const float* A, const float* B, int M, int N, int K, int lda, int ldb, int mode) {
  if (!b200rl_gemm_tc_supported(A, B, M, N, K, lda, ldb, 0, 1)) return 0;
  if (N % 4 || N > 1536) return 0;
  if (mode == 1 && (N % 384 != 0)) return 0;      // three thirds of R = 128 * k columns each
  return 1;
}

extern "C" int b200rl_gemm_ln(const float* A, const float* W, int M, int N, int K, int lda, int ldw_, const float* gamma,
                              const float* beta, float eps, int act, float* pre, long long ldpre, float* out, long long ldout,
                              int mode, const float* h_prev, long long ldh, float* h_out, long long ldho, float* h_out2,
                              long long ldho2, cudaStream_t st) {
  RL_CHECK_ARG(A && W && gamma && beta, "null pointer");
  RL_CHECK_ARG(b200rl_gemm_ln_supported(A, W, M, N, K, lda, ldw_, mode), "shape not eligible for the fused product + LayerNorm");
  RL_CHECK_ARG(mode == 0 ? out != nullptr : (h_prev && h_out), "missing output");
  RL_CHECK_ARG(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(pre) |
                 reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(h_prev) | reinterpret_cast<uintptr_t>(h_out) |
                 reinterpret_cast<uintptr_t>(h_out2)) & 15) == 0 &&
                   ((ldpre | ldout | ldh | ldho | ldho2) & 3) == 0,
               "fused product + LayerNorm needs 16-byte aligned rows");
  FusedTail ft;
  if (int rc = gemm_tc_impl(A, W, nullptr, nullptr, M, N, K, lda, ldw_, N, 0, 1, 0, st, &ft)) return rc;
  const int blocks = ceil_div((long long)M * 32, 256);
#define LAUNCH_SPLITK_LN(NV_)                                                                                              \
  splitk_ln_kernel<NV_><<<blocks, 256, 0, st>>>(ft.part, ft.ks, ft.mpad, ft.ldw, M, N, gamma, beta, eps, act, pre, ldpre, out, \
                                                ldout, mode, h_prev, ldh, h_out, ldho, h_out2, ldho2)
  if (mode == 1) {
    if (N == 384) LAUNCH_SPLITK_LN(3); else if (N == 768) LAUNCH_SPLITK_LN(6); else if (N == 1152) LAUNCH_SPLITK_LN(9);
    else LAUNCH_SPLITK_LN(12);
  } else if (N <= 128) LAUNCH_SPLITK_LN(1);
  else if (N <= 256) LAUNCH_SPLITK_LN(2);
  else if (N <= 512) LAUNCH_SPLITK_LN(4);
  else if (N <= 1024) LAUNCH_SPLITK_LN(8);
  else LAUNCH_SPLITK_LN(12);
#undef LAUNCH_SPLITK_LN
  RL_CHECK_LAUNCH();
  return B200RL_OK;
}

// ---- convolution weight gradient on the tensor cores, operands read in place (no im2col, no transposes):
//   G[(tap, cb), cs] = sum over small pixels p of big[patch(p)][tap][cb] * small[p][cs]
// A = gathered big image, MN-major (rows (tap, cb), K = pixels); B = small [P][Cs], MN-major; split-K over pixels.
extern "C" int b200rl_conv_wgrad_mn_supported(int NB, int h, int w, int Cs, int Cb) {
  if (w <= 0 || h <= 0 || (w & (w - 1)) || (h & (h - 1))) return 0;
  const long long P = (long long)NB * h * w;
  if (Cb % 32 || Cs < 48 || Cs % 4 || P % 32 || P < 1024 || P > 2000000000LL) return 0;
  // the 32 pixels of a k-block must be one box: part of a row, whole rows of one image, or whole images
  const int bw = w < 32 ? w : 32, bh = (32 / bw) < h ? (32 / bw) : h, bn = 32 / (bw * bh);
  return (bw * bh * bn == 32) && (NB % bn == 0);
}

extern "C" int b200rl_conv_wgrad_mn(const float* small_, const float* big, float* G, int NB, int h, int w, int Cs, int Cb,
                                    cudaStream_t st) {
  RL_CHECK_ARG(small_ && big && G, "null pointer");
  RL_CHECK_ARG(b200rl_conv_wgrad_mn_supported(NB, h, w, Cs, Cb), "shape not eligible for the in-place wgrad path");
  const int P = NB * h * w, M = 16 * Cb, N = Cs;
  TileGeo g = {};
  g.mode = MODE_GEMM; g.a_mn = 1; g.b_mn = 1; g.wg = 1; g.passes = g_passes;
  g.h = h; g.w = w; g.NB = NB; g.Cout = Cb;
  g.bw = w < 32 ? w : 32; g.bh = (32 / g.bw) < h ? (32 / g.bw) : h; g.bn = 32 / (g.bw * g.bh);
  CUtensorMap ma, mb;
  if (int rc = get_map4(big, Cb, 2 * w, 2 * h, NB, g.bw, g.bh, g.bn, 2, &ma, true)) return rc;
  if (int rc = get_map(small_, P, Cs, Cs, 32, &mb, true)) return rc;
  const int BN = (N <= 64) ? 64 : 128;
  dim3 grid((N + BN - 1) / BN, M / BM);
  const int tiles = grid.x * grid.y, nkb = P / BK;
  int sp = (2 * kNumSMs + tiles - 1) / tiles;
  if (sp > nkb / 8) sp = nkb / 8;
  if (sp < 1) sp = 1;
  const int per = (nkb + sp - 1) / sp;
  g.ksplits = (nkb + per - 1) / per;
  grid.z = g.ksplits;
  g.mtiles = (int)grid.y;
  g.ntiles = (int)grid.x;
  if (g.ksplits > 1) {
    g.mpad = (int)grid.y * BM;
    g.ldw = (int)grid.x * BN;
    if (int rc = split_workspace((size_t)g.ksplits * g.mpad * g.ldw, st, &g.part)) return rc;
  }
  grid.x = 1;
  grid.y = (unsigned)(g.mtiles * g.ntiles);
  DISPATCH_GEMM_TC(BN, g.passes, grid, st, ma, mb, G, nullptr, M, N, P, N, 0, g);
  RL_CHECK_LAUNCH();
  if (g.ksplits > 1) {
    splitk_reduce_kernel<<<ceil_div((long long)M * ((N + 3) / 4), 256), 256, 0, st>>>(g.part, G, nullptr, M, N, N, g.ksplits, g.mpad,
                                                                                      g.ldw, 0);
    RL_CHECK_LAUNCH();
  }
  return B200RL_OK;
}
