// Warp-specialised tcgen05 / TMA / TMEM GEMM for sm_100a, bf16 operands, fp32 accumulate.
//
//   D[M,N] = A[M,K] * B[N,K]^T          (each operand either K-major or MN-major in HBM)
//
// One persistent CTA per SM, 576 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma 128xBNx16, fp32 accumulators in TMEM,
//                               tcgen05.commit releases smem slots / publishes the accumulator)
//   warps 2..17 epilogue       (tcgen05.ld 32x32b: thread == accumulator row; the Epi functor consumes
//                               16-column chunks, so row-wise reductions such as the catalog softmax
//                               statistics are thread-local; the four warps that share a TMEM sub-partition
//                               take interleaved chunks. ncu showed the epilogue, not the MMA, sets the tile
//                               time and that it is issue-latency bound: 4 warps per scheduler, <=113 registers)
// The TMEM accumulator is double buffered (2*BN columns) so the epilogue of tile i overlaps the
// mainloop of tile i+1. Split-K is supported (tile index carries the split; Epi decides how to merge).
//
// Operand layouts ("major"), as seen by the tensor core:
//   K-major  operand X[MN][K]  (K contiguous in HBM): TMA box {64 k, rows}, smem rows of 128 B,
//            UMMA descriptor SWIZZLE_128B, SBO = 1024 B, k-step advances the start address by 32 B.
//   MN-major operand X[K][MN]  (MN contiguous in HBM): TMA boxes {64 mn, 64 k} (8 KB each),
//            UMMA descriptor SWIZZLE_128B, LBO = 8 KB (next 64-wide MN chunk), SBO = 1 KB
//            (next 8 k), k-step (16 k) advances the start address by 2 KB.
// The second form is what lets the backward GEMMs (dgrad over the catalog, wgrad over the batch)
// read the very same HBM tensors the forward wrote, with no transposed copies.
#pragma once
#include <cuda.h>
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace ltg {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 16;                      // four per TMEM sub-partition, interleaved 16-column chunks
constexpr int GEMM_CW = 16;                             // epilogue chunk width (columns per tcgen05.ld)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// Same load, delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
// Tile store: shared memory (the box layout of the tensor map, e.g. 128B-swizzled rows) -> global; rows / columns outside the tensor
// are clipped. Completion is tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes complete
// One lane of a converged warp (elect.sync). The single-thread roles (TMA producer, MMA issuer) run their loops with the WHOLE warp and
// gate only the issuing instructions with this predicate: every operand (shared-memory descriptors, TMEM address, barrier address) is
// then warp-uniform and lives in uniform registers. Written as `if (lane == 0) { loop }` the same code made the compiler treat every
// operand as divergent -- five R2UR.BROADCAST + ELECT + a BRA.U.ANY "waterfall" loop per tcgen05.mma, ~120 dependent scalar
// instructions per k block in ONE thread: ncu (round 2) showed the MMA warp busy issuing, the producer waiting for free slots, the
// epilogue warps waiting for accumulators, and 323 cycles per 128x256x16 MMA against a tensor-pipe floor of 128.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Same, arriving on the barrier at this offset in every CTA of cta_mask (slot release for multicast operands).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (cute/arch/mma_sm100_desc.hpp documents the bit fields)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_base() {
  // version = 1 (bits 46..47), layout = SWIZZLE_128B (2, bits 61..63), SBO = 1024 B (bits 32..45)
  return ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)(1024 >> 4) << 32);
}
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t saddr) {  // K-major, LBO ignored (=1)
  return umma_desc_base() | ((uint64_t)1 << 16) | (uint64_t)((saddr & 0x3FFFF) >> 4);
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return umma_desc_base() | ((uint64_t)(lbo_bytes >> 4) << 16) | (uint64_t)((saddr & 0x3FFFF) >> 4);
}
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4)                    // D format fp32
         | (1u << 7) | (1u << 10)     // A, B = bf16
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// Problem description
// ----------------------------------------------------------------------------------------------
struct GemmShape {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks;  // 128-row blocks, BN-col blocks, 64-deep k blocks
  int splits, kb_per_split;          // split-K: split s covers k blocks [s*kb_per_split, ...)
};

// EPI_SMEM: per-CTA scratch an epilogue asks for (Epi::kSmem); it comes out of the pipeline's stage budget
template <int BN, int EPI_SMEM = 0>
__host__ __device__ constexpr int gemm_stages() {
  constexpr int per_stage = GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2;
  constexpr int s = (196608 - EPI_SMEM) / per_stage;
  return s > 8 ? 8 : s;
}

template <int BN, int EPI_SMEM = 0>
__host__ __device__ constexpr size_t gemm_smem_bytes() {
  return (size_t)gemm_stages<BN, EPI_SMEM>() * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_SMEM;
}

// ----------------------------------------------------------------------------------------------
// The kernel. Epi: struct with `Params`, ctor(const Params&, row, n0, n_blk, split, shape),
// `chunk(col0, v[32])`, `finish()`.
// ----------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int CM, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmShape shape, const __grid_constant__ typename Epi::Params ep) {
  constexpr int STAGES = gemm_stages<BN, Epi::kSmem>();
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = umma_idesc(GEMM_BM, BN, A_MN, B_MN);
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be 64, 128, 192 or 256");  // UMMA N: multiple of 16 up to 256; B boxes are 64 wide
  static_assert(CM == 1 || CM == 2 || CM == 4, "cluster size along M");
  // CM > 1: the CM CTAs of a cluster work on CM consecutive 128-row blocks of the SAME n-block / k-range. Each loads its own A
  // tile and 1/CM of the shared B tile, multicast to all of them, so the B operand crosses L2->SM once per cluster instead of
  // once per CTA. A stage may be refilled only when every CTA of the cluster has consumed it: tcgen05.commit multicasts the
  // release to all CM empty barriers (init count CM).
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CM) - 1u);
  const uint32_t crank = CM > 1 ? cluster_ctarank() : 0u;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_scratch = smem + STAGES * (A_BYTES + B_BYTES) + 256;   // Epi::kSmem bytes behind the barrier block

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_trigger();   // (PDL, ltg_common.cuh) the prologue below touches no global data: it overlaps the previous kernel's tail
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CM); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], GEMM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if constexpr (CM > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_cta();      // every prerequisite grid has completed and its writes are visible from here on

  // cluster tiles: (group of CM row blocks, n block, split); CTA `crank` of the cluster takes row block m_group*CM + crank
  const int m_groups = (shape.m_blocks + CM - 1) / CM;
  const int num_tiles = m_groups * shape.n_blocks * shape.splits;
  const int tile0 = blockIdx.x / CM;
  const int tile_stride = gridDim.x / CM;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in the loop, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_stride) {
        const int m_blk = (t % m_groups) * CM + (int)crank;
        const int rest = t / m_groups;
        const int n_blk = rest % shape.n_blocks;
        const int split = rest / shape.n_blocks;
        const int kb0 = split * shape.kb_per_split;
        const int kb1 = min(shape.k_blocks, kb0 + shape.kb_per_split);
        const int m0 = m_blk * GEMM_BM, n0 = n_blk * BN;   // m0 may lie beyond M for the padding block of the last group: TMA zero-fills
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
            uint8_t* sa = smem_a + stage * A_BYTES;
            uint8_t* sb = smem_b + stage * B_BYTES;
            const int k0 = kb * GEMM_BK;
            if constexpr (A_MN) {
#pragma unroll
              for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, k0);
            } else {
              tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
            }
            if constexpr (CM == 1) {
              if constexpr (B_MN) {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, k0);
              } else {
                tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
              }
            } else if constexpr (B_MN) {
              // every 64(n) x 64(k) box is split along k: this CTA fetches k rows [crank*64/CM, +64/CM) of each box for everybody
              constexpr int KR = 64 / CM;
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d_mc(sb + j * 8192 + crank * (KR * 128), &tmB, &full_bar[stage], n0 + j * 64, k0 + (int)crank * KR, MC_MASK);
            } else {
              // the BN x 64(k) tile is split along n: this CTA fetches rows [crank*BN/CM, +BN/CM) for everybody
              constexpr int NR = BN / CM;
              tma_load_2d_mc(sb + crank * (NR * 128), &tmB, &full_bar[stage], k0, n0 + (int)crank * NR, MC_MASK);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
      for (int t = tile0; t < num_tiles; t += tile_stride) {
        const int split = (t / m_groups) / shape.n_blocks;
        const int kb0 = split * shape.kb_per_split;
        const int kb1 = min(shape.k_blocks, kb0 + shape.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (kb0 >= kb1) {  // empty split (fixed split count, short K): the epilogue stores zeros
          if (elect_one()) umma_commit(&tfull_bar[acc]);
          __syncwarp();
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = sa0 + stage * A_BYTES;
          const uint32_t sb = sb0 + stage * B_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              const uint64_t da = A_MN ? umma_desc_mn(sa + k * 2048, 8192) : umma_desc_k(sa + k * 32);
              const uint64_t db = B_MN ? umma_desc_mn(sb + k * 2048, 8192) : umma_desc_k(sb + k * 32);
              umma_bf16(tmem_d, da, db, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if constexpr (CM == 1) umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
            else umma_commit_mc(&empty_bar[stage], MC_MASK);       // ... in every CTA of the cluster
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (16 warps; TMEM sub-partition = warp % 4, chunk phase = quarter) =====================
    const int sub = warp & 3;  // hardware rule: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int quarter = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = tile0; t < num_tiles; t += tile_stride) {
      const int m_blk = (t % m_groups) * CM + (int)crank;
      const int rest = t / m_groups;
      const int n_blk = rest % shape.n_blocks;
      const int split = rest / shape.n_blocks;
      const int n0 = n_blk * BN;
      const int row = m_blk * GEMM_BM + sub * 32 + lane;
      // The epilogue functor may need global data per chunk that does not depend on the accumulator (bias slice, stored activation for
      // the fused tanh'/dropout backward). ncu (round 2) showed the epilogue warps stalled on exactly those loads, one L2 round trip
      // per chunk (30% of all samples of the dz12 GEMM, 21% of the decoder forward): they are software pipelined -- the first chunk's
      // loads are issued BEFORE waiting for the accumulator, chunk i+1's before chunk i is processed.
      Epi epi(ep, row, n0, n_blk * 4 + quarter, split, shape, epi_scratch);
      int c = quarter * GEMM_CW;
      bool have = c < BN && n0 + c < shape.N;   // warp-uniform
      if (have) epi.preload(n0 + c);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      while (have) {
        const int cn = c + 4 * GEMM_CW;
        const bool next = cn < BN && n0 + cn < shape.N;
        float v[GEMM_CW];
        if (split * shape.kb_per_split < shape.k_blocks) {
          tmem_ld16(taddr + c, v);
        } else {
#pragma unroll
          for (int i = 0; i < GEMM_CW; ++i) v[i] = 0.f;  // empty split: nothing was accumulated
        }
        epi.chunk(n0 + c, v, next ? n0 + cn : -1);
        c = cn; have = next;
      }
      epi.finish();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CM > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// Epilogues
// ----------------------------------------------------------------------------------------------

// MUFU.TANH: max abs error ~5e-4, below the bf16 rounding of every tensor the GEMM epilogues write through tanh
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MUFU.EX2 without the denormal-range fix-up exp2f() wraps around it (FSETP + two FMULs per call: a quarter of the arithmetic of the
// catalog-softmax epilogue in the ncu instruction mix of round 2). Arguments here are <= 0 and results feed a sum of at least one 1.0.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Generic: out = dropout(act(alpha*acc + bias[col])) -> fp32 and/or bf16, or fp32 atomic accumulate (split-K).
// A designated column (`aux_col`) can be diverted to `aux_out[row]` (used to get column sums for free
// from a ones-column in the B operand).
struct EpiStore {
  struct Params {
    float* out_f32; int ld_f32;
    __nv_bfloat16* out_bf16; int ld_bf16;
    const float* bias;      // [N] or null
    int act;                // 0 none, 1 tanh
    int atomic;             // 1: red.add into out_f32 (which the caller zeroed)
    int64_t split_stride;   // > 0: split s stores its partial tile at out_f32 + s*split_stride (the consumer sums the partials)
    float alpha;
    float keep;             // dropout keep prob; >= 1 or <= 0 disables
    uint64_t seed; uint32_t rng_stream, rng_step; const uint32_t* rng_step_dev; int rng_ld;  // group idx = (row*rng_ld + col)/4
    int aux_col; float* aux_out;  // aux_col < 0 disables
    const __nv_bfloat16* dact_src; int dact_ld; float dact_keep;  // multiply by d/da dropout(tanh(a)) recovered from the stored activation
  };
  const Params& p;
  int row, M, N, split;
  uint32_t thr16, key;
  float inv_keep;
  bool drop;
  uint4 pre0, pre1;   // stored activation of the NEXT chunk (dact_src), loaded one chunk ahead
  bool pre_vec;
  static constexpr int kSmem = 0;
  __device__ __forceinline__ void preload(int col0) {
    pre_vec = false;
    if (p.dact_src == nullptr || row >= M) return;
    const __nv_bfloat16* src = p.dact_src + (size_t)row * p.dact_ld + col0;
    if (col0 + GEMM_CW <= N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      pre0 = __ldg(reinterpret_cast<const uint4*>(src));
      pre1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
      pre_vec = true;
    }
  }
  __device__ EpiStore(const Params& p_, int row_, int, int, int split_, const GemmShape& s, uint8_t*) : p(p_), row(row_), M(s.M), N(s.N), split(split_) {
    drop = p.keep > 0.f && p.keep < 1.f;
    thr16 = drop ? ltg_keep_threshold16(p.keep) : 65536u;
    inv_keep = drop ? 1.0f / p.keep : 1.0f;
    key = 0;
    if (drop) key = ltg_hash_key(p.seed, p.rng_stream, p.rng_step + (p.rng_step_dev != nullptr ? *p.rng_step_dev : 0u));
  }
  __device__ void chunk(int col0, float (&v)[GEMM_CW], int next_col0) {
    if (row >= M) return;
    constexpr int CW = GEMM_CW;
    const bool full = col0 + CW <= N;
    const uint4 cur0 = pre0, cur1 = pre1;
    const bool cur_vec = pre_vec;
    if (next_col0 >= 0) preload(next_col0);
    if (p.bias != nullptr) {
      if (full && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0)) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int i = 0; i < CW / 4; ++i) {
          const float4 b = __ldg(b4 + i);
          v[4 * i] = fmaf(v[4 * i], p.alpha, b.x); v[4 * i + 1] = fmaf(v[4 * i + 1], p.alpha, b.y);
          v[4 * i + 2] = fmaf(v[4 * i + 2], p.alpha, b.z); v[4 * i + 3] = fmaf(v[4 * i + 3], p.alpha, b.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = fmaf(v[i], p.alpha, (col0 + i < N) ? __ldg(p.bias + col0 + i) : 0.f);
      }
    } else if (p.alpha != 1.0f) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] *= p.alpha;
    }
    if (p.act == 1) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = tanh_approx(v[i]);
    }
    if (p.dact_src != nullptr) {
      // backward through y = dropout(tanh(a)) stored post-dropout: dy/da = mask/keep * (1 - tanh^2) = (1 - (y*keep)^2)/keep where y != 0
      const __nv_bfloat16* src = p.dact_src + (size_t)row * p.dact_ld + col0;
      const bool dd = p.dact_keep > 0.f && p.dact_keep < 1.f;
      const float kk = dd ? p.dact_keep : 1.0f, ik = 1.0f / kk;
      if (cur_vec) {
#pragma unroll
        for (int q = 0; q < CW / 8; ++q) {
          const uint4 u = q == 0 ? cur0 : cur1;
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 y = unpack_bf16x2(w[j]);
            const float t0 = y.x * kk, t1 = y.y * kk;
            v[8 * q + 2 * j] *= (dd && y.x == 0.f) ? 0.f : (1.0f - t0 * t0) * ik;
            v[8 * q + 2 * j + 1] *= (dd && y.y == 0.f) ? 0.f : (1.0f - t1 * t1) * ik;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          if (col0 + i < N) {
            const float y = __bfloat162float(src[i]);
            const float t = y * kk;
            v[i] *= (dd && y == 0.f) ? 0.f : (1.0f - t * t) * ik;
          }
        }
      }
    }
    if (drop) {
      // col0 % 16 == 0 and rng_ld % 4 == 0: one 64-bit hash per group of four adjacent columns
      const uint64_t gbase = ((uint64_t)row * (uint64_t)p.rng_ld + (uint64_t)col0) >> 2;
      const uint32_t thr_hi = thr16 << 16;   // (h >> 16) < thr16  <=>  h < thr16 << 16   (thr16 <= 65535 when drop)
#pragma unroll
      for (int q = 0; q < CW / 4; ++q) {
        const uint2 h = ltg_hash_quad(key, gbase + q);
        v[4 * q] = (h.x & 0xFFFFu) < thr16 ? v[4 * q] * inv_keep : 0.f;
        v[4 * q + 1] = h.x < thr_hi ? v[4 * q + 1] * inv_keep : 0.f;
        v[4 * q + 2] = (h.y & 0xFFFFu) < thr16 ? v[4 * q + 2] * inv_keep : 0.f;
        v[4 * q + 3] = h.y < thr_hi ? v[4 * q + 3] * inv_keep : 0.f;
      }
    }
    if (p.aux_col >= col0 && p.aux_col < col0 + CW) {
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (col0 + i == p.aux_col) {
          if (p.atomic) atomicAdd(p.aux_out + row, v[i]); else p.aux_out[row] = v[i];
        }
    }
    const int nlim = (p.aux_col >= 0 && p.aux_col < N) ? p.aux_col : N;  // columns >= aux_col are not part of `out`
    if (p.out_f32 != nullptr) {
      float* o = p.out_f32 + (size_t)split * (size_t)p.split_stride + (size_t)row * p.ld_f32 + col0;
      if (p.atomic) {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (col0 + i < nlim) atomicAdd(o + i, v[i]);
      } else if (col0 + CW <= nlim && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < CW; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (col0 + i < nlim) o[i] = v[i];
      }
    }
    if (p.out_bf16 != nullptr) {
      __nv_bfloat16* o = p.out_bf16 + (size_t)row * p.ld_bf16 + col0;
      if (col0 + CW <= nlim && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < CW; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(v[i], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
          u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
          *reinterpret_cast<uint4*>(o + i) = u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (col0 + i < nlim) o[i] = __float2bfloat16(v[i]);
      }
    }
  }
  __device__ void finish() {}
};

// Weight-gradient GEMM whose epilogue IS the optimizer step (decoder W_p1^T [I,600], single GPU): the accumulator of output
// element (item, col) is the complete gradient (K = batch, no split), so TF-Adam (adam_kernels.cu: same formula, same operation
// order) is applied in place: p, m, v fp32 [M, H] read and written once, bf16 shadow written, and the [I,600] fp32 gradient never
// exists in HBM (-8 B/param of the 30 B/param the separate gradient store + Adam sweep moved; one 60 us kernel less per step).
// Thread == row, 16 consecutive columns == 64 contiguous bytes per array. Column `aux_col` (the ones column of the B operand)
// carries the bias gradient and goes to aux_out[row].
struct EpiAdam {
  struct Params {
    float* p; float* m; float* v; __nv_bfloat16* shadow; int ld;   // [M, ld], ld % 4 == 0
    int n_cols;                                                    // columns < n_cols are parameters (n_cols % 4 == 0)
    int aux_col; float* aux_out;
    float lr_t; const float* scal; float b1, b2, eps;
  };
  static constexpr int kSmem = GEMM_EPI_WARPS * 32 * 17 * 4;   // per-warp 32 x 16 fp32 transpose tile (pitch 17: conflict-free)
  const Params& p;
  int row, M;
  float lr_t;
  float* sw;
  int n_end;
  __device__ EpiAdam(const Params& p_, int row_, int n0, int, int, const GemmShape& s, uint8_t* scratch) : p(p_), row(row_), M(s.M) {
    lr_t = p.lr_t < 0.f ? p.scal[LTG_S_LR_T] : p.lr_t;
    sw = reinterpret_cast<float*>(scratch) + ((threadIdx.x >> 5) - 2) * (32 * 17);
    n_end = n0 + 128;                                            // launched with BN = 128 only
    const int quarter = ((threadIdx.x >> 5) - 2) >> 2;
    prefetch_chunk(n0 + quarter * GEMM_CW);                      // this warp's first chunk of the tile
  }
  // L2 prefetch of the p/m/v pieces this lane will load for the chunk at col0. ncu on the first version: 2.9 TB/s, IPC 0.85 --
  // every warp walked load -> wait ~1 us -> update -> store chunk after chunk; with the next chunk already on its way to L2 the
  // loads cost an L2 hit.
  __device__ __forceinline__ void prefetch_chunk(int col0) const {
    const int lane = threadIdx.x & 31;
    const int c4 = (lane & 3) * 4;
    if (col0 + c4 + 4 > p.n_cols) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row - lane + (lane >> 2) + 8 * j;
      if (r < M) {
        const size_t off = (size_t)r * p.ld + col0 + c4;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.p + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.m + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.v + off));
      }
    }
  }
  // The accumulator arrives with thread == row (64 B per thread per array, 2400 B apart: 32 half-used sectors per request, measured
  // 3x slower than the separate Adam sweep). The warp therefore transposes its 32 x 16 chunk through shared memory so that four
  // lanes cover the 64 contiguous bytes of a row and a request touches 8 rows x 64 B = 16 full sectors.
  __device__ __forceinline__ void preload(int) {}
  __device__ void chunk(int col0, float (&g)[GEMM_CW], int) {
    const int lane = threadIdx.x & 31;
    if (row < M && p.aux_col >= col0 && p.aux_col < col0 + GEMM_CW) {
#pragma unroll
      for (int i = 0; i < GEMM_CW; ++i)
        if (col0 + i == p.aux_col) p.aux_out[row] = g[i];
    }
    if (col0 + 4 * GEMM_CW < n_end) prefetch_chunk(col0 + 4 * GEMM_CW);   // the chunk this warp handles next (quarters interleave)
#pragma unroll
    for (int i = 0; i < GEMM_CW; ++i) sw[lane * 17 + i] = g[i];
    __syncwarp();
    const int row_base = row - lane;
    const int c4 = (lane & 3) * 4;
    const bool col_ok = col0 + c4 + 4 <= p.n_cols;
    float4 pp[4], mm[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                // all loads first: 12 independent 16-byte requests per thread
      const int r = row_base + (lane >> 2) + 8 * j;
      if (col_ok && r < M) {
        const size_t off = (size_t)r * p.ld + col0 + c4;
        pp[j] = ld_stream_f4(p.p + off); mm[j] = ld_stream_f4(p.m + off); vv[j] = ld_stream_f4(p.v + off);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = (lane >> 2) + 8 * j;
      const int r = row_base + rl;
      if (col_ok && r < M) {
        const float gx = sw[rl * 17 + c4], gy = sw[rl * 17 + c4 + 1], gz = sw[rl * 17 + c4 + 2], gw = sw[rl * 17 + c4 + 3];
        float4 m4 = mm[j], v4 = vv[j], p4 = pp[j];
        ltg_adam4(p4, m4, v4, make_float4(gx, gy, gz, gw), lr_t, p.b1, p.b2, p.eps);
        const size_t off = (size_t)r * p.ld + col0 + c4;
        st_stream_f4(p.p + off, p4); st_stream_f4(p.m + off, m4); st_stream_f4(p.v + off, v4);
        uint2 sh; sh.x = pack_bf16x2(p4.x, p4.y); sh.y = pack_bf16x2(p4.z, p4.w);
        *reinterpret_cast<uint2*>(p.shadow + off) = sh;
      }
    }
    __syncwarp();
  }
  __device__ void finish() {}
};

// Decoder forward: logits = acc + b_dec -> bf16 stash, plus per-(n-block,chunk-parity,row) softmax partials
// (row max, sum exp(x - max)) so the [B, I] logits never exist in fp32 in HBM.
// Restates MultiVAE.py:169 (matmul + bias) and the reductions inside log_softmax/softmax (108, 143).
struct EpiLogitsStats {
  struct Params {
    __nv_bfloat16* logits; int ld;   // [M, ld] bf16 (may be null: statistics only)
    const float* bias;               // [N]
    float2* partial;                 // [4*n_blocks, M] (max, sumexp): one entry per (n block, chunk phase); null: logits only
  };
  const Params& p;
  int row, M, N, slot;
  float mx, sum;
  float4 pb[GEMM_CW / 4];   // bias slice of the NEXT chunk, loaded one chunk ahead
  bool pre_vec;
  static constexpr int kSmem = 0;
  __device__ __forceinline__ void preload(int col0) {
    pre_vec = false;
    if (row >= M) return;
    if (col0 + GEMM_CW <= N && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0)) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < GEMM_CW / 4; ++i) pb[i] = __ldg(b4 + i);
      pre_vec = true;
    }
  }
  __device__ EpiLogitsStats(const Params& p_, int row_, int, int slot_, int, const GemmShape& s, uint8_t*)
      : p(p_), row(row_), M(s.M), N(s.N), slot(slot_), mx(-INFINITY), sum(0.f) {}
  __device__ void chunk(int col0, float (&v)[GEMM_CW], int next_col0) {
    if (row >= M) return;
    constexpr int CW = GEMM_CW;
    constexpr float LOG2E = 1.4426950408889634f;
    if (pre_vec) {
      float4 cb[CW / 4];
#pragma unroll
      for (int i = 0; i < CW / 4; ++i) cb[i] = pb[i];
      if (next_col0 >= 0) preload(next_col0);
#pragma unroll
      for (int i = 0; i < CW / 4; ++i) {
        const float4 b = cb[i];
        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    } else {
      if (next_col0 >= 0) preload(next_col0);
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = (col0 + i < N) ? v[i] + __ldg(p.bias + col0 + i) : -INFINITY;
    }
    if (p.partial != nullptr) {   // (phase A only samples from the logits: no statistics wanted)
      // four independent chains for the max and for the sum
      float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
      for (int i = 4; i < CW; i += 4) {
        m0 = fmaxf(m0, v[i]); m1 = fmaxf(m1, v[i + 1]); m2 = fmaxf(m2, v[i + 2]); m3 = fmaxf(m3, v[i + 3]);
      }
      const float nm = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));  // finite: every processed chunk has col0 < N
      const float nml = nm * LOG2E;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        s0 += ex2_approx(fmaf(v[i], LOG2E, -nml)); s1 += ex2_approx(fmaf(v[i + 1], LOG2E, -nml));
        s2 += ex2_approx(fmaf(v[i + 2], LOG2E, -nml)); s3 += ex2_approx(fmaf(v[i + 3], LOG2E, -nml));
      }
      sum = sum * ex2_approx((mx - nm) * LOG2E) + ((s0 + s1) + (s2 + s3));   // first chunk: mx = -inf -> factor 0
      mx = nm;
    }
    if (p.logits != nullptr) {
      __nv_bfloat16* o = p.logits + (size_t)row * p.ld + col0;
      if (col0 + CW <= N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < CW; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(v[i], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
          u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
          *reinterpret_cast<uint4*>(o + i) = u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (col0 + i < N) o[i] = __float2bfloat16(v[i]);
      }
    }
  }
  __device__ void finish() {
    if (row < M && p.partial != nullptr) p.partial[(size_t)slot * M + row] = make_float2(mx, sum);
  }
};

// ----------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled ltg_get_encode_tiled();

// bf16 2-D tensor [outer][inner] with row pitch ld (elements), box {box_inner (=64), box_outer}, 128B swizzle, zero OOB fill.
inline int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled fn = ltg_get_encode_tiled();
  if (fn == nullptr) return LTG_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    ltg_set_last_error("TMA operand must be 16-byte aligned with a 16-byte-multiple pitch", __FILE__, __LINE__);
    return LTG_ERR_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ltg_set_last_error("cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    return LTG_ERR_DRIVER;
  }
  return LTG_OK;
}

int ltg_num_sms();

int ltg_gemm_cluster_override();  // -1: heuristic; 1/2/4 forces the cluster size (env LTG_GEMM_CM, experiments)

template <int BN, bool A_MN, bool B_MN, int CM, class Epi>
int launch_gemm_cm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, const GemmShape& s, const typename Epi::Params& ep,
                   cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc;
  if (A_MN) rc = make_tmap_bf16(&tmA, A, (uint64_t)s.M, (uint64_t)s.K, (uint64_t)lda, 64, 64);
  else rc = make_tmap_bf16(&tmA, A, (uint64_t)s.K, (uint64_t)s.M, (uint64_t)lda, 64, GEMM_BM);
  if (rc) return rc;
  if (B_MN) rc = make_tmap_bf16(&tmB, B, (uint64_t)s.N, (uint64_t)s.K, (uint64_t)ldb, 64, 64 / CM);
  else rc = make_tmap_bf16(&tmB, B, (uint64_t)s.K, (uint64_t)s.N, (uint64_t)ldb, 64, BN / CM);
  if (rc) return rc;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, CM, Epi>;
  static int max_clusters = 0;  // per instantiation
  const size_t smem = gemm_smem_bytes<BN, Epi::kSmem>();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CM; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = ltg_pdl_enabled() ? 2 : 1;
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
    if (CM > 1) {
      cfg.gridDim = dim3(ltg_num_sms() / CM * CM);
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (e != cudaSuccess || n <= 0) { ltg_set_last_error("cudaOccupancyMaxActiveClusters failed", __FILE__, __LINE__); return LTG_ERR_CUDA; }
      max_clusters = n;
    } else {
      max_clusters = ltg_num_sms();
    }
  }
  const int m_groups = (s.m_blocks + CM - 1) / CM;
  const int tiles = m_groups * s.n_blocks * s.splits;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cfg.gridDim = dim3(clusters * CM);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, s, ep);
  if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
  return LTG_OK;
}

// A: [M,K] (a_mn=false, pitch lda over K) or stored [K,M] (a_mn=true, pitch lda over M). Same for B with N.
template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K, int splits,
                const typename Epi::Params& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LTG_OK;
  GemmShape s;
  s.M = M; s.N = N; s.K = K;
  s.m_blocks = (M + GEMM_BM - 1) / GEMM_BM;
  s.n_blocks = (N + BN - 1) / BN;
  s.k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  if (splits < 1) splits = 1;
  s.kb_per_split = (s.k_blocks + splits - 1) / splits;
  s.splits = splits;  // trailing splits may be empty (kb0 >= k_blocks): they store zeros, so a consumer can rely on a fixed count
  // cluster along M: row blocks of the same n-block share the B operand (TMA multicast). 4-CTA clusters can only be
  // placed on 132 of the 148 SMs (GPCs with 18 SMs strand two), so they are used when the problem has exactly one group of
  // 3-4 row blocks (batch 500: the whole M extent shares every B tile); longer M runs as CTA pairs on all 148 SMs.
  int cm = (s.m_blocks == 3 || s.m_blocks == 4) ? 4 : (s.m_blocks >= 2 ? 2 : 1);
  const int ov = ltg_gemm_cluster_override();
  if (ov == 1 || ov == 2 || ov == 4) cm = ov;
  if (cm == 4) return launch_gemm_cm<BN, A_MN, B_MN, 4, Epi>(A, lda, B, ldb, s, ep, stream);
  if (cm == 2) return launch_gemm_cm<BN, A_MN, B_MN, 2, Epi>(A, lda, B, ldb, s, ep, stream);
  return launch_gemm_cm<BN, A_MN, B_MN, 1, Epi>(A, lda, B, ldb, s, ep, stream);
}

}  // namespace ltg
