// Discriminator (discriminator.py:16-55) as ONE persistent tcgen05 kernel: per 128 pairs the two branch layers, the concatenated
// fc1 layer and the sigmoid head are chained through TMEM and shared memory and, for the D update (train.py:300), the backward pass
// down to the hidden activation follows in the same tile: dz3 (head backward), dz12 = (dz3 W3^T) * d/da dropout(tanh(a)) as a
// fourth MMA on the resident dz3 tile. What leaves the CTA is what the weight-gradient GEMMs contract over the pairs: Hd, dz3, dz12.
//
// Round 1 had the forward only (one CTA per tile, 2 waves of 148) followed by a separate 36 us GEMM for dz12; the timeline of round 2
// showed the tile as a serial latency chain: operands of branch 2 were requested only after MMA 1 had retired, every CTA paid its own
// prologue, and the backward GEMM re-read dz3 / Hd from HBM. Now:
//   * one persistent CTA per SM loops over tiles (barrier phases carry over), TMEM / barriers are set up once;
//   * shared memory is two regions that change roles inside a tile, so both branches' operands are requested at tile start:
//       R2 (112 KB): X_pop + W1  ->  hidden activation Hd (A operand of MMA 3, written by the epilogue warps in the 128B-swizzled
//                    K-major layout TMA would have produced)  ->  ring for the W3^T tiles of MMA 4
//       R1 ( 96 KB): X_niche + W2  ->  ring for the W3 tiles of MMA 3  ->  dz3 tile (A operand of MMA 4)
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue (tcgen05.ld: thread == row; tanh + counter-hash dropout,
//     head, loss, dz3 / dw4 / db4, dact).
//
//   MMA 1  [128 x 192] = Xp W1            -> TMEM cols   0..191        MMA 3  [128 x 304] = Hd W3      -> cols 0..303
//   MMA 2  [128 x 256] = Xn W2            -> TMEM cols 192..447        MMA 4  [128 x 416] = dz3 W3^T   -> cols 0..415
//
// Dropout masks are the same counter hash, streams and indices as the unfused GEMM epilogues (EpiStore), so both paths produce
// identical activations.
#include "gemm_sm100.cuh"
#include "../../include/ltgan.h"

namespace {
using namespace ltg;

constexpr int DF_N1 = 192;                  // UMMA N of branch 1 (ld1 <= 192)
constexpr int DF_N2 = 256;                  // UMMA N of branch 2 (k3 - off2 <= 256)
constexpr int DF_N3A = 256, DF_N3B = 48;    // fc1: N = 304 as two instructions (ld3 <= 304)
constexpr int DF_MAXH3 = DF_N3A + DF_N3B;
constexpr int DF_N4 = 208;                  // dz12: two instructions of N = 208 (k3 <= 416)
constexpr int DF_KB3_MAX = 7;               // k3 <= 448 (64-k blocks of Hd)
constexpr int DF_KB4 = (DF_MAXH3 + 63) / 64;   // 5 k blocks over ld3 (dz3 tile)
constexpr int DF_R1 = 96 * 1024;            // region 1
constexpr int DF_R2 = DF_KB3_MAX * 16384;   // region 2 (112 KB)
constexpr int DF_W3_STAGE = 5 * 8192;       // one 64-k block of W3 (MN-major): five 64-column boxes
constexpr int DF_W3T_HALF = DF_N4 * 128;    // one 64-k' block of W3^T rows [0,208) or [208,416), K-major: 26 KB
constexpr int DF_W3T_STAGE = 2 * DF_W3T_HALF;
constexpr int DF_W4_PAD = (DF_MAXH3 + 15) / 16 * 16 + 16;   // head weights staged in shared memory (zero beyond ld3)
constexpr size_t DF_SMEM = 1024 + DF_R1 + DF_R2 + 256 + 4 * 128 * 4 + DF_MAXH3 * 4 + 32 + DF_W4_PAD * 4;
constexpr int DF_CH3 = (DF_MAXH3 + 63) / 64;   // fc1 chunks per epilogue warp (5)
static_assert(DF_W3_STAGE * 2 <= DF_R1 && DF_W3T_STAGE * 2 <= DF_R2 && DF_KB4 * 16384 <= DF_R1, "region sizes");

struct DiscFusedParams {
  int P, ld1, ld2, ld3, off2, one3, h2, k3, kb3;
  int rng_row0;                // row index of pair 0 in the dropout counter space (a launch over a slice of a larger pair batch)
  const float* w4; const float* b4; const int32_t* label;
  float keep; uint64_t seed; uint32_t rng_stream, rng_step; const uint32_t* rng_step_dev;
  __nv_bfloat16* Hd; float* y; float* scal; __nv_bfloat16* dz3; float* dw4; float* db4; __nv_bfloat16* dz12;
  unsigned long long* trace;   // debug: per-CTA phase timestamps (tools/disc_trace.py), NULL in production
};

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// slot layout: trace[(cta * 2 + tile_iter) * 32 + k], first two tile iterations of every CTA
#define DF_STAMP(k) do { if (p.trace != nullptr && it < 2) p.trace[((size_t)blockIdx.x * 2 + it) * 32 + (k)] = gtimer(); } while (0)
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// v <- dropout(v) with one 64-bit hash per four adjacent columns (same indices and bits as EpiStore::chunk)
__device__ __forceinline__ void drop16(float (&v)[16], uint32_t key, uint32_t thr16, float inv_keep, int row, int rng_ld, int col0) {
  const uint64_t gbase = ((uint64_t)row * (uint64_t)rng_ld + (uint64_t)col0) >> 2;
  const uint32_t thr_hi = thr16 << 16;   // thr16 <= 65535 whenever dropout is on
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint2 h = ltg_hash_quad(key, gbase + q);
    v[4 * q] = (h.x & 0xFFFFu) < thr16 ? v[4 * q] * inv_keep : 0.f;
    v[4 * q + 1] = h.x < thr_hi ? v[4 * q + 1] * inv_keep : 0.f;
    v[4 * q + 2] = (h.y & 0xFFFFu) < thr16 ? v[4 * q + 2] * inv_keep : 0.f;
    v[4 * q + 3] = h.y < thr_hi ? v[4 * q + 3] * inv_keep : 0.f;
  }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// w4[c .. c+16) from the zero-padded shared-memory copy; c is a multiple of 16 (warp-uniform address: broadcast reads)
__device__ __forceinline__ void load_w16(float (&w)[16], const float* s_w4, int c) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(s_w4 + c)[i];
    w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
  }
}

// 16-byte piece (8 columns starting at g0, a multiple of 8) of row r of a K-major, 128B-swizzled activation tile
__device__ __forceinline__ void tile_store(uint8_t* tile, int r, int g0, uint4 u) {
  const int kb = g0 >> 6, ch = (g0 & 63) >> 3;
  *reinterpret_cast<uint4*>(tile + kb * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = u;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
disc_fused_kernel(const __grid_constant__ CUtensorMap tmXp, const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW3T,
                  const __grid_constant__ CUtensorMap tmHd, const __grid_constant__ DiscFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* r1 = smem;                        // [96 KB]
  uint8_t* r2 = smem + DF_R1;                // [112 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(r2 + DF_R2);
  uint64_t* bar_ld = bars;                   // [2] operands of MMA 1 (R2) / MMA 2 (R1) landed
  uint64_t* bar_mma = bars + 2;              // [4] MMA 1 / 2 / 3 / 4 retired
  uint64_t* bar_hd = bars + 6;               // hidden activation published by the 16 epilogue warps
  uint64_t* bar_dz3 = bars + 7;              // dz3 tile published
  uint64_t* bar_done = bars + 8;             // the epilogue warps have drained TMEM for this tile
  uint64_t* full3 = bars + 9;                // [2] W3 ring (R1)
  uint64_t* empty3 = bars + 11;              // [2]
  uint64_t* full4 = bars + 13;               // [2] W3^T ring (R2)
  uint64_t* empty4 = bars + 15;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* s_part = reinterpret_cast<float*>(r2 + DF_R2 + 256);   // [4][128] per-quarter partial dot products
  float* s_dw4 = s_part + 4 * 128;                               // [DF_MAXH3]
  float* s_acc = s_dw4 + DF_MAXH3;                               // loss, sum_y, sum ds, n_generated
  float* s_w4 = s_acc + 8;                                       // [DF_W4_PAD] head weights (16-byte aligned: DF_MAXH3 % 4 == 0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool bwd = p.dz3 != nullptr;
  const bool bwd4 = bwd && p.dz12 != nullptr;   // MMA 4 / dz12 in this kernel
  const int n_tiles = (p.P + GEMM_BM - 1) / GEMM_BM;

  pdl_trigger();   // (PDL, ltg_common.cuh) barrier init / TMEM allocation overlap the previous kernel's tail
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmXp); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmXn); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW3);
    if (bwd4) tma_prefetch_desc(&tmW3T); else tma_prefetch_desc(&tmHd);
    mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_mma[i], 1);
    mbar_init(bar_hd, GEMM_EPI_WARPS); mbar_init(bar_dz3, GEMM_EPI_WARPS); mbar_init(bar_done, GEMM_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&full3[i], 1); mbar_init(&empty3[i], 1); mbar_init(&full4[i], 1); mbar_init(&empty4[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int j = threadIdx.x; j < DF_MAXH3 + 4; j += GEMM_THREADS) s_dw4[j] = 0.f;   // s_dw4 and s_acc are contiguous
  pdl_wait_cta();      // the head weights below are the first global data this kernel reads
  // The head weights are read by every epilogue thread for every chunk, twice per tile in the D update: from global memory that was one
  // L1/L2 round trip per chunk in front of the FMAs (ncu source view, round 2: the top long-scoreboard lines of the kernel)
  for (int j = threadIdx.x; j < DF_W4_PAD; j += GEMM_THREADS) s_w4[j] = j < p.ld3 ? __ldg(p.w4 + j) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in the loop, one elected lane issues: see elect_one) =====================
    {
      int st3 = 0; uint32_t ph3 = 0; int st4 = 0; uint32_t ph4 = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int m0 = t * GEMM_BM;
        const uint32_t par = it & 1, prev = par ^ 1;
        // both regions are free once the last MMA of the previous tile that reads them has retired
        DF_STAMP(0);
        if (it > 0) mbar_wait(&bar_mma[bwd4 ? 3 : 2], prev);
        DF_STAMP(1);
        if (elect_one()) {
          {  // R2: X_pop [2 x 16 KB] + W1 [2 k blocks x three 64-column boxes]
            uint8_t* xs = r2; uint8_t* ws = r2 + 32768;
            mbar_expect_tx(&bar_ld[0], 32768 + 2 * 3 * 8192);
            for (int kb = 0; kb < 2; ++kb) {
              tma_load_2d(xs + kb * 16384, &tmXp, &bar_ld[0], kb * 64, m0);
              for (int j = 0; j < 3; ++j) tma_load_2d(ws + kb * (3 * 8192) + j * 8192, &tmW1, &bar_ld[0], j * 64, kb * 64);
            }
          }
          {  // R1: X_niche + W2 [2 k blocks x four boxes]
            uint8_t* xs = r1; uint8_t* ws = r1 + 32768;
            mbar_expect_tx(&bar_ld[1], 32768 + 2 * 4 * 8192);
            for (int kb = 0; kb < 2; ++kb) {
              tma_load_2d(xs + kb * 16384, &tmXn, &bar_ld[1], kb * 64, m0);
              for (int j = 0; j < 4; ++j) tma_load_2d(ws + kb * (4 * 8192) + j * 8192, &tmW2, &bar_ld[1], j * 64, kb * 64);
            }
          }
        }
        __syncwarp();
        mbar_wait(&bar_mma[1], par);             // MMA 2 has read R1: it becomes the W3 ring
        DF_STAMP(2);
        for (int kb = 0; kb < p.kb3; ++kb) {
          mbar_wait(&empty3[st3], ph3 ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full3[st3], DF_W3_STAGE);
            for (int j = 0; j < 5; ++j) tma_load_2d(r1 + st3 * DF_W3_STAGE + j * 8192, &tmW3, &full3[st3], j * 64, kb * 64);
          }
          __syncwarp();
          if (++st3 == 2) { st3 = 0; ph3 ^= 1; }
        }
        if (bwd4) {
          DF_STAMP(3);
          mbar_wait(&bar_mma[2], par);           // MMA 3 has read Hd in R2: it becomes the W3^T ring
          DF_STAMP(4);
          for (int kb = 0; kb < DF_KB4; ++kb) {
            mbar_wait(&empty4[st4], ph4 ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full4[st4], DF_W3T_STAGE);
              tma_load_2d(r2 + st4 * DF_W3T_STAGE, &tmW3T, &full4[st4], kb * 64, 0);
              tma_load_2d(r2 + st4 * DF_W3T_STAGE + DF_W3T_HALF, &tmW3T, &full4[st4], kb * 64, DF_N4);
            }
            __syncwarp();
            if (++st4 == 2) { st4 = 0; ph4 ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
    {
      constexpr uint32_t ID1 = umma_idesc(GEMM_BM, DF_N1, false, true), ID2 = umma_idesc(GEMM_BM, DF_N2, false, true);
      constexpr uint32_t ID3A = umma_idesc(GEMM_BM, DF_N3A, false, true), ID3B = umma_idesc(GEMM_BM, DF_N3B, false, true);
      constexpr uint32_t ID4 = umma_idesc(GEMM_BM, DF_N4, false, false);
      const uint32_t a1 = smem_u32(r1), a2 = smem_u32(r2);
      int st3 = 0; uint32_t ph3 = 0; int st4 = 0; uint32_t ph4 = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t par = it & 1, prev = par ^ 1;
        if (it > 0) { mbar_wait(bar_done, prev); tc_fence_after(); }   // the previous tile's accumulators have been read out
        DF_STAMP(8);
        mbar_wait(&bar_ld[0], par);
        DF_STAMP(9);
        tc_fence_after();
        if (elect_one()) {
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, umma_desc_k(a2 + kb * 16384 + k * 32), umma_desc_mn(a2 + 32768 + kb * (3 * 8192) + k * 2048, 8192), ID1,
                        (kb | k) ? 1u : 0u);
          umma_commit(&bar_mma[0]);
        }
        __syncwarp();
        mbar_wait(&bar_ld[1], par);
        DF_STAMP(10);
        tc_fence_after();
        if (elect_one()) {
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + DF_N1, umma_desc_k(a1 + kb * 16384 + k * 32), umma_desc_mn(a1 + 32768 + kb * (4 * 8192) + k * 2048, 8192), ID2,
                        (kb | k) ? 1u : 0u);
          umma_commit(&bar_mma[1]);
        }
        __syncwarp();
        mbar_wait(bar_hd, par);                  // Hd tile complete in R2, TMEM columns 0..447 drained
        DF_STAMP(11);
        tc_fence_after();
        // The hidden activation leaves the CTA as TMA tile stores of the very tile MMA 3 reads (same 128B-swizzled K-major blocks):
        // full 128-byte lines written by the copy engine instead of 16 bytes per thread and row from the epilogue warps (204 store
        // instructions of 32 sectors each per tile). Rows >= P and columns >= k3 are clipped by the tensor map. (With MMA 4 in the kernel
        // the epilogue warps keep their own stores: they read the activation back through the generic proxy.)
        // lane 0 issues the stores and later waits for them (a bulk async-group belongs to the thread that committed it)
        if (!bwd4 && lane == 0) {
          for (int kb = 0; kb < p.kb3; ++kb) tma_store_2d(&tmHd, r2 + kb * 16384, kb * 64, t * GEMM_BM);
          bulk_commit();
        }
        __syncwarp();
        for (int kb = 0; kb < p.kb3; ++kb) {
          mbar_wait(&full3[st3], ph3);
          tc_fence_after();
          const uint32_t w3 = a1 + st3 * DF_W3_STAGE;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_desc_k(a2 + kb * 16384 + k * 32);
              umma_bf16(tmem_base, da, umma_desc_mn(w3 + k * 2048, 8192), ID3A, (kb | k) ? 1u : 0u);
              umma_bf16(tmem_base + DF_N3A, da, umma_desc_mn(w3 + 4 * 8192 + k * 2048, 8192), ID3B, (kb | k) ? 1u : 0u);
            }
            umma_commit(&empty3[st3]);
          }
          __syncwarp();
          if (++st3 == 2) { st3 = 0; ph3 ^= 1; }
        }
        if (!bwd4 && lane == 0) bulk_wait_read0();   // R2 is handed back to the producer by the commit below: the tile stores have read it
        __syncwarp();
        if (elect_one()) umma_commit(&bar_mma[2]);
        __syncwarp();
        DF_STAMP(12);
        if (bwd4) {
          mbar_wait(bar_dz3, par);               // dz3 tile complete in R1, fc1 accumulators drained
          DF_STAMP(13);
          tc_fence_after();
          for (int kb = 0; kb < DF_KB4; ++kb) {
            mbar_wait(&full4[st4], ph4);
            tc_fence_after();
            const uint32_t w = a2 + st4 * DF_W3T_STAGE;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_k(a1 + kb * 16384 + k * 32);
                umma_bf16(tmem_base, da, umma_desc_k(w + k * 32), ID4, (kb | k) ? 1u : 0u);
                umma_bf16(tmem_base + DF_N4, da, umma_desc_k(w + DF_W3T_HALF + k * 32), ID4, (kb | k) ? 1u : 0u);
              }
              umma_commit(&empty4[st4]);
            }
            __syncwarp();
            if (++st4 == 2) { st4 = 0; ph4 ^= 1; }
          }
          if (elect_one()) umma_commit(&bar_mma[3]);
          __syncwarp();
          DF_STAMP(14);
        }
      }
      if (lane == 0) bulk_wait0();               // the last tile's activation stores are complete before the CTA retires
    }
  } else {
    // ===================== epilogue =====================
    const int sub = warp & 3, quarter = (warp - 2) >> 2;
    const int rl = sub * 32 + lane;            // row inside the tile == TMEM lane
    const bool drop = p.keep > 0.f && p.keep < 1.f;
    const uint32_t thr16 = drop ? ltg_keep_threshold16(p.keep) : 65536u;
    const float inv_keep = drop ? 1.0f / p.keep : 1.0f;
    const uint32_t step = p.rng_step + (p.rng_step_dev != nullptr ? *p.rng_step_dev : 0u);
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const uint32_t key1 = drop ? ltg_hash_key(p.seed, p.rng_stream, step) : 0u;
    const uint32_t key2 = drop ? ltg_hash_key(p.seed, p.rng_stream + 1, step) : 0u;
    const uint32_t key3 = drop ? ltg_hash_key(p.seed, p.rng_stream + 2, step) : 0u;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row = t * GEMM_BM + rl;
      const bool row_ok = row < p.P;
      __nv_bfloat16* hrow = p.Hd + (size_t)(row_ok ? row : 0) * p.k3;

      // ---- branch 1: columns [0, off2) of the hidden activation
      if (warp == 2 && lane == 0) DF_STAMP(16);
      mbar_wait(&bar_mma[0], par);
      if (warp == 2 && lane == 0) DF_STAMP(17);
      tc_fence_after();
      for (int c = quarter * 16; c < p.off2; c += 64) {
        float v[16];
        tmem_ld16(taddr + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
        if (drop) drop16(v, key1, thr16, inv_keep, p.rng_row0 + row, p.ld1, c);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g0 = c + 8 * h;
          if (g0 < p.off2) {
            const uint4 u = pack8(v + 8 * h);
            tile_store(r2, rl, g0, u);
            if (bwd4 && row_ok) *reinterpret_cast<uint4*>(hrow + g0) = u;
          }
        }
      }
      // ---- branch 2: columns [off2, k3): h2 activations, the ones column (fc1 bias row of W3), zero padding
      if (warp == 2 && lane == 0) DF_STAMP(18);
      mbar_wait(&bar_mma[1], par);
      if (warp == 2 && lane == 0) DF_STAMP(19);
      tc_fence_after();
      for (int c = quarter * 16; c < p.k3 - p.off2; c += 64) {
        float v[16];
        tmem_ld16(taddr + DF_N1 + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
        if (drop) drop16(v, key2, thr16, inv_keep, p.rng_row0 + row, p.ld2, c);
        if (c + 16 > p.h2) {                     // warp-uniform: only the last chunk(s) hold the ones column / padding
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int g = p.off2 + c + i;
            if (g >= p.off2 + p.h2) v[i] = (g == p.one3) ? 1.0f : 0.f;
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g0 = p.off2 + c + 8 * h;
          if (g0 < p.k3) {
            const uint4 u = pack8(v + 8 * h);
            tile_store(r2, rl, g0, u);
            if (bwd4 && row_ok) *reinterpret_cast<uint4*>(hrow + g0) = u;
          }
        }
      }
      // K padding of the last 64-k block: W3 rows >= k3 are zero-filled by TMA, the A side must be finite
      if (quarter == 0)
        for (int g0 = p.k3; g0 < p.kb3 * 64; g0 += 8) tile_store(r2, rl, g0, make_uint4(0, 0, 0, 0));
      fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hd);
      if (warp == 2 && lane == 0) DF_STAMP(20);

      // ---- fc1 + head
      uint32_t yq[DF_CH3][8];                    // this thread's fc1 activations (bf16 pairs), kept for the backward pass
      float s = 0.f;
      mbar_wait(&bar_mma[2], par);
      if (warp == 2 && lane == 0) DF_STAMP(21);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < DF_CH3; ++j) {
        const int c = quarter * 16 + 64 * j;
        if (c < p.ld3) {
          float v[16];
          tmem_ld16(taddr + c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
          if (drop) drop16(v, key3, thr16, inv_keep, p.rng_row0 + row, p.ld3, c);
          float w[16];
          load_w16(w, s_w4, c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t u = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            yq[j][i] = u;
            const float2 f = unpack_bf16x2(u);   // the head sees the bf16-rounded activation, like the unfused path
            s = fmaf(f.x, w[2 * i], s);
            s = fmaf(f.y, w[2 * i + 1], s);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) yq[j][i] = 0u;
        }
      }
      if (!bwd4) {                               // TMEM is drained for this tile (with MMA 4 that happens after its epilogue)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_done);
      }
      if (warp == 2 && lane == 0) DF_STAMP(22);
      s_part[quarter * 128 + rl] = s;
      named_bar_sync(1, GEMM_EPI_WARPS * 32);
      s = s_part[rl] + s_part[128 + rl] + s_part[256 + rl] + s_part[384 + rl] + __ldg(p.b4);
      named_bar_sync(1, GEMM_EPI_WARPS * 32);    // s_part is rewritten by the next tile
      const float y = 1.0f / (1.0f + __expf(-s));
      const int lab = row_ok ? p.label[row] : -1;
      // -log(sigmoid(s)) = softplus(-s) ; -log(1 - sigmoid(s)) = softplus(s)   (train.py:142); label < 0: pair dropped
      const float sp = (lab == 0) ? -s : s;
      const float ds = lab < 0 ? 0.f : ((lab == 0) ? (y - 1.0f) : y);
      if (quarter == 0) {
        if (row_ok && p.y != nullptr) p.y[row] = y;
        float l = lab < 0 ? 0.f : fmaxf(sp, 0.f) + log1pf(__expf(-fabsf(sp)));
        float sy = lab == 1 ? y : 0.f, ng = lab == 1 ? 1.f : 0.f, sd = ds;
        l = warp_sum(l); sy = warp_sum(sy); ng = warp_sum(ng); sd = warp_sum(sd);
        if (lane == 0) { atomicAdd(&s_acc[0], l); atomicAdd(&s_acc[1], sy); atomicAdd(&s_acc[2], sd); atomicAdd(&s_acc[3], ng); }
      }
      if (warp == 2 && lane == 0) DF_STAMP(23);
      if (bwd) {
        __nv_bfloat16* drow = p.dz3 + (size_t)(row_ok ? row : 0) * p.ld3;
#pragma unroll
        for (int j = 0; j < DF_CH3; ++j) {
          const int c = quarter * 16 + 64 * j;
          if (c < p.ld3) {                       // warp-uniform
            float d[16], gw[16], w[16];
            load_w16(w, s_w4, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = unpack_bf16x2(yq[j][i]);
              const float w0 = w[2 * i], w1 = w[2 * i + 1];
              float a0, a1;
              if (drop) {
                const float t0 = f.x * p.keep, t1 = f.y * p.keep;
                a0 = f.x == 0.f ? 0.f : (1.0f - t0 * t0) * inv_keep;
                a1 = f.y == 0.f ? 0.f : (1.0f - t1 * t1) * inv_keep;
              } else {
                a0 = 1.0f - f.x * f.x; a1 = 1.0f - f.y * f.y;
              }
              d[2 * i] = ds * w0 * a0; d[2 * i + 1] = ds * w1 * a1;
              gw[2 * i] = ds * f.x; gw[2 * i + 1] = ds * f.y;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (c + 8 * h < p.ld3) {
                const uint4 u = pack8(d + 8 * h);
                if (row_ok) *reinterpret_cast<uint4*>(drow + c + 8 * h) = u;
                if (bwd4) tile_store(r1, rl, c + 8 * h, row_ok ? u : make_uint4(0, 0, 0, 0));
              } else if (bwd4) {
                tile_store(r1, rl, c + 8 * h, make_uint4(0, 0, 0, 0));
              }
            }
            // dw4[c + i] += sum over the warp's 32 rows of gw[i]: halving exchange, 16 shuffles instead of 80
            float r8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool up = (lane & 16) != 0;
              const float keepv = up ? gw[8 + i] : gw[i], send = up ? gw[i] : gw[8 + i];
              r8[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float r4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool up = (lane & 8) != 0;
              const float keepv = up ? r8[4 + i] : r8[i], send = up ? r8[i] : r8[4 + i];
              r4[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            float r2v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const bool up = (lane & 4) != 0;
              const float keepv = up ? r4[2 + i] : r4[i], send = up ? r4[i] : r4[2 + i];
              r2v[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            float r1v;
            {
              const bool up = (lane & 2) != 0;
              const float keepv = up ? r2v[1] : r2v[0], send = up ? r2v[0] : r2v[1];
              r1v = keepv + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            r1v += __shfl_xor_sync(0xffffffffu, r1v, 1);
            const int col = c + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0 && col < p.ld3) atomicAdd(&s_dw4[col], r1v);
          } else if (bwd4 && c < DF_KB4 * 64) {  // K padding of the dz3 tile (columns [ld3, 320))
            tile_store(r1, rl, c, make_uint4(0, 0, 0, 0));
            tile_store(r1, rl, c + 8, make_uint4(0, 0, 0, 0));
          }
        }
      }
      if (bwd4) {
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dz3);
        if (warp == 2 && lane == 0) DF_STAMP(24);
        // ---- dz12 = (dz3 W3^T) * d/da dropout(tanh(a)), recovered from the stored activation Hd (L2-hot: this CTA wrote it above)
        const float kk = drop ? p.keep : 1.0f;
        __nv_bfloat16* orow = p.dz12 + (size_t)(row_ok ? row : 0) * p.k3;
        int c = quarter * 16;
        uint4 h0 = make_uint4(0, 0, 0, 0), h1 = h0;
        if (row_ok && c < p.k3) { h0 = *reinterpret_cast<const uint4*>(hrow + c); if (c + 8 < p.k3) h1 = *reinterpret_cast<const uint4*>(hrow + c + 8); }
        mbar_wait(&bar_mma[3], par);
        if (warp == 2 && lane == 0) DF_STAMP(25);
        tc_fence_after();
        for (; c < p.k3; c += 64) {
          const uint4 c0 = h0, c1 = h1;
          const int cn = c + 64;
          if (row_ok && cn < p.k3) {             // next chunk's activation, one chunk ahead of its use
            h0 = *reinterpret_cast<const uint4*>(hrow + cn);
            h1 = (cn + 8 < p.k3) ? *reinterpret_cast<const uint4*>(hrow + cn + 8) : make_uint4(0, 0, 0, 0);
          }
          float v[16];
          tmem_ld16(taddr + c, v);
          const uint32_t hw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 yv = unpack_bf16x2(hw[i]);
            const float t0 = yv.x * kk, t1 = yv.y * kk;
            v[2 * i] *= (drop && yv.x == 0.f) ? 0.f : (1.0f - t0 * t0) * inv_keep;
            v[2 * i + 1] *= (drop && yv.y == 0.f) ? 0.f : (1.0f - t1 * t1) * inv_keep;
          }
          if (row_ok) {
            *reinterpret_cast<uint4*>(orow + c) = pack8(v);
            if (c + 8 < p.k3) *reinterpret_cast<uint4*>(orow + c + 8) = pack8(v + 8);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_done);
        if (warp == 2 && lane == 0) DF_STAMP(26);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) {
    atomicAdd(p.scal + LTG_S_D_LOSS, s_acc[0]);
    atomicAdd(p.scal + LTG_S_SUM_Y, s_acc[1]);
    atomicAdd(p.scal + LTG_S_CNT, s_acc[3]);
    if (bwd && p.db4 != nullptr) atomicAdd(p.db4, s_acc[2]);
  }
  if (bwd && p.dw4 != nullptr)
    for (int j = threadIdx.x; j < p.ld3; j += GEMM_THREADS) atomicAdd(p.dw4 + j, s_dw4[j]);
}

static unsigned long long* g_df_trace = nullptr;
}  // namespace

/* debug only: device buffer of 148 * 2 * 32 u64 that receives per-phase %globaltimer stamps of the next launches (NULL: off) */
extern "C" int ltg_disc_fused_set_trace(void* buf) { g_df_trace = reinterpret_cast<unsigned long long*>(buf); return LTG_OK; }

extern "C" int ltg_disc_fused_supported(int k1, int ld1, int ld2, int ld3, int off2, int one3, int h2, int k3) {
  return (k1 <= 128 && ld1 <= DF_N1 && off2 <= DF_N1 && off2 % 8 == 0 && ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && k3 % 8 == 0 &&
          k3 - off2 <= DF_N2 && ld2 <= DF_N2 && ld3 <= DF_MAXH3 && k3 <= 2 * DF_N4 && k3 <= DF_KB3_MAX * 64 && one3 == off2 + h2 && one3 < k3 &&
          h2 <= ld2)
             ? 1 : 0;
}

extern "C" int ltg_disc_fwd_fused(const void* Xp_bf16, const void* Xn_bf16, int P, int k1, const void* W1_bf16, int ld1, const void* W2_bf16,
                                  int ld2, int h2, const void* W3_bf16, int ld3, int k3, int off2, int one3, const float* w4, const float* b4,
                                  const int32_t* label, float keep, uint64_t seed, uint32_t rng_stream, uint32_t rng_step,
                                  const uint32_t* rng_step_dev, void* Hd_bf16, float* y, float* scal, void* dz3_bf16, float* dw4, float* db4,
                                  void* dz12_bf16, int rng_row0, void* stream) {
  LTG_REQUIRE(Xp_bf16 && Xn_bf16 && W1_bf16 && W2_bf16 && W3_bf16 && w4 && b4 && label && Hd_bf16 && scal);
  LTG_REQUIRE(rng_row0 >= 0);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(w4) & 15) == 0);
  LTG_REQUIRE(dz12_bf16 == nullptr || dz3_bf16 != nullptr);
  LTG_REQUIRE(ltg_disc_fused_supported(k1, ld1, ld2, ld3, off2, one3, h2, k3));
  if (P <= 0) return LTG_OK;
  CUtensorMap tmXp, tmXn, tmW1, tmW2, tmW3, tmW3T, tmHd;
  int rc;
  if ((rc = make_tmap_bf16(&tmXp, Xp_bf16, 128, (uint64_t)P, 128, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmXn, Xn_bf16, 128, (uint64_t)P, 128, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmW1, W1_bf16, (uint64_t)ld1, (uint64_t)k1, (uint64_t)ld1, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmW2, W2_bf16, (uint64_t)ld2, (uint64_t)k1, (uint64_t)ld2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmW3, W3_bf16, (uint64_t)ld3, (uint64_t)k3, (uint64_t)ld3, 64, 64))) return rc;
  // the same W3 [k3, ld3] seen as the K-major B operand of dz12 = dz3 W3^T: rows = output column k, 64-wide slices of n
  if ((rc = make_tmap_bf16(&tmW3T, W3_bf16, (uint64_t)ld3, (uint64_t)k3, (uint64_t)ld3, 64, DF_N4))) return rc;
  // the activation Hd [P, k3] as the destination of the tile stores: the 64-column K blocks of the shared-memory tile
  if ((rc = make_tmap_bf16(&tmHd, Hd_bf16, (uint64_t)k3, (uint64_t)P, (uint64_t)k3, 64, GEMM_BM))) return rc;
  DiscFusedParams p;
  p.P = P; p.ld1 = ld1; p.ld2 = ld2; p.ld3 = ld3; p.off2 = off2; p.one3 = one3; p.h2 = h2; p.k3 = k3; p.kb3 = (k3 + 63) / 64;
  p.w4 = w4; p.b4 = b4; p.label = label; p.keep = keep; p.seed = seed; p.rng_stream = rng_stream; p.rng_step = rng_step;
  p.rng_step_dev = rng_step_dev; p.rng_row0 = rng_row0;
  p.Hd = reinterpret_cast<__nv_bfloat16*>(Hd_bf16); p.y = y; p.scal = scal; p.dz3 = reinterpret_cast<__nv_bfloat16*>(dz3_bf16);
  p.dw4 = dw4; p.db4 = db4; p.dz12 = reinterpret_cast<__nv_bfloat16*>(dz12_bf16); p.trace = g_df_trace;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(disc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM);
    if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
    opted = true;
  }
  const int n_tiles = (P + GEMM_BM - 1) / GEMM_BM;
  const int grid = n_tiles < ltg_num_sms() ? n_tiles : ltg_num_sms();
  ltg_launch(disc_fused_kernel, dim3(grid), dim3(GEMM_THREADS), DF_SMEM, (cudaStream_t)stream, tmXp, tmXn, tmW1, tmW2, tmW3, tmW3T, tmHd, p);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
