// The "middle" of the MultiVAE on the 5th-generation tensor cores: everything between the encoder gather and the decoder GEMM,
// forward and backward, each as ONE kernel that chains two tcgen05 GEMMs through TMEM and shared memory.
//
//   forward  (MultiVAE.py:151-162,178-181,168-172):  [mu|logvar] = h1 W_q1 + b_q1 ; KL ; z = mu + is_training*eps*exp(logvar/2) ;
//                                                     h2 = tanh(z W_p0 + b_p0)
//   backward (autodiff of the same lines, train.py:164):  dz = dh2pre W_p0^T ; dmulv = f(dz, KL) ; dh1 = dmulv W_q1^T ;
//                                                     dh1pre = dh1 (1-h1^2) ; bias gradients
//
// These layers are 0.36 GFLOP at batch 500 and sit on the critical path of every phase. The mma.sync version (mid_kernels.cu: 160
// CTAs, cp.async ring, four launches, 16 + 28 us alone) is bound by L2 latency per k-step and slows down 2.5x whenever an HBM sweep
// runs beside it (timeline of round 2: 34 + 34 us under the decoder Adam). Here a CTA owns 128 batch rows and one third of the
// second GEMM's output columns (grid = ceil(B/128) x 3, so 12 CTAs at batch 500 -- the rest of the GPU stays free for the sweeps):
//
//   warp 0      TMA producer: A tile + weight tile per 64-deep k block through a 2-stage mbarrier ring (128B swizzle)
//   warp 1      MMA issuer:   GEMM 1 into TMEM, then (once the epilogue warps have published the intermediate tile in shared
//               memory, in the K-major 128B-swizzled layout TMA would have produced) GEMM 2 into TMEM
//   warps 2..9  epilogue:     tcgen05.ld (thread == row), latent head / its backward, bf16 intermediate -> shared memory (+ global,
//               by the CTA that owns third 0), final activation / tanh' and the bias-gradient column sums
//
// GEMM 1 is recomputed by the three CTAs of a row block (it is two thirds of 0.7 MFLOP per user: nothing next to the decoder), which
// keeps the kernel free of cluster exchanges. Numerics are those of mid_kernels.cu: bf16 operands, fp32 accumulate, tanhf / expf.
#include "gemm_sm100.cuh"
#include "../../include/ltgan.h"

namespace {
using namespace ltg;

constexpr int H = LTG_H;      // 600
constexpr int L = LTG_L;      // 200
constexpr int MT_EPW = 8;                          // epilogue warps (two per TMEM sub-partition, interleaved chunks)
constexpr int MT_THREADS = 64 + 32 * MT_EPW;
constexpr int NT3 = 3;                             // column thirds of the second GEMM
constexpr int KB_H = (H + 63) / 64;                // 10 k blocks over 600
constexpr int KB_L = (L + 63) / 64;                // 4 k blocks over 200
constexpr int KB_2L = (2 * L + 63) / 64;           // 7 k blocks over 400
constexpr int NQ = 208;                            // UMMA N covering one third (200 columns, multiple of 16)

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-byte piece (8 columns starting at c0, a multiple of 8) of row r of a K-major, 128B-swizzled [128 x 64*kb] bf16 tile
__device__ __forceinline__ void tile_store8(uint8_t* tile, int r, int c0, uint4 u) {
  const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
  *reinterpret_cast<uint4*>(tile + kb * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = u;
}
__device__ __forceinline__ uint4 pack8f(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
__device__ __forceinline__ void ld8f(float (&o)[8], const float* p) {   // 32-byte aligned
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// Column sums of v[0..8) over the 32 rows (lanes) of the warp by halving exchange (7 + 1 shuffles instead of 40):
// afterwards lane l (l % 4 == 0) holds the sum of column (l >> 2) & 7 ... returned in `out`, valid where (lane & 3) == 0.
__device__ __forceinline__ float warp_colsum8(const float (&v)[8], int lane, int& col) {
  float r4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = (lane & 16) != 0;
    const float keepv = up ? v[4 + i] : v[i], send = up ? v[i] : v[4 + i];
    r4[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float r2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = (lane & 8) != 0;
    const float keepv = up ? r4[2 + i] : r4[i], send = up ? r4[i] : r4[2 + i];
    r2[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float r1;
  {
    const bool up = (lane & 4) != 0;
    const float keepv = up ? r2[1] : r2[0], send = up ? r2[0] : r2[1];
    r1 = keepv + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
  r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
  col = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  return r1;
}

// ================================================================================================================================
// forward
// ================================================================================================================================
constexpr int F_A_BYTES = 16384;                   // h1 k block: 128 rows x 128 B
constexpr int F_B_BYTES = 7 * 8192;                // W_q1 k block, MN-major: seven 64(n) x 64(k) boxes (400 -> 448 columns)
constexpr int F_STAGE = F_A_BYTES + F_B_BYTES;     // 72 KB
constexpr int F_RING = 2 * F_STAGE;                // 144 KB; afterwards holds the W_p0 third: 4 k blocks x four 64 x 64 boxes = 128 KB
constexpr int F_Z = KB_L * 16384;                  // z tile, 128 x 256, K-major
constexpr size_t F_SMEM = 1024 + F_RING + F_Z + 256;

struct MidFwdParams {
  int B; const float* bq1; const float* bp0; const float* eps; int64_t uid0; float is_training; uint64_t seed; uint32_t step;
  const uint32_t* step_dev; float* mulv; __nv_bfloat16* z; int ld_z; float* zmu; __nv_bfloat16* h2; int ld_h2; float* scal;
};

__global__ void __launch_bounds__(MT_THREADS, 1)
mid_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmWq1, const __grid_constant__ CUtensorMap tmWp0,
                  const __grid_constant__ MidFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* zt = smem + F_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zt + F_Z);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* bar_acc1 = bars + 4;    // GEMM 1 retired
  uint64_t* bar_w2 = bars + 5;      // W_p0 third landed
  uint64_t* bar_z = bars + 6;       // z tile published (MT_EPW arrivals)
  uint64_t* bar_acc2 = bars + 7;    // GEMM 2 retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / NT3) * GEMM_BM;
  const int third = blockIdx.x % NT3;
  const int n0 = third * L;          // this CTA's columns of h2: [n0, n0 + 200)
  const bool writer = third == 0;    // the intermediates (mulv, z, zmu, KL) are identical in the three CTAs: one of them stores

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH1); tma_prefetch_desc(&tmWq1); tma_prefetch_desc(&tmWp0);
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(bar_acc1, 1); mbar_init(bar_w2, 1); mbar_init(bar_z, MT_EPW); mbar_init(bar_acc2, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // (both single-thread roles run their loops with the whole warp and gate the issuing instructions with elect_one(): operands stay
  // warp-uniform, see gemm_sm100.cuh)
  if (warp == 0) {
    {
      for (int kb = 0; kb < KB_H; ++kb) {
        const int st = kb & 1;
        if (kb >= 2) mbar_wait(&empty[st], ((kb >> 1) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&full[st], F_STAGE);
          uint8_t* sa = ring + st * F_STAGE;
          uint8_t* sb = sa + F_A_BYTES;
          tma_load_2d(sa, &tmH1, &full[st], kb * 64, m0);
#pragma unroll
          for (int j = 0; j < 7; ++j) tma_load_2d(sb + j * 8192, &tmWq1, &full[st], j * 64, kb * 64);
        }
        __syncwarp();
      }
      mbar_wait(bar_acc1, 0);            // GEMM 1 has read the whole ring: it now receives this CTA's third of W_p0
      if (elect_one()) {
        mbar_expect_tx(bar_w2, KB_L * 4 * 8192);
        for (int kb = 0; kb < KB_L; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(ring + kb * 32768 + j * 8192, &tmWp0, bar_w2, n0 + j * 64, kb * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t ID_A = umma_idesc(GEMM_BM, 256, false, true), ID_B = umma_idesc(GEMM_BM, 144, false, true);
      constexpr uint32_t ID_2 = umma_idesc(GEMM_BM, NQ, false, true);
      for (int kb = 0; kb < KB_H; ++kb) {
        const int st = kb & 1;
        mbar_wait(&full[st], (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + st * F_STAGE), sb = sa + F_A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_k(sa + k * 32);
            umma_bf16(tmem_base, da, umma_desc_mn(sb + k * 2048, 8192), ID_A, (kb | k) ? 1u : 0u);                    // columns 0..255
            umma_bf16(tmem_base + 256, da, umma_desc_mn(sb + 4 * 8192 + k * 2048, 8192), ID_B, (kb | k) ? 1u : 0u);   // columns 256..399
          }
          umma_commit(&empty[st]);
          if (kb == KB_H - 1) umma_commit(bar_acc1);
        }
        __syncwarp();
      }
      mbar_wait(bar_z, 0);               // z tile complete in shared memory, TMEM columns 0..399 drained
      mbar_wait(bar_w2, 0);
      tc_fence_after();
      const uint32_t zs = smem_u32(zt), ws = smem_u32(ring);
      if (elect_one()) {
        for (int ks = 0; ks < (L + 15) / 16; ++ks) {   // 13 k steps of 16 (columns >= 200 of the z tile are zero)
          const int kb = ks >> 2, k = ks & 3;
          umma_bf16(tmem_base, umma_desc_k(zs + kb * 16384 + k * 32), umma_desc_mn(ws + kb * 32768 + k * 2048, 8192), ID_2, ks ? 1u : 0u);
        }
        umma_commit(bar_acc2);
      }
      __syncwarp();
    }
  } else {
    const int sub = warp & 3, half = (warp - 2) >> 2;
    const int rl = sub * 32 + lane, row = m0 + rl;
    const bool row_ok = row < p.B;
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const uint32_t step = p.step + (p.step_dev != nullptr ? *p.step_dev : 0u);
    // ---- latent head
    mbar_wait(bar_acc1, 0);
    tc_fence_after();
    float kl = 0.f;
    for (int c = half * 8; c < L; c += 16) {
      float mu[8], lv[8], bm[8], bl[8];
      tmem_ld8(taddr + c, mu);
      tmem_ld8(taddr + L + c, lv);
      ld8f(bm, p.bq1 + c); ld8f(bl, p.bq1 + L + c);
      float zz[8], dd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mu[i] += bm[i]; lv[i] += bl[i];
        float e = 0.f;
        if (p.is_training != 0.f && row_ok) {
          if (p.eps != nullptr) {
            e = p.eps[(size_t)row * L + c + i];
          } else {
            const uint64_t gi = (uint64_t)(p.uid0 + row) * (uint64_t)L + (uint64_t)(c + i);
            Philox4 rr = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), LTG_STREAM_EPS, step, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
            e = sqrtf(-2.0f * logf(ltg_u01(rr.x))) * cospif(2.0f * ltg_u01(rr.y));
          }
        }
        dd[i] = p.is_training * e * expf(0.5f * lv[i]);
        zz[i] = row_ok ? mu[i] + dd[i] : 0.f;
        if (row_ok) kl += 0.5f * (-lv[i] + expf(lv[i]) + mu[i] * mu[i] - 1.0f);
      }
      const uint4 zu = pack8f(zz);
      tile_store8(zt, rl, c, zu);
      if (writer && row_ok) {
        st8f(p.mulv + (size_t)row * 2 * L + c, mu);
        st8f(p.mulv + (size_t)row * 2 * L + L + c, lv);
        st8f(p.zmu + (size_t)row * L + c, dd);
        *reinterpret_cast<uint4*>(p.z + (size_t)row * p.ld_z + c) = zu;
      }
    }
    if (half == 0)
      for (int c = L; c < KB_L * 64; c += 8) tile_store8(zt, rl, c, make_uint4(0, 0, 0, 0));   // K padding of the z tile
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_z);
    if (writer) {
      kl = warp_sum(kl);
      if (lane == 0) atomicAdd(p.scal + LTG_S_KL_SUM, kl);
    }
    // ---- h2 = tanh(z W_p0 + b)
    mbar_wait(bar_acc2, 0);
    tc_fence_after();
    for (int c = half * 16; c < L; c += 32) {
      float v[16];
      tmem_ld16(taddr + c, v);
      if (row_ok) {
        __nv_bfloat16* o = p.h2 + (size_t)row * p.ld_h2 + n0 + c;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (c + 8 * hh >= L) break;        // the last chunk holds 8 valid columns
          float b[8], t[8];
          ld8f(b, p.bp0 + n0 + c + 8 * hh);
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = tanhf(v[8 * hh + i] + b[i]);
          *reinterpret_cast<uint4*>(o + 8 * hh) = pack8f(t);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================================================
// backward
// ================================================================================================================================
constexpr int B_A_BYTES = 16384;                   // dh2pre k block
constexpr int B_B_BYTES = NQ * 128;                // weight k block, K-major: 208 rows x 128 B (26 KB)
constexpr int B_STAGE = B_A_BYTES + B_B_BYTES;     // 42 KB
constexpr int B_RING = 2 * B_STAGE;
constexpr int B_D = KB_2L * 16384;                 // dmulv tile, 128 x 448, K-major
constexpr size_t B_SMEM = 1024 + B_RING + B_D + 256;
constexpr int ACC2 = 256;                          // TMEM column of the second accumulator

struct MidBwdParams {
  int B; float inv_bg; float anneal; const float* scal; const float* mulv; const float* zmu; const __nv_bfloat16* h1; int ld_h1;
  __nv_bfloat16* dmulv; float* dh1pre; __nv_bfloat16* dh1pre_b; float* db_q1; float* db_q0;
};

__global__ void __launch_bounds__(MT_THREADS, 1)
mid_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmWp0, const __grid_constant__ CUtensorMap tmWq1,
                  const __grid_constant__ MidBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* dt = smem + B_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dt + B_D);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* bar_acc1 = bars + 4;
  uint64_t* bar_d = bars + 5;       // dmulv tile published (MT_EPW arrivals)
  uint64_t* bar_acc2 = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / NT3) * GEMM_BM;
  const int third = blockIdx.x % NT3;
  const int n0 = third * L;          // this CTA's columns of dh1: [n0, n0 + 200)
  const bool writer = third == 0;
  constexpr int KB_ALL = KB_H + KB_2L;   // one ring for both GEMMs: k blocks 0..9 carry dh2pre + W_p0, 10..16 carry W_q1 rows only

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmD2); tma_prefetch_desc(&tmWp0); tma_prefetch_desc(&tmWq1);
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(bar_acc1, 1); mbar_init(bar_d, MT_EPW); mbar_init(bar_acc2, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      for (int kb = 0; kb < KB_ALL; ++kb) {
        const int st = kb & 1;
        if (kb >= 2) mbar_wait(&empty[st], ((kb >> 1) - 1) & 1);
        if (elect_one()) {
          uint8_t* sa = ring + st * B_STAGE;
          uint8_t* sb = sa + B_A_BYTES;
          if (kb < KB_H) {
            mbar_expect_tx(&full[st], B_STAGE);
            tma_load_2d(sa, &tmD2, &full[st], kb * 64, m0);
            tma_load_2d(sb, &tmWp0, &full[st], kb * 64, 0);
          } else {
            mbar_expect_tx(&full[st], B_B_BYTES);
            tma_load_2d(sb, &tmWq1, &full[st], (kb - KB_H) * 64, n0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t ID = umma_idesc(GEMM_BM, NQ, false, false);
      const uint32_t ds = smem_u32(dt);
      for (int kb = 0; kb < KB_ALL; ++kb) {
        const int st = kb & 1;
        if (kb == KB_H) mbar_wait(bar_d, 0);          // dmulv tile complete in shared memory
        mbar_wait(&full[st], (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + st * B_STAGE), sb = sa + B_A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kb < KB_H) umma_bf16(tmem_base, umma_desc_k(sa + k * 32), umma_desc_k(sb + k * 32), ID, (kb | k) ? 1u : 0u);
            else umma_bf16(tmem_base + ACC2, umma_desc_k(ds + (kb - KB_H) * 16384 + k * 32), umma_desc_k(sb + k * 32), ID, ((kb - KB_H) | k) ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          if (kb == KB_H - 1) umma_commit(bar_acc1);
          if (kb == KB_ALL - 1) umma_commit(bar_acc2);
        }
        __syncwarp();
      }
    }
  } else {
    const int sub = warp & 3, half = (warp - 2) >> 2;
    const int rl = sub * 32 + lane, row = m0 + rl;
    const bool row_ok = row < p.B;
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const float anneal = p.anneal < 0.f ? p.scal[LTG_S_ANNEAL] : p.anneal;
    // ---- latent backward: dmu = dz + anneal*mu/Bg ; dlogvar = dz*zmu/2 + anneal*(exp(logvar)-1)/(2 Bg)
    mbar_wait(bar_acc1, 0);
    tc_fence_after();
    for (int c = half * 8; c < L; c += 16) {
      float gz[8], dmu[8], dlv[8];
      tmem_ld8(taddr + c, gz);
      if (row_ok) {
        float mu[8], lv[8], zm[8];
        ld8f(mu, p.mulv + (size_t)row * 2 * L + c); ld8f(lv, p.mulv + (size_t)row * 2 * L + L + c); ld8f(zm, p.zmu + (size_t)row * L + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dmu[i] = gz[i] + anneal * mu[i] * p.inv_bg;
          dlv[i] = gz[i] * zm[i] * 0.5f + anneal * 0.5f * (expf(lv[i]) - 1.0f) * p.inv_bg;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dmu[i] = 0.f; dlv[i] = 0.f; }
      }
      const uint4 um = pack8f(dmu), ul = pack8f(dlv);
      tile_store8(dt, rl, c, um);
      tile_store8(dt, rl, L + c, ul);
      if (writer) {
        if (row_ok) {
          *reinterpret_cast<uint4*>(p.dmulv + (size_t)row * 2 * L + c) = um;
          *reinterpret_cast<uint4*>(p.dmulv + (size_t)row * 2 * L + L + c) = ul;
        }
        int col;
        const float sm = warp_colsum8(dmu, lane, col);
        const float sl = warp_colsum8(dlv, lane, col);
        if ((lane & 3) == 0) { atomicAdd(p.db_q1 + c + col, sm); atomicAdd(p.db_q1 + L + c + col, sl); }
      }
    }
    if (half == 0)
      for (int c = 2 * L; c < KB_2L * 64; c += 8) tile_store8(dt, rl, c, make_uint4(0, 0, 0, 0));   // K padding
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_d);
    // ---- dh1pre = (dmulv W_q1^T) (1 - h1^2), db_q0
    mbar_wait(bar_acc2, 0);
    tc_fence_after();
    for (int c = half * 16; c < L; c += 32) {
      float v[16];
      tmem_ld16(taddr + ACC2 + c, v);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (c + 8 * hh >= L) break;        // warp-uniform: the last chunk holds 8 valid columns
        const int col0 = n0 + c + 8 * hh;
        float o[8];
        if (row_ok) {
          const uint4 hu = *reinterpret_cast<const uint4*>(p.h1 + (size_t)row * p.ld_h1 + col0);
          const uint32_t hw[4] = {hu.x, hu.y, hu.z, hu.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 hv = unpack_bf16x2(hw[i]);
            o[2 * i] = v[8 * hh + 2 * i] * (1.0f - hv.x * hv.x);
            o[2 * i + 1] = v[8 * hh + 2 * i + 1] * (1.0f - hv.y * hv.y);
          }
          st8f(p.dh1pre + (size_t)row * H + col0, o);
          *reinterpret_cast<uint4*>(p.dh1pre_b + (size_t)row * H + col0) = pack8f(o);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = 0.f;
        }
        int col;
        const float s = warp_colsum8(o, lane, col);
        if ((lane & 3) == 0) atomicAdd(p.db_q0 + col0 + col, s);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int opt_in(const void* kern, size_t bytes, bool* done) {
  if (*done) return LTG_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
  *done = true;
  return LTG_OK;
}

}  // namespace

extern "C" int ltg_vae_mid_fwd_tc(const void* h1_bf16, int ld_h1, const void* Wq1_bf16, const float* b_q1, const void* Wp0_bf16,
                                  const float* b_p0, const float* eps, int B, int64_t uid0, float is_training, uint64_t seed, uint32_t step,
                                  const uint32_t* step_dev, float* mulv, void* z_bf16, int ld_z, float* zmu, void* h2_bf16, int ld_h2,
                                  float* scal, void* stream) {
  LTG_REQUIRE(h1_bf16 && Wq1_bf16 && b_q1 && Wp0_bf16 && b_p0 && mulv && z_bf16 && zmu && h2_bf16 && scal);
  LTG_REQUIRE(ld_h1 % 8 == 0 && ld_h1 >= H && ld_h2 % 8 == 0 && ld_h2 >= H && ld_z % 8 == 0 && ld_z >= L);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(b_q1) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_p0) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(mulv) & 15) == 0 && (reinterpret_cast<uintptr_t>(zmu) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(z_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(h2_bf16) & 15) == 0);
  if (B <= 0) return LTG_OK;
  CUtensorMap tmH1, tmWq1, tmWp0;
  int rc;
  if ((rc = make_tmap_bf16(&tmH1, h1_bf16, H, (uint64_t)B, (uint64_t)ld_h1, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmWq1, Wq1_bf16, 2 * L, H, 2 * L, 64, 64))) return rc;      // stored [600][400]: MN-major B of GEMM 1
  if ((rc = make_tmap_bf16(&tmWp0, Wp0_bf16, H, L, H, 64, 64))) return rc;              // stored [200][600]: MN-major B of GEMM 2
  MidFwdParams p;
  p.B = B; p.bq1 = b_q1; p.bp0 = b_p0; p.eps = eps; p.uid0 = uid0; p.is_training = is_training; p.seed = seed; p.step = step;
  p.step_dev = step_dev; p.mulv = mulv; p.z = reinterpret_cast<__nv_bfloat16*>(z_bf16); p.ld_z = ld_z; p.zmu = zmu;
  p.h2 = reinterpret_cast<__nv_bfloat16*>(h2_bf16); p.ld_h2 = ld_h2; p.scal = scal;
  static bool opted = false;
  if ((rc = opt_in(reinterpret_cast<const void*>(mid_fwd_tc_kernel), F_SMEM, &opted))) return rc;
  const int grid = ((B + GEMM_BM - 1) / GEMM_BM) * NT3;
  mid_fwd_tc_kernel<<<grid, MT_THREADS, F_SMEM, (cudaStream_t)stream>>>(tmH1, tmWq1, tmWp0, p);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_vae_mid_bwd_tc(const void* dh2pre_bf16, const void* Wp0_bf16, const void* Wq1_bf16, const float* mulv, const float* zmu,
                                  const void* h1_bf16, int ld_h1, int B, int B_global, float anneal, const float* scal, void* dmulv_bf16,
                                  float* dh1pre, void* dh1pre_bf16, float* db_q1, float* db_q0, void* stream) {
  LTG_REQUIRE(dh2pre_bf16 && Wp0_bf16 && Wq1_bf16 && mulv && zmu && h1_bf16 && dmulv_bf16 && dh1pre && dh1pre_bf16);
  LTG_REQUIRE(db_q1 && db_q0 && (anneal >= 0.f || scal != nullptr));
  LTG_REQUIRE(ld_h1 % 8 == 0 && ld_h1 >= H);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(mulv) & 15) == 0 && (reinterpret_cast<uintptr_t>(zmu) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(h1_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dmulv_bf16) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(dh1pre) & 15) == 0 && (reinterpret_cast<uintptr_t>(dh1pre_bf16) & 15) == 0);
  if (B <= 0) return LTG_OK;
  CUtensorMap tmD2, tmWp0, tmWq1;
  int rc;
  if ((rc = make_tmap_bf16(&tmD2, dh2pre_bf16, H, (uint64_t)B, H, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmWp0, Wp0_bf16, H, L, H, 64, NQ))) return rc;              // stored [200][600]: K-major B of dz = dh2pre W_p0^T
  if ((rc = make_tmap_bf16(&tmWq1, Wq1_bf16, 2 * L, H, 2 * L, 64, NQ))) return rc;      // stored [600][400]: K-major B of dh1 = dmulv W_q1^T
  MidBwdParams p;
  p.B = B; p.inv_bg = 1.0f / (float)B_global; p.anneal = anneal; p.scal = scal; p.mulv = mulv; p.zmu = zmu;
  p.h1 = reinterpret_cast<const __nv_bfloat16*>(h1_bf16); p.ld_h1 = ld_h1; p.dmulv = reinterpret_cast<__nv_bfloat16*>(dmulv_bf16);
  p.dh1pre = dh1pre; p.dh1pre_b = reinterpret_cast<__nv_bfloat16*>(dh1pre_bf16); p.db_q1 = db_q1; p.db_q0 = db_q0;
  static bool opted = false;
  if ((rc = opt_in(reinterpret_cast<const void*>(mid_bwd_tc_kernel), B_SMEM, &opted))) return rc;
  const int grid = ((B + GEMM_BM - 1) / GEMM_BM) * NT3;
  mid_bwd_tc_kernel<<<grid, MT_THREADS, B_SMEM, (cudaStream_t)stream>>>(tmD2, tmWp0, tmWq1, p);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
