// Discriminator forward (discriminator.py:16-55) as ONE tcgen05 kernel per 128 pairs: the two branch layers, the concatenated
// fc1 layer and the sigmoid head chained through TMEM and shared memory, plus (D step) the head's backward.
//
// The unfused chain (3 GEMM launches + head) is bound by per-launch fixed costs, not by math: each launch pays the pipeline
// fill, a 128 x N epilogue pass and a teardown for a K = 128 (or 408) mainloop that takes well under a microsecond, and the head
// re-reads the fc1 activation from HBM. Here a CTA keeps its 128 x 408 hidden activation in shared memory as the A operand of
// the third MMA (written by the epilogue warps in the 128B-swizzled K-major layout TMA would have produced), the fc1
// accumulator never leaves TMEM, and y / loss / dz3 / dw4 / db4 come out of the same epilogue.
//
//   warp 0      TMA producer: X_pop + W1, then X_niche + W2 (same 96 KB buffer, refilled when MMA 1 retires), then W3 streamed
//               through a 2-stage ring of 64-k blocks in that buffer
//   warp 1      MMA issuer:   [128 x 192] += Xp W1 -> TMEM cols 0..191 ; [128 x 256] += Xn W2 -> cols 192..447 ;
//               [128 x 304] += Hd W3 -> cols 0..303 (N = 256 + 48), after the epilogue warps have published Hd
//   warps 2..17 epilogue:     tanh + dropout -> bf16 -> Hd (global, for the backward GEMMs, and shared, for MMA 3);
//               head: row dot with w4 (thread == row; the 4 warps of a TMEM sub-partition take interleaved chunks and combine
//               through shared memory), sigmoid, loss, and dz3 / dw4 / db4 when the backward is requested.
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
constexpr int DF_KB3_MAX = 7;               // k3 <= 448
constexpr int DF_BUF = 96 * 1024;           // operand buffer (phases 1-2: X tile + weights; phase 3: W3 ring)
constexpr int DF_W3_STAGE = 5 * 8192;       // one 64-k block of W3: five 64-column boxes
constexpr int DF_HD = DF_KB3_MAX * 16384;   // hidden activation tile, K-major, 128B swizzle
constexpr size_t DF_SMEM = 1024 + DF_BUF + DF_HD + 256 + 4 * 128 * 4 + DF_MAXH3 * 4 + 32;
constexpr int DF_CH3 = (DF_MAXH3 + 63) / 64;   // fc1 chunks per epilogue warp (5)

struct DiscFusedParams {
  int P, ld1, ld2, ld3, off2, one3, h2, k3, kb3;
  const float* w4; const float* b4; const int32_t* label;
  float keep; uint64_t seed; uint32_t rng_stream, rng_step; const uint32_t* rng_step_dev;
  __nv_bfloat16* Hd; float* y; float* scal; __nv_bfloat16* dz3; float* dw4; float* db4;
};

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

// w4[c .. c+16) (zero beyond n); c is a multiple of 16 and w4 is 16-byte aligned
__device__ __forceinline__ void load_w16(float (&w)[16], const float* __restrict__ w4, int c, int n) {
  if (c + 16 <= n) {   // warp-uniform
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(w4 + c) + i);
      w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = c + i < n ? __ldg(w4 + c + i) : 0.f;
  }
}

// 16-byte piece (8 columns starting at g0, a multiple of 8) of row r of the K-major, 128B-swizzled activation tile
__device__ __forceinline__ void hd_smem_store(uint8_t* hd, int r, int g0, uint4 u) {
  const int kb = g0 >> 6, ch = (g0 & 63) >> 3;
  *reinterpret_cast<uint4*>(hd + kb * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = u;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
disc_fused_kernel(const __grid_constant__ CUtensorMap tmXp, const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ DiscFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* buf = smem;                       // [96 KB]
  uint8_t* hd = smem + DF_BUF;               // [7][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hd + DF_HD);
  uint64_t* bar_ld = bars;                   // [2] operands of MMA 1 / MMA 2 landed
  uint64_t* bar_mma = bars + 2;              // [3] MMA 1 / 2 / 3 retired
  uint64_t* bar_hd = bars + 5;               // hidden activation published by the 16 epilogue warps
  uint64_t* full3 = bars + 6;                // [2] W3 ring
  uint64_t* empty3 = bars + 8;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* s_part = reinterpret_cast<float*>(hd + DF_HD + 256);   // [4][128] per-quarter partial dot products
  float* s_dw4 = s_part + 4 * 128;                               // [DF_MAXH3]
  float* s_acc = s_dw4 + DF_MAXH3;                               // loss, sum_y, sum ds, n_generated

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GEMM_BM;
  const bool bwd = p.dz3 != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmXp); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmXn); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW3);
    mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1);
    mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); mbar_init(&bar_mma[2], 1);
    mbar_init(bar_hd, GEMM_EPI_WARPS);
    mbar_init(&full3[0], 1); mbar_init(&full3[1], 1); mbar_init(&empty3[0], 1); mbar_init(&empty3[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int j = threadIdx.x; j < DF_MAXH3 + 4; j += GEMM_THREADS) s_dw4[j] = 0.f;   // s_dw4 and s_acc are contiguous
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint8_t* xs = buf;
      uint8_t* ws = buf + 32768;
      mbar_expect_tx(&bar_ld[0], 32768 + 2 * 3 * 8192);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(xs + kb * 16384, &tmXp, &bar_ld[0], kb * 64, m0);
        for (int j = 0; j < 3; ++j) tma_load_2d(ws + kb * (3 * 8192) + j * 8192, &tmW1, &bar_ld[0], j * 64, kb * 64);
      }
      mbar_wait(&bar_mma[0], 0);               // MMA 1 has read the buffer
      mbar_expect_tx(&bar_ld[1], 32768 + 2 * 4 * 8192);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(xs + kb * 16384, &tmXn, &bar_ld[1], kb * 64, m0);
        for (int j = 0; j < 4; ++j) tma_load_2d(ws + kb * (4 * 8192) + j * 8192, &tmW2, &bar_ld[1], j * 64, kb * 64);
      }
      mbar_wait(&bar_mma[1], 0);               // MMA 2 has read the buffer: it becomes the W3 ring
      for (int kb = 0; kb < p.kb3; ++kb) {
        const int st = kb & 1;
        if (kb >= 2) mbar_wait(&empty3[st], ((kb >> 1) - 1) & 1);
        mbar_expect_tx(&full3[st], DF_W3_STAGE);
        for (int j = 0; j < 5; ++j) tma_load_2d(buf + st * DF_W3_STAGE + j * 8192, &tmW3, &full3[st], j * 64, kb * 64);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t xs = smem_u32(buf), ws = smem_u32(buf + 32768), hs = smem_u32(hd);
      constexpr uint32_t ID1 = umma_idesc(GEMM_BM, DF_N1, false, true), ID2 = umma_idesc(GEMM_BM, DF_N2, false, true);
      constexpr uint32_t ID3A = umma_idesc(GEMM_BM, DF_N3A, false, true), ID3B = umma_idesc(GEMM_BM, DF_N3B, false, true);
      mbar_wait(&bar_ld[0], 0);
      tc_fence_after();
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_desc_k(xs + kb * 16384 + k * 32), umma_desc_mn(ws + kb * (3 * 8192) + k * 2048, 8192), ID1, (kb | k) ? 1u : 0u);
      umma_commit(&bar_mma[0]);
      mbar_wait(&bar_ld[1], 0);
      tc_fence_after();
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + DF_N1, umma_desc_k(xs + kb * 16384 + k * 32), umma_desc_mn(ws + kb * (4 * 8192) + k * 2048, 8192), ID2,
                    (kb | k) ? 1u : 0u);
      umma_commit(&bar_mma[1]);
      mbar_wait(bar_hd, 0);                    // Hd tile complete in shared memory, TMEM columns 0..447 drained
      tc_fence_after();
      for (int kb = 0; kb < p.kb3; ++kb) {
        const int st = kb & 1;
        mbar_wait(&full3[st], (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t w3 = smem_u32(buf + st * DF_W3_STAGE);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_desc_k(hs + kb * 16384 + k * 32);
          umma_bf16(tmem_base, da, umma_desc_mn(w3 + k * 2048, 8192), ID3A, (kb | k) ? 1u : 0u);
          umma_bf16(tmem_base + DF_N3A, da, umma_desc_mn(w3 + 4 * 8192 + k * 2048, 8192), ID3B, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty3[st]);
      }
      umma_commit(&bar_mma[2]);
    }
  } else {
    // ===================== epilogue =====================
    const int sub = warp & 3, quarter = (warp - 2) >> 2;
    const int rl = sub * 32 + lane;            // row inside the tile == TMEM lane
    const int row = m0 + rl;
    const bool row_ok = row < p.P;
    const bool drop = p.keep > 0.f && p.keep < 1.f;
    const uint32_t thr16 = drop ? ltg_keep_threshold16(p.keep) : 65536u;
    const float inv_keep = drop ? 1.0f / p.keep : 1.0f;
    const uint32_t step = p.rng_step + (p.rng_step_dev != nullptr ? *p.rng_step_dev : 0u);
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    __nv_bfloat16* hrow = p.Hd + (size_t)(row_ok ? row : 0) * p.k3;

    // ---- branch 1: columns [0, off2) of the hidden activation
    {
      const uint32_t key = drop ? ltg_hash_key(p.seed, p.rng_stream, step) : 0u;
      mbar_wait(&bar_mma[0], 0);
      tc_fence_after();
      for (int c = quarter * 16; c < p.off2; c += 64) {
        float v[16];
        tmem_ld16(taddr + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
        if (drop) drop16(v, key, thr16, inv_keep, row, p.ld1, c);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g0 = c + 8 * h;
          if (g0 < p.off2) {
            const uint4 u = pack8(v + 8 * h);
            hd_smem_store(hd, rl, g0, u);
            if (row_ok) *reinterpret_cast<uint4*>(hrow + g0) = u;
          }
        }
      }
    }
    // ---- branch 2: columns [off2, k3): h2 activations, the ones column (fc1 bias row of W3), zero padding
    {
      const uint32_t key = drop ? ltg_hash_key(p.seed, p.rng_stream + 1, step) : 0u;
      mbar_wait(&bar_mma[1], 0);
      tc_fence_after();
      for (int c = quarter * 16; c < p.k3 - p.off2; c += 64) {
        float v[16];
        tmem_ld16(taddr + DF_N1 + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
        if (drop) drop16(v, key, thr16, inv_keep, row, p.ld2, c);
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
            hd_smem_store(hd, rl, g0, u);
            if (row_ok) *reinterpret_cast<uint4*>(hrow + g0) = u;
          }
        }
      }
      // K padding of the last 64-k block: W3 rows >= k3 are zero-filled by TMA, the A side must be finite
      if (quarter == 0)
        for (int g0 = p.k3; g0 < p.kb3 * 64; g0 += 8) hd_smem_store(hd, rl, g0, make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_hd);

    // ---- fc1 + head
    {
      const uint32_t key = drop ? ltg_hash_key(p.seed, p.rng_stream + 2, step) : 0u;
      uint32_t yq[DF_CH3][8];                  // this thread's fc1 activations (bf16 pairs), kept for the backward pass
      float s = 0.f;
      mbar_wait(&bar_mma[2], 0);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < DF_CH3; ++j) {
        const int c = quarter * 16 + 64 * j;
        if (c < p.ld3) {
          float v[16];
          tmem_ld16(taddr + c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = tanh_approx(v[i]);
          if (drop) drop16(v, key, thr16, inv_keep, row, p.ld3, c);
          float w[16];
          load_w16(w, p.w4, c, p.ld3);
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
      s_part[quarter * 128 + rl] = s;
      named_bar_sync(1, GEMM_EPI_WARPS * 32);
      s = s_part[rl] + s_part[128 + rl] + s_part[256 + rl] + s_part[384 + rl] + __ldg(p.b4);
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
      if (bwd) {
        __nv_bfloat16* drow = p.dz3 + (size_t)(row_ok ? row : 0) * p.ld3;
#pragma unroll
        for (int j = 0; j < DF_CH3; ++j) {
          const int c = quarter * 16 + 64 * j;
          if (c < p.ld3) {                       // warp-uniform
            float d[16], gw[16], w[16];
            load_w16(w, p.w4, c, p.ld3);
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
            if (row_ok) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (c + 8 * h < p.ld3) *reinterpret_cast<uint4*>(drow + c + 8 * h) = pack8(d + 8 * h);
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
            float r2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const bool up = (lane & 4) != 0;
              const float keepv = up ? r4[2 + i] : r4[i], send = up ? r4[i] : r4[2 + i];
              r2[i] = keepv + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            float r1;
            {
              const bool up = (lane & 2) != 0;
              const float keepv = up ? r2[1] : r2[0], send = up ? r2[0] : r2[1];
              r1 = keepv + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
            const int col = c + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0 && col < p.ld3) atomicAdd(&s_dw4[col], r1);
          }
        }
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

}  // namespace

extern "C" int ltg_disc_fused_supported(int k1, int ld1, int ld2, int ld3, int off2, int one3, int h2, int k3) {
  return (k1 <= 128 && ld1 <= DF_N1 && off2 <= DF_N1 && off2 % 8 == 0 && ld1 % 8 == 0 && ld2 % 8 == 0 && ld3 % 8 == 0 && k3 % 8 == 0 &&
          k3 - off2 <= DF_N2 && ld2 <= DF_N2 && ld3 <= DF_MAXH3 && k3 <= DF_KB3_MAX * 64 && one3 == off2 + h2 && one3 < k3 && h2 <= ld2)
             ? 1 : 0;
}

extern "C" int ltg_disc_fwd_fused(const void* Xp_bf16, const void* Xn_bf16, int P, int k1, const void* W1_bf16, int ld1, const void* W2_bf16,
                                  int ld2, int h2, const void* W3_bf16, int ld3, int k3, int off2, int one3, const float* w4, const float* b4,
                                  const int32_t* label, float keep, uint64_t seed, uint32_t rng_stream, uint32_t rng_step,
                                  const uint32_t* rng_step_dev, void* Hd_bf16, float* y, float* scal, void* dz3_bf16, float* dw4, float* db4,
                                  void* stream) {
  LTG_REQUIRE(Xp_bf16 && Xn_bf16 && W1_bf16 && W2_bf16 && W3_bf16 && w4 && b4 && label && Hd_bf16 && scal);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(w4) & 15) == 0);
  LTG_REQUIRE(ltg_disc_fused_supported(k1, ld1, ld2, ld3, off2, one3, h2, k3));
  if (P <= 0) return LTG_OK;
  CUtensorMap tmXp, tmXn, tmW1, tmW2, tmW3;
  int rc;
  if ((rc = make_tmap_bf16(&tmXp, Xp_bf16, 128, (uint64_t)P, 128, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmXn, Xn_bf16, 128, (uint64_t)P, 128, 64, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16(&tmW1, W1_bf16, (uint64_t)ld1, (uint64_t)k1, (uint64_t)ld1, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmW2, W2_bf16, (uint64_t)ld2, (uint64_t)k1, (uint64_t)ld2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmW3, W3_bf16, (uint64_t)ld3, (uint64_t)k3, (uint64_t)ld3, 64, 64))) return rc;
  DiscFusedParams p;
  p.P = P; p.ld1 = ld1; p.ld2 = ld2; p.ld3 = ld3; p.off2 = off2; p.one3 = one3; p.h2 = h2; p.k3 = k3; p.kb3 = (k3 + 63) / 64;
  p.w4 = w4; p.b4 = b4; p.label = label; p.keep = keep; p.seed = seed; p.rng_stream = rng_stream; p.rng_step = rng_step;
  p.rng_step_dev = rng_step_dev;
  p.Hd = reinterpret_cast<__nv_bfloat16*>(Hd_bf16); p.y = y; p.scal = scal; p.dz3 = reinterpret_cast<__nv_bfloat16*>(dz3_bf16);
  p.dw4 = dw4; p.db4 = db4;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(disc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM);
    if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
    opted = true;
  }
  disc_fused_kernel<<<(P + GEMM_BM - 1) / GEMM_BM, GEMM_THREADS, DF_SMEM, (cudaStream_t)stream>>>(tmXp, tmXn, tmW1, tmW2, tmW3, p);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
