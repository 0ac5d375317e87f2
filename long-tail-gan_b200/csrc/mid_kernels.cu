// Fused "middle" of the MultiVAE: everything between the encoder gather and the decoder GEMM, forward and backward.
//
//   forward  (MultiVAE.py:151-162,178-181,168-172):  [mu|logvar] = h1 W_q1 + b_q1 ; KL ; z = mu + is_training*eps*exp(logvar/2) ;
//                                                     h2 = tanh(z W_p0 + b_p0)
//   backward (autodiff of the same lines, train.py:164):  dh2pre = dh2 (1-h2^2) ; dz = dh2pre W_p0^T ; dmulv = f(dz, KL) ;
//                                                     dh1 = dmulv W_q1^T ; dh1pre = dh1 (1-h1^2) ; bias gradients
//
// At batch 500 these are 0.36 GFLOP: as six tcgen05 GEMM launches + three element-wise launches they cost ~10 us each
// of fixed latency on the critical path of the step. Here one CTA owns 16 batch rows, keeps every intermediate in shared
// memory, streams its slice of the (L2-resident, 0.7 MB) bf16 weights through a 4-stage cp.async ring and multiplies with
// mma.sync m16n8k16 (bf16 in, fp32 accumulate). A first single-kernel version (32 CTAs) was instruction-issue bound on 32 SMs
// (ncu: IPC 1.4, 39 us); the columns of every 16-row tile are therefore split over 5 CTAs (160 CTAs) and each direction is two
// short kernels (GEMM + latent head, GEMM + tanh). The weight-gradient GEMMs (K = batch) stay on the tcgen05 kernel.
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int H = LTG_H;      // 600
constexpr int L = LTG_L;      // 200
constexpr int MT = 16;        // batch rows per CTA
constexpr int MID_THREADS = 256;
constexpr int NWARP = MID_THREADS / 32;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(s_u32(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(s_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// shared-memory pitch (elements) whose byte size is an odd multiple of 16 mod 128: ldmatrix rows hit 8 distinct bank groups
__host__ __device__ constexpr int pad_pitch(int n) { return (((n * 2) % 128) / 16) % 2 == 1 ? n : n + 8; }

// C[16 x NC] = A[16 x K] * Wop for a SLICE of NC output columns (a multiple of 8), A in shared memory (row-major, pitch PA,
// zero beyond K), the weight in global memory:
//   B_NK = false: W stored [K][N] (n contiguous, pitch ldw)      -- forward layers
//   B_NK = true : W stored [N][K] (k contiguous, pitch ldw)      -- backward: multiplication by the transposed weight
// The slice is the union of two column ranges: local 8-column tiles [0, T0) start at global column base0, tiles [T0, NT) at
// base1 (the latent head needs mu_j and logvar_j = column j and 200 + j in the same CTA). Warp w owns local tiles w, w+8, ...
template <int NC, int T0, int K, bool B_NK>
struct MidGemm {
  static constexpr int NT = NC / 8;
  static constexpr int TPW = (NT + NWARP - 1) / NWARP;
  static constexpr int KS = (K + 15) / 16;
  static constexpr int WP = B_NK ? 24 : pad_pitch(NC);                    // panel pitch (elements)
  static constexpr int PANEL = (B_NK ? NC : 16) * WP;                     // elements per k-panel
  static constexpr int STAGES = 4;                                        // panels in flight (the loop is L2-latency bound)
  static constexpr int SMEM_ELEMS = STAGES * PANEL;

  __device__ static __forceinline__ int gcol(int tile, int base0, int base1) { return tile < T0 ? base0 + 8 * tile : base1 + 8 * (tile - T0); }

  __device__ static void load_panel(__nv_bfloat16* sW, const __nv_bfloat16* __restrict__ W, int ldw, int ks, int base0, int base1) {
    const int k0 = ks * 16;
    if constexpr (B_NK) {
      // local rows n = 0..NC-1 (global row gcol(n/8) + n%8), 16 k values = two 16-byte chunks per row
      for (int c = threadIdx.x; c < NC * 2; c += MID_THREADS) {
        const int n = c >> 1, h = c & 1;
        const int k = k0 + h * 8;
        const bool ok = k < K;                                            // K % 8 == 0: a chunk is all-valid or all-padding
        const int gn = gcol(n >> 3, base0, base1) + (n & 7);
        cp_async16(sW + n * WP + h * 8, ok ? W + (size_t)gn * ldw + k : W, ok);
      }
    } else {
      for (int c = threadIdx.x; c < 16 * NT; c += MID_THREADS) {
        const int r = c / NT, q = c - r * NT;
        const bool ok = k0 + r < K;
        cp_async16(sW + r * WP + q * 8, ok ? W + (size_t)(k0 + r) * ldw + gcol(q, base0, base1) : W, ok);
      }
    }
  }

  __device__ static void run(float (&acc)[TPW][4], const __nv_bfloat16* sA, int PA, const __nv_bfloat16* __restrict__ W, int ldw,
                             __nv_bfloat16* sW, int base0, int base1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < TPW; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) {
      if (st < KS) load_panel(sW + st * PANEL, W, ldw, st, base0, base1);
      cp_async_commit();                 // always commit: group counting stays uniform
    }
    for (int ks = 0; ks < KS; ++ks) {
      cp_async_wait<STAGES - 2>();       // panel ks has landed (for this thread's copies)
      __syncthreads();                   // ... for everybody's, and everybody has finished computing on panel ks-1
      const int nxt = ks + STAGES - 1;
      if (nxt < KS) load_panel(sW + (nxt % STAGES) * PANEL, W, ldw, nxt, base0, base1);   // refills the buffer of panel ks-1
      cp_async_commit();
      const __nv_bfloat16* cur = sW + (ks % STAGES) * PANEL;
      uint32_t a[4];
      ldsm_x4(a, sA + (lane & 15) * PA + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int tile = warp + NWARP * i;
        if (tile < NT) {
          uint32_t b[2];
          const int l16 = lane & 15;
          if constexpr (B_NK) ldsm_x2(b, cur + (tile * 8 + (l16 & 7)) * WP + (l16 >> 3) * 8);
          else ldsm_x2_trans(b, cur + l16 * WP + tile * 8);
          mma_bf16(acc[i], a, b);
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();                     // the caller may reuse the panel buffers / the A tile
  }
};

constexpr int NG = 5;                                        // column groups per 16-row tile -> ceil(B/16) * 5 CTAs
constexpr int ZG = L / NG;                                   // 40 latent columns per group
constexpr int HG = H / NG;                                   // 120 hidden columns per group
using G1F = MidGemm<2 * ZG, ZG / 8, H, false>;               // [mu_g | logvar_g] = h1 W_q1[:, cols]      K=600, 80 columns
using G2F = MidGemm<HG, HG / 8, L, false>;                   // h2pre_g = z W_p0[:, cols]                  K=200, 120 columns
using G1B = MidGemm<ZG, ZG / 8, H, true>;                    // dz_g = dh2pre W_p0[rows g]^T               K=600, 40 columns
using G2B = MidGemm<HG, HG / 8, 2 * L, true>;                // dh1_g = dmulv W_q1[rows g]^T               K=400, 120 columns

constexpr int PA_H = pad_pitch(((H + 15) / 16) * 16);        // 608 -> 616
constexpr int PA_L = pad_pitch(((L + 15) / 16) * 16);        // 208 -> 216
constexpr int PA_2L = pad_pitch(2 * L);                      // 400 -> 408

constexpr size_t FWD_A_SMEM = (size_t)(MT * PA_H + G1F::SMEM_ELEMS) * 2 + (size_t)MT * 2 * ZG * 4;
constexpr size_t FWD_B_SMEM = (size_t)(MT * PA_L + G2F::SMEM_ELEMS) * 2;
constexpr size_t BWD_A_SMEM = (size_t)(MT * PA_H + G1B::SMEM_ELEMS) * 2 + (size_t)(MT * ZG + 2 * ZG) * 4;
constexpr size_t BWD_B_SMEM = (size_t)(MT * PA_2L + G2B::SMEM_ELEMS) * 2;

// 16-row x `cols`-column bf16 tile from global memory into shared memory (pitch PA), zero beyond B rows / `cols`
__device__ __forceinline__ void load_a_tile(__nv_bfloat16* sA, int PA, const __nv_bfloat16* __restrict__ src, int ld, int r0, int B, int cols) {
  for (int c = threadIdx.x; c < MT * (PA / 8); c += MID_THREADS) {
    const int r = c / (PA / 8), q = c - r * (PA / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < B && q * 8 < cols) v = *reinterpret_cast<const uint4*>(src + (size_t)(r0 + r) * ld + q * 8);
    *reinterpret_cast<uint4*>(sA + r * PA + q * 8) = v;
  }
}

// ---- forward, part A: [mu_g | logvar_g] = h1 W_q1 + b, KL, z_g = mu + is_training * eps * exp(logvar/2) --------------------------
__global__ void __launch_bounds__(MID_THREADS)
vae_mid_fwd_a_kernel(const __nv_bfloat16* __restrict__ h1, int ld_h1, const __nv_bfloat16* __restrict__ Wq1, const float* __restrict__ bq1,
                     const float* __restrict__ eps, int B, int64_t uid0, float is_training, uint64_t seed, uint32_t step,
                     const uint32_t* __restrict__ step_dev, float* __restrict__ mulv, __nv_bfloat16* __restrict__ z, int ld_z,
                     float* __restrict__ zmu, float* __restrict__ scal) {
  pdl_trigger();
  pdl_wait_cta();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [16][PA_H]  h1 tile
  __nv_bfloat16* sW = sA + MT * PA_H;
  float* sC = reinterpret_cast<float*>(sW + G1F::SMEM_ELEMS);              // [16][80]    mu_g | logvar_g
  __shared__ float s_red[NWARP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * MT, grp = blockIdx.y;
  const int base0 = grp * ZG, base1 = L + grp * ZG;
  if (step_dev != nullptr) step += *step_dev;
  load_a_tile(sA, PA_H, h1, ld_h1, r0, B, H);
  __syncthreads();
  float acc[G1F::TPW][4];
  G1F::run(acc, sA, PA_H, Wq1, 2 * L, sW, base0, base1);
  {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < G1F::TPW; ++i) {
      const int tile = warp + NWARP * i;
      if (tile < G1F::NT) {
        const int lc = tile * 8 + 2 * t;                                   // local column: [0,40) mu, [40,80) logvar
        const int gc = G1F::gcol(tile, base0, base1) + 2 * t;
        const float b0 = __ldg(bq1 + gc), b1 = __ldg(bq1 + gc + 1);
        sC[g * 2 * ZG + lc] = acc[i][0] + b0; sC[g * 2 * ZG + lc + 1] = acc[i][1] + b1;
        sC[(g + 8) * 2 * ZG + lc] = acc[i][2] + b0; sC[(g + 8) * 2 * ZG + lc + 1] = acc[i][3] + b1;
      }
    }
  }
  __syncthreads();
  float kl = 0.f;
  for (int c = tid; c < MT * ZG; c += MID_THREADS) {
    const int r = c / ZG, jl = c - r * ZG;
    const int row = r0 + r, j = base0 + jl;
    if (row < B) {
      const float mu = sC[r * 2 * ZG + jl], lv = sC[r * 2 * ZG + ZG + jl];
      mulv[(size_t)row * 2 * L + j] = mu;
      mulv[(size_t)row * 2 * L + L + j] = lv;
      kl += 0.5f * (-lv + expf(lv) + mu * mu - 1.0f);
      float e = 0.f;
      if (is_training != 0.f) {
        if (eps != nullptr) {
          e = eps[(size_t)row * L + j];
        } else {
          const uint64_t gi = (uint64_t)(uid0 + row) * (uint64_t)L + (uint64_t)j;
          Philox4 rr = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), LTG_STREAM_EPS, step, (uint32_t)seed, (uint32_t)(seed >> 32));
          e = sqrtf(-2.0f * logf(ltg_u01(rr.x))) * cospif(2.0f * ltg_u01(rr.y));
        }
      }
      const float d = is_training * e * expf(0.5f * lv);
      zmu[(size_t)row * L + j] = d;
      z[(size_t)row * ld_z + j] = __float2bfloat16(mu + d);
    }
  }
  kl = warp_sum(kl);
  if (lane == 0) s_red[warp] = kl;
  __syncthreads();
  if (tid == 0) {
    float tsum = 0.f;
    for (int w = 0; w < NWARP; ++w) tsum += s_red[w];
    atomicAdd(scal + LTG_S_KL_SUM, tsum);
  }
}

// ---- forward, part B: h2_g = tanh(z W_p0[:, cols g] + b) ------------------------------------------------------------------------
__global__ void __launch_bounds__(MID_THREADS)
vae_mid_fwd_b_kernel(const __nv_bfloat16* __restrict__ z, int ld_z, const __nv_bfloat16* __restrict__ Wp0, const float* __restrict__ bp0, int B,
                     __nv_bfloat16* __restrict__ h2, int ld_h2) {
  pdl_trigger();
  pdl_wait_cta();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sZ = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [16][PA_L]
  __nv_bfloat16* sW = sZ + MT * PA_L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * MT, base = blockIdx.y * HG;
  load_a_tile(sZ, PA_L, z, ld_z, r0, B, L);
  __syncthreads();
  float acc[G2F::TPW][4];
  G2F::run(acc, sZ, PA_L, Wp0, H, sW, base, base);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < G2F::TPW; ++i) {
    const int tile = warp + NWARP * i;
    if (tile < G2F::NT) {
      const int c0 = base + tile * 8 + 2 * t;
      const float b0 = __ldg(bp0 + c0), b1 = __ldg(bp0 + c0 + 1);
      if (r0 + g < B) *reinterpret_cast<uint32_t*>(h2 + (size_t)(r0 + g) * ld_h2 + c0) = pack_bf16x2(tanhf(acc[i][0] + b0), tanhf(acc[i][1] + b1));
      if (r0 + g + 8 < B) *reinterpret_cast<uint32_t*>(h2 + (size_t)(r0 + g + 8) * ld_h2 + c0) = pack_bf16x2(tanhf(acc[i][2] + b0), tanhf(acc[i][3] + b1));
    }
  }
}

// column sums of a fragment over the 16 rows of the tile -> atomicAdd(dst[col]) by the lanes with g == 0
__device__ __forceinline__ void frag_colsum_atomic(float v0, float v1, float* dst, int c0) {
  // lanes sharing t (lane & 3) hold the same columns for rows g = lane >> 2: reduce over g with xor 4, 8, 16
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
  if ((threadIdx.x & 31) < 4) { atomicAdd(dst + c0, v0); atomicAdd(dst + c0 + 1, v1); }
}

// ---- backward, part A: dz_g = dh2pre W_p0[rows g]^T ; dmulv_g = latent backward + anneal * dKL  (dh2pre comes from ltg_tanh_bwd,
//      which sums the split-K partials with full-machine parallelism; folding that sum in here made every group re-read them) ------
__global__ void __launch_bounds__(MID_THREADS)
vae_mid_bwd_a_kernel(const __nv_bfloat16* __restrict__ dh2pre, const __nv_bfloat16* __restrict__ Wp0, const float* __restrict__ mulv,
                     const float* __restrict__ zmu, int B, float inv_bg, float anneal, const float* __restrict__ scal,
                     __nv_bfloat16* __restrict__ dmulv, float* __restrict__ db_q1) {
  pdl_trigger();
  pdl_wait_cta();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [16][PA_H]   dh2pre tile
  __nv_bfloat16* sW = sA + MT * PA_H;
  float* sC = reinterpret_cast<float*>(sW + G1B::SMEM_ELEMS);              // [16][40] dz_g
  float* sCol = sC + MT * ZG;                                              // [80] column sums of dmulv_g
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * MT, grp = blockIdx.y;
  if (anneal < 0.f) anneal = scal[LTG_S_ANNEAL];
  for (int c = tid; c < 2 * ZG; c += MID_THREADS) sCol[c] = 0.f;
  load_a_tile(sA, PA_H, dh2pre, H, r0, B, H);
  __syncthreads();
  float acc[G1B::TPW][4];
  G1B::run(acc, sA, PA_H, Wp0, H, sW, grp * ZG, grp * ZG);     // dz_g[16][40]
  {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < G1B::TPW; ++i) {
      const int tile = warp + NWARP * i;
      if (tile < G1B::NT) {
        const int c0 = tile * 8 + 2 * t;
        sC[g * ZG + c0] = acc[i][0]; sC[g * ZG + c0 + 1] = acc[i][1];
        sC[(g + 8) * ZG + c0] = acc[i][2]; sC[(g + 8) * ZG + c0 + 1] = acc[i][3];
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < MT * 2 * ZG; e += MID_THREADS) {
    const int r = e / (2 * ZG), lc = e - r * (2 * ZG);
    const int row = r0 + r;
    if (row < B) {
      const bool is_mu = lc < ZG;
      const int jl = is_mu ? lc : lc - ZG;
      const int j = grp * ZG + jl;
      const float gz = sC[r * ZG + jl];
      float o;
      if (is_mu) {
        o = gz + anneal * mulv[(size_t)row * 2 * L + j] * inv_bg;
      } else {
        const float lv = mulv[(size_t)row * 2 * L + L + j];
        o = gz * zmu[(size_t)row * L + j] * 0.5f + anneal * 0.5f * (expf(lv) - 1.0f) * inv_bg;
      }
      dmulv[(size_t)row * 2 * L + (is_mu ? j : L + j)] = __float2bfloat16(o);
      atomicAdd(&sCol[lc], o);
    }
  }
  __syncthreads();
  for (int c = tid; c < 2 * ZG; c += MID_THREADS) atomicAdd(db_q1 + (c < ZG ? grp * ZG + c : L + grp * ZG + (c - ZG)), sCol[c]);
}

// ---- backward, part B: dh1_g = dmulv W_q1[rows g]^T ; dh1pre = dh1 (1 - h1^2) ; db_q0 --------------------------------------------
__global__ void __launch_bounds__(MID_THREADS)
vae_mid_bwd_b_kernel(const __nv_bfloat16* __restrict__ dmulv, const __nv_bfloat16* __restrict__ Wq1, const __nv_bfloat16* __restrict__ h1,
                     int ld_h1, int B, float* __restrict__ dh1pre, __nv_bfloat16* __restrict__ dh1pre_b, float* __restrict__ db_q0) {
  pdl_trigger();
  pdl_wait_cta();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sD = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [16][PA_2L]  dmulv tile
  __nv_bfloat16* sW = sD + MT * PA_2L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * MT, base = blockIdx.y * HG;
  load_a_tile(sD, PA_2L, dmulv, 2 * L, r0, B, 2 * L);
  __syncthreads();
  float acc[G2B::TPW][4];
  G2B::run(acc, sD, PA_2L, Wq1, 2 * L, sW, base, base);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < G2B::TPW; ++i) {
    const int tile = warp + NWARP * i;   // warp-uniform
    if (tile < G2B::NT) {
      const int c0 = base + tile * 8 + 2 * t;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = r0 + g + 8 * hh;
        if (row < B) {
          const float2 hv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h1 + (size_t)row * ld_h1 + c0));
          o[2 * hh] = acc[i][2 * hh] * (1.0f - hv.x * hv.x);
          o[2 * hh + 1] = acc[i][2 * hh + 1] * (1.0f - hv.y * hv.y);
          *reinterpret_cast<float2*>(dh1pre + (size_t)row * H + c0) = make_float2(o[2 * hh], o[2 * hh + 1]);
          *reinterpret_cast<uint32_t*>(dh1pre_b + (size_t)row * H + c0) = pack_bf16x2(o[2 * hh], o[2 * hh + 1]);
        }
      }
      frag_colsum_atomic(o[0] + o[2], o[1] + o[3], db_q0, c0);
    }
  }
}

template <class Kern>
int opt_in_smem(Kern kern, size_t bytes, bool* done) {
  if (*done) return LTG_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
  *done = true;
  return LTG_OK;
}

}  // namespace

extern "C" int ltg_vae_mid_fwd(const void* h1_bf16, int ld_h1, const void* Wq1_bf16, const float* b_q1, const void* Wp0_bf16,
                               const float* b_p0, const float* eps, int B, int64_t uid0, float is_training, uint64_t seed, uint32_t step,
                               const uint32_t* step_dev, float* mulv, void* z_bf16, int ld_z, float* zmu, void* h2_bf16, int ld_h2,
                               float* scal, void* stream) {
  LTG_REQUIRE(h1_bf16 && Wq1_bf16 && b_q1 && Wp0_bf16 && b_p0 && mulv && z_bf16 && zmu && h2_bf16 && scal);
  LTG_REQUIRE(ld_h1 % 8 == 0 && ld_h1 >= H && ld_h2 % 2 == 0 && ld_h2 >= H && ld_z % 8 == 0 && ld_z >= L);
  if (B <= 0) return LTG_OK;
  static bool o1 = false, o2 = false;
  int rc = opt_in_smem(vae_mid_fwd_a_kernel, FWD_A_SMEM, &o1);
  if (rc) return rc;
  rc = opt_in_smem(vae_mid_fwd_b_kernel, FWD_B_SMEM, &o2);
  if (rc) return rc;
  const dim3 grid((B + MT - 1) / MT, NG);
  ltg_launch(vae_mid_fwd_a_kernel, dim3(grid), dim3(MID_THREADS), FWD_A_SMEM, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(h1_bf16), ld_h1, reinterpret_cast<const __nv_bfloat16*>(Wq1_bf16), b_q1, eps, B, uid0, is_training,
      seed, step, step_dev, mulv, reinterpret_cast<__nv_bfloat16*>(z_bf16), ld_z, zmu, scal);
  LTG_CHECK_LAUNCH();
  ltg_launch(vae_mid_fwd_b_kernel, dim3(grid), dim3(MID_THREADS), FWD_B_SMEM, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(z_bf16), ld_z, reinterpret_cast<const __nv_bfloat16*>(Wp0_bf16), b_p0, B,
      reinterpret_cast<__nv_bfloat16*>(h2_bf16), ld_h2);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_vae_mid_bwd(const void* dh2pre_bf16, const void* Wp0_bf16, const void* Wq1_bf16, const float* mulv, const float* zmu,
                               const void* h1_bf16, int ld_h1, int B, int B_global, float anneal, const float* scal, void* dmulv_bf16,
                               float* dh1pre, void* dh1pre_bf16, float* db_q1, float* db_q0, void* stream) {
  LTG_REQUIRE(dh2pre_bf16 && Wp0_bf16 && Wq1_bf16 && mulv && zmu && h1_bf16 && dmulv_bf16 && dh1pre && dh1pre_bf16);
  LTG_REQUIRE(db_q1 && db_q0 && (anneal >= 0.f || scal != nullptr));
  LTG_REQUIRE(ld_h1 % 2 == 0 && ld_h1 >= H);
  if (B <= 0) return LTG_OK;
  static bool o1 = false, o2 = false;
  int rc = opt_in_smem(vae_mid_bwd_a_kernel, BWD_A_SMEM, &o1);
  if (rc) return rc;
  rc = opt_in_smem(vae_mid_bwd_b_kernel, BWD_B_SMEM, &o2);
  if (rc) return rc;
  const dim3 grid((B + MT - 1) / MT, NG);
  ltg_launch(vae_mid_bwd_a_kernel, dim3(grid), dim3(MID_THREADS), BWD_A_SMEM, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(dh2pre_bf16), reinterpret_cast<const __nv_bfloat16*>(Wp0_bf16), mulv, zmu, B,
      1.0f / (float)B_global, anneal, scal, reinterpret_cast<__nv_bfloat16*>(dmulv_bf16), db_q1);
  LTG_CHECK_LAUNCH();
  ltg_launch(vae_mid_bwd_b_kernel, dim3(grid), dim3(MID_THREADS), BWD_B_SMEM, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(dmulv_bf16), reinterpret_cast<const __nv_bfloat16*>(Wq1_bf16),
      reinterpret_cast<const __nv_bfloat16*>(h1_bf16), ld_h1, B, dh1pre, reinterpret_cast<__nv_bfloat16*>(dh1pre_bf16), db_q0);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
