// a13: TensorFlow-1 semantics Adam (train.py:160-164), fused with the bf16 shadow write, and the
// encoder-weight variant whose (sparse) gradient is built compactly for the batch's active items.
//
//   m <- b1*m + (1-b1)*g ;  v <- b2*v + (1-b2)*g*g ;  p <- p - lr_t * m / (sqrt(v) + eps)
//   lr_t = lr*sqrt(1-b2^t)/(1-b1^t)  (epsilon outside the bias correction: TF's "epsilon hat")
// Dense over every element (F7): rows whose gradient is zero still move through their momentum.
//
// Pure streaming kernels: 16-byte loads/stores with L1 no-allocate, 26 B of HBM traffic per parameter
// (p,m,v read + write, bf16 shadow write) plus 4 B for an explicit gradient.
#include <stdlib.h>
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int H = LTG_H;

__device__ __forceinline__ void adam_update4(float4& p, float4& m, float4& v, const float4 g, float lr_t, float b1, float b2, float eps) {
  ltg_adam4(p, m, v, g, lr_t, b1, b2, eps);
}

// Non-persistent grid: every CTA owns one fixed chunk of ADAM_UN x 256 float4 and exits. The sweeps run on low-priority side
// branches of the step graph beside the latency-bound critical chain; a grid-stride kernel with 8 resident CTAs per SM held every
// thread slot of the GPU for its whole 60 us and the critical chain's kernels waited for it (timeline: vae_mid_bwd_b 58 us under the
// sweep vs 12 us alone). With short-lived CTAs the block scheduler hands freed slots to the higher-priority kernels first.
constexpr int ADAM_UN = 2;

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int n_partials,
            int64_t partial_stride, __nv_bfloat16* __restrict__ shadow, int64_t n, float lr_t, const float* __restrict__ scal, float b1,
            float b2, float eps) {
  pdl_trigger();
  pdl_wait_cta();
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = n >> 2;
  const int64_t base = (int64_t)blockIdx.x * (256 * ADAM_UN) + threadIdx.x;
  const uint64_t pol = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  float4 pp[ADAM_UN], mm[ADAM_UN], vv[ADAM_UN], gg[ADAM_UN];
#pragma unroll
  for (int u = 0; u < ADAM_UN; ++u) {          // all loads of the chunk first: 8 independent 16-byte requests per thread
    const int64_t i = base + u * 256;
    if (i < n4) {
      pp[u] = ld_stream_f4_hint(p + 4 * i, pol); mm[u] = ld_stream_f4_hint(m + 4 * i, pol); vv[u] = ld_stream_f4_hint(v + 4 * i, pol);
      gg[u] = ld_stream_f4_hint(g + 4 * i, pol);
    }
  }
  if (n_partials > 1) {
    // split-K partials of the producing GEMM: rounds of 4 partials x ADAM_UN chunks = 8 independent 16-byte requests per thread
    // (the discriminator update sums 16 partials on the critical chain of the step: 5 dependent rounds of 3 before)
    constexpr int PR = 4;
    for (int sp = 1; sp < n_partials; sp += PR) {
      float4 o[PR][ADAM_UN];
#pragma unroll
      for (int q = 0; q < PR; ++q)
#pragma unroll
        for (int u = 0; u < ADAM_UN; ++u) {
          const int64_t i = base + u * 256;
          o[q][u] = (sp + q < n_partials && i < n4) ? ld_stream_f4(g + (size_t)(sp + q) * partial_stride + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < ADAM_UN; ++u) {
        const float4 a = o[0][u], b = o[1][u], c = o[2][u], d = o[3][u];
        gg[u].x += (a.x + b.x) + (c.x + d.x); gg[u].y += (a.y + b.y) + (c.y + d.y);
        gg[u].z += (a.z + b.z) + (c.z + d.z); gg[u].w += (a.w + b.w) + (c.w + d.w);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < ADAM_UN; ++u) {
    const int64_t i = base + u * 256;
    if (i >= n4) continue;
    adam_update4(pp[u], mm[u], vv[u], gg[u], lr_t, b1, b2, eps);
    st_stream_f4_hint(p + 4 * i, pp[u], pol); st_stream_f4_hint(m + 4 * i, mm[u], pol); st_stream_f4_hint(v + 4 * i, vv[u], pol);
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp[u].x, pp[u].y); s.y = pack_bf16x2(pp[u].z, pp[u].w);
      st_b64_hint(shadow + 4 * i, s, pol_keep);
    }
  }
  // tail (n % 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float gt = g[i];
    for (int sp = 1; sp < n_partials; ++sp) gt += g[(size_t)sp * partial_stride + i];
    float mt = m[i], vt = v[i], pt = p[i];
    ltg_adam1(pt, mt, vt, gt, lr_t, b1, b2, eps);
    m[i] = mt; v[i] = vt; p[i] = pt;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pt);
  }
}

// Adam whose gradient is the sum of split-K partials (the discriminator update: 161 k parameters, 16+ partials of the three
// weight-gradient GEMMs), laid out so that EVERY load of the launch is issued in one round: a CTA owns 32 float4 and its eight
// warps each fetch four partials of them (warp 0 also p, m, v and partial 0), fold them as (a+b)+(c+d) and hand the result over
// through shared memory; warp 0 adds the eight sums in partial order and runs the update. adam_kernel's loop over rounds of four
// partials waited for one full memory round trip per round -- on the critical chain of the step, beside HBM-saturating sweeps, that
// was 29 us for 0.6 MB of parameters (timeline of round 2). Same operations in the same order as that loop: bit-identical result.
constexpr int AP_WARPS = 8;
constexpr int AP_MAX_PARTIALS = 1 + 4 * AP_WARPS;

__global__ void __launch_bounds__(32 * AP_WARPS)
adam_partials_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int n_partials,
                     int64_t partial_stride, __nv_bfloat16* __restrict__ shadow, int64_t n, float lr_t, const float* __restrict__ scal,
                     float b1, float b2, float eps) {
  pdl_trigger();
  pdl_wait_cta();
  __shared__ float4 s_sum[AP_WARPS][32];
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = n >> 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = i < n4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 pp = zero, mm = zero, vv = zero, gg = zero, o[4];
  if (w == 0 && ok) { pp = ld_stream_f4(p + 4 * i); mm = ld_stream_f4(m + 4 * i); vv = ld_stream_f4(v + 4 * i); gg = ld_stream_f4(g + 4 * i); }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int sp = 1 + 4 * w + q;
    o[q] = (ok && sp < n_partials) ? ld_stream_f4(g + (size_t)sp * partial_stride + 4 * i) : zero;
  }
  s_sum[w][lane] = make_float4((o[0].x + o[1].x) + (o[2].x + o[3].x), (o[0].y + o[1].y) + (o[2].y + o[3].y),
                               (o[0].z + o[1].z) + (o[2].z + o[3].z), (o[0].w + o[1].w) + (o[2].w + o[3].w));
  __syncthreads();
  if (w == 0 && ok) {
    for (int k = 0; 1 + 4 * k < n_partials; ++k) {
      const float4 s = s_sum[k][lane];
      gg.x += s.x; gg.y += s.y; gg.z += s.z; gg.w += s.w;
    }
    adam_update4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    *reinterpret_cast<float4*>(p + 4 * i) = pp; *reinterpret_cast<float4*>(m + 4 * i) = mm; *reinterpret_cast<float4*>(v + 4 * i) = vv;
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(shadow + 4 * i) = s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // tail (n % 4)
    const int64_t t = (n4 << 2) + threadIdx.x;
    float gt = g[t];
    for (int sp = 1; sp < n_partials; ++sp) gt += g[(size_t)sp * partial_stride + t];
    float mt = m[t], vt = v[t], pt = p[t];
    ltg_adam1(pt, mt, vt, gt, lr_t, b1, b2, eps);
    m[t] = mt; v[t] = vt; p[t] = pt;
    if (shadow != nullptr) shadow[t] = __float2bfloat16(pt);
  }
}

// ---------------------------------------------------------------------------------------------
// Encoder weight W_q0 [I, 600]: its gradient X^T dh1pre is non-zero only on the items that occur in the batch
// ("active" items, about a third of the catalog at B = 500). It is built compactly -- G[slot, :] for the active
// items only -- by one CTA per active item, and the dense TF-Adam sweep (every row moves) then streams
// p/m/v exactly like adam_kernel and fetches its gradient row through slot_of_item[] (-1: zero gradient).
// A first version rebuilt the gradient inside the streaming kernel; ncu showed 60% long-scoreboard stalls on the
// dependent index -> coefficient -> row gathers and 2.8 TB/s instead of 6.
// ---------------------------------------------------------------------------------------------
constexpr int WG_GROUPS = 4;           // entry groups per CTA (hot items sit in hundreds of batch rows)
constexpr int WG_LANES = 160;          // 150 float4 column slots, padded to 5 warps
constexpr int WG_THREADS = WG_GROUPS * WG_LANES;
constexpr int H4 = H / 4;

__global__ void __launch_bounds__(WG_THREADS)
enc_wgrad_compact_kernel(float* __restrict__ G, const int32_t* __restrict__ act_ptr, const int32_t* __restrict__ csc_row,
                         const int32_t* __restrict__ csc_pos, const float* __restrict__ coef, const float* __restrict__ dh1, int ld) {
  __shared__ float4 s_part[WG_GROUPS - 1][H4];
  const int slot = blockIdx.x;
  const int grp = threadIdx.x / WG_LANES;
  const int c4 = threadIdx.x - grp * WG_LANES;
  const int e0 = __ldg(act_ptr + slot), e1 = __ldg(act_ptr + slot + 1);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < H4) {
    constexpr int U = 4;
    for (int e = e0 + grp * U; e < e1; e += WG_GROUPS * U) {
      float c[U];
      float4 d[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const bool ok = e + q < e1;
        c[q] = ok ? __ldg(coef + __ldg(csc_pos + e + q)) : 0.f;
        d[q] = __ldg(reinterpret_cast<const float4*>(dh1 + (size_t)(ok ? __ldg(csc_row + e + q) : 0) * ld) + c4);
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        g.x = fmaf(c[q], d[q].x, g.x); g.y = fmaf(c[q], d[q].y, g.y); g.z = fmaf(c[q], d[q].z, g.z); g.w = fmaf(c[q], d[q].w, g.w);
      }
    }
    if (grp > 0) s_part[grp - 1][c4] = g;
  }
  __syncthreads();
  if (grp == 0 && c4 < H4) {
#pragma unroll
    for (int k = 0; k < WG_GROUPS - 1; ++k) {
      const float4 o = s_part[k][c4];
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    reinterpret_cast<float4*>(G + (size_t)slot * H)[c4] = g;
  }
}

__global__ void __launch_bounds__(256)
enc_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int n_items,
                const int32_t* __restrict__ slot_of_item, const float* __restrict__ G, float lr_t, const float* __restrict__ scal,
                float b1, float b2, float eps, int rows) {
  pdl_trigger();
  pdl_wait_cta();
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t base = (int64_t)blockIdx.x * (256 * ADAM_UN) + threadIdx.x;   // one fixed chunk per CTA (see adam_kernel)
  const uint64_t pol = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  float4 pp[ADAM_UN], mm[ADAM_UN], vv[ADAM_UN], gg[ADAM_UN];
  bool on[ADAM_UN];
#pragma unroll
  for (int u = 0; u < ADAM_UN; ++u) {
    const int64_t i = base + u * 256;
    on[u] = false;
    if (i >= n4) continue;
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    const int slot = __ldg(slot_of_item + item);
    // rows: 0 every row; 1 only rows without a gradient (their update needs nothing from this step's backward pass, so it can run
    // while the forward chain leaves HBM idle); 2 only the batch's active rows
    if ((rows == 1 && slot >= 0) || (rows == 2 && slot < 0)) continue;
    on[u] = true;
    pp[u] = ld_stream_f4_hint(p + 4 * i, pol); mm[u] = ld_stream_f4_hint(m + 4 * i, pol); vv[u] = ld_stream_f4_hint(v + 4 * i, pol);
    gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) gg[u] = __ldg(reinterpret_cast<const float4*>(G + (size_t)slot * H) + c4);
  }
#pragma unroll
  for (int u = 0; u < ADAM_UN; ++u) {
    if (!on[u]) continue;
    const int64_t i = base + u * 256;
    adam_update4(pp[u], mm[u], vv[u], gg[u], lr_t, b1, b2, eps);
    st_stream_f4_hint(p + 4 * i, pp[u], pol); st_stream_f4_hint(m + 4 * i, mm[u], pol); st_stream_f4_hint(v + 4 * i, vv[u], pol);
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp[u].x, pp[u].y); s.y = pack_bf16x2(pp[u].z, pp[u].w);
      st_b64_hint(shadow + 4 * i, s, pol_keep);
    }
  }
}

// dense gradient (parity checks / data-parallel all-reduce): dW[item, :] = G[slot_of_item[item], :] or 0
__global__ void __launch_bounds__(256)
enc_wgrad_expand_kernel(float* __restrict__ dW, int n_items, const int32_t* __restrict__ slot_of_item, const float* __restrict__ G) {
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    const int slot = __ldg(slot_of_item + item);
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) gg = __ldg(reinterpret_cast<const float4*>(G + (size_t)slot * H) + c4);
    st_stream_f4(dW + 4 * i, gg);
  }
}

// Xc (dense bf16 coefficient matrix of the batch, written by ltg_enc_gather_fwd) back to all-zero after its consumer: clears
// exactly the entries the forward wrote (one 2-byte store per interaction instead of a fill of the whole matrix)
__global__ void __launch_bounds__(256)
enc_xc_clear_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int B, const int32_t* __restrict__ slot_of_item,
                    __nv_bfloat16* __restrict__ xc, int ld_xc) {
  pdl_trigger();
  pdl_wait_cta();
  const int e0 = indptr[0];
  const int n = indptr[B] - e0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    // row of entry e: binary search in indptr[0..B]
    int lo = 0, hi = B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (indptr[mid] - e0 <= e) lo = mid; else hi = mid; }
    xc[(size_t)lo * ld_xc + slot_of_item[indices[e0 + e]]] = __float2bfloat16(0.f);
  }
}

int grid_for(int64_t work_items, int threads, int max_blocks) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

// 148 SMs x 8 resident 256-thread CTAs: one full wave, grid-stride over the rest
static const int kStreamBlocks = 148 * 8;

// Dynamic shared memory requested by the Adam sweeps only to bound how many of their CTAs an SM holds at once (LTG_ADAM_CTAS_PER_SM,
// default 8 = no bound): a sweep that keeps fewer bytes in flight leaves HBM queues shorter for the latency-bound kernels beside it.
static size_t adam_throttle_smem() {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("LTG_ADAM_CTAS_PER_SM");
    const int n = e != nullptr ? atoi(e) : 8;
    v = (n >= 8 || n <= 0) ? 0 : (long)((227 * 1024) / n - 1024) / 1024 * 1024;
    if (v > 0) {
      cudaFuncSetAttribute(adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v);
      cudaFuncSetAttribute(enc_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v);
    }
  }
  return (size_t)v;
}

extern "C" int ltg_adam(float* p, float* m, float* v, const float* g, int n_partials, int64_t partial_stride, void* shadow_bf16, int64_t n,
                        float lr_t, const float* scal, float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(p && m && v && g && n_partials >= 1);
  LTG_REQUIRE(n_partials == 1 || (partial_stride % 4 == 0 && partial_stride >= n));
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(g)) & 15) == 0);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  if (n <= 0) return LTG_OK;
  static int one_round = -1;   // LTG_ADAM_ONE_ROUND=0: the loop over rounds of partials in adam_kernel (A/B switch)
  if (one_round < 0) { const char* e = getenv("LTG_ADAM_ONE_ROUND"); one_round = (e != nullptr && e[0] == '0') ? 0 : 1; }
  if (one_round && n_partials > 1 && n_partials <= AP_MAX_PARTIALS) {
    ltg_launch(adam_partials_kernel, dim3((unsigned)(((n >> 2) + 31) / 32 > 0 ? ((n >> 2) + 31) / 32 : 1)), dim3(32 * AP_WARPS), 0, (cudaStream_t)stream,
        p, m, v, g, n_partials, partial_stride, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n, lr_t, scal, beta1, beta2, eps);
    LTG_CHECK_LAUNCH();
    return LTG_OK;
  }
  ltg_launch(adam_kernel, dim3(grid_for(((n >> 2) + ADAM_UN - 1) / ADAM_UN, 256, 1 << 30)), dim3(256), n > (1 << 20) ? adam_throttle_smem() : 0, (cudaStream_t)stream, 
      p, m, v, g, n_partials, partial_stride, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n, lr_t, scal, beta1, beta2, eps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_wgrad_compact(float* G, int n_active, const int32_t* act_ptr, const int32_t* csc_row, const int32_t* csc_pos,
                                     const float* coef, const float* dh1pre, int ld_dh1, void* stream) {
  LTG_REQUIRE(G && act_ptr && csc_row && csc_pos && coef && dh1pre);
  LTG_REQUIRE(ld_dh1 % 4 == 0 && ld_dh1 >= H);
  if (n_active <= 0) return LTG_OK;
  enc_wgrad_compact_kernel<<<n_active, WG_THREADS, 0, (cudaStream_t)stream>>>(G, act_ptr, csc_row, csc_pos, coef, dh1pre, ld_dh1);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_adam(float* p, float* m, float* v, void* shadow_bf16, int n_items, const int32_t* slot_of_item, const float* G,
                            float lr_t, const float* scal, float beta1, float beta2, float eps, int rows, void* stream) {
  LTG_REQUIRE(p && m && v && slot_of_item && G && rows >= 0 && rows <= 2);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  if (n_items <= 0) return LTG_OK;
  ltg_launch(enc_adam_kernel, dim3(grid_for(((int64_t)n_items * H4 + ADAM_UN - 1) / ADAM_UN, 256, 1 << 30)), dim3(256), adam_throttle_smem(), (cudaStream_t)stream, 
      p, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n_items, slot_of_item, G, lr_t, scal, beta1, beta2, eps, rows);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_wgrad_expand(float* dW, int n_items, const int32_t* slot_of_item, const float* G, void* stream) {
  LTG_REQUIRE(dW && slot_of_item && G);
  if (n_items <= 0) return LTG_OK;
  enc_wgrad_expand_kernel<<<grid_for((int64_t)n_items * H4, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(dW, n_items, slot_of_item, G);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_xc_clear(const int32_t* indptr, const int32_t* indices, int B, int nnz_hint, const int32_t* slot_of_item, void* xc_bf16,
                                int ld_xc, void* stream) {
  LTG_REQUIRE(indptr && indices && slot_of_item && xc_bf16);
  if (B <= 0 || nnz_hint <= 0) return LTG_OK;
  ltg_launch(enc_xc_clear_kernel, dim3(grid_for(nnz_hint, 256, kStreamBlocks)), dim3(256), 0, (cudaStream_t)stream, 
      indptr, indices, B, slot_of_item, reinterpret_cast<__nv_bfloat16*>(xc_bf16), ld_xc);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

// out[i] = sum_s src[s * stride + i]: the split-K partials of the discriminator weight-gradient GEMMs folded into the one gradient
// arena the data-parallel exchange reads (replaces a torch.sum launch inside the captured step)
namespace {
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ src, int n_partials, int64_t stride, int64_t n4, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = ld_stream_f4(src + 4 * i), b = make_float4(0.f, 0.f, 0.f, 0.f), c = b, d = b;
  int sp = 1;
  for (; sp + 3 <= n_partials; sp += 3) {
    const float4 x = ld_stream_f4(src + (size_t)sp * stride + 4 * i), y = ld_stream_f4(src + (size_t)(sp + 1) * stride + 4 * i),
                 z = ld_stream_f4(src + (size_t)(sp + 2) * stride + 4 * i);
    b.x += x.x; b.y += x.y; b.z += x.z; b.w += x.w; c.x += y.x; c.y += y.y; c.z += y.z; c.w += y.w; d.x += z.x; d.y += z.y; d.z += z.z; d.w += z.w;
  }
  for (; sp < n_partials; ++sp) {
    const float4 x = ld_stream_f4(src + (size_t)sp * stride + 4 * i);
    b.x += x.x; b.y += x.y; b.z += x.z; b.w += x.w;
  }
  *reinterpret_cast<float4*>(out + 4 * i) = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
}
}  // namespace

extern "C" int ltg_sum_partials(const float* src, int n_partials, int64_t stride, int64_t n, float* out, void* stream) {
  LTG_REQUIRE(src && out && n_partials >= 1 && n % 4 == 0 && stride % 4 == 0);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  if (n <= 0) return LTG_OK;
  const int64_t n4 = n / 4;
  sum_partials_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n_partials, stride, n4, out);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
