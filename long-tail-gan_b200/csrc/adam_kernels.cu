// a13: TensorFlow-1 semantics Adam (train.py:160-164), fused with the bf16 shadow write, and the
// encoder-weight variant whose (sparse) gradient is built compactly for the batch's active items.
//
//   m <- b1*m + (1-b1)*g ;  v <- b2*v + (1-b2)*g*g ;  p <- p - lr_t * m / (sqrt(v) + eps)
//   lr_t = lr*sqrt(1-b2^t)/(1-b1^t)  (epsilon outside the bias correction: TF's "epsilon hat")
// Dense over every element (F7): rows whose gradient is zero still move through their momentum.
//
// Pure streaming kernels: 16-byte loads/stores with L1 no-allocate, 26 B of HBM traffic per parameter
// (p,m,v read + write, bf16 shadow write) plus 4 B for an explicit gradient.
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int H = LTG_H;

__device__ __forceinline__ void adam_update4(float4& p, float4& m, float4& v, const float4 g, float lr_t, float b1, float b2, float eps) {
  m.x = b1 * m.x + (1.f - b1) * g.x; m.y = b1 * m.y + (1.f - b1) * g.y;
  m.z = b1 * m.z + (1.f - b1) * g.z; m.w = b1 * m.w + (1.f - b1) * g.w;
  v.x = b2 * v.x + (1.f - b2) * g.x * g.x; v.y = b2 * v.y + (1.f - b2) * g.y * g.y;
  v.z = b2 * v.z + (1.f - b2) * g.z * g.z; v.w = b2 * v.w + (1.f - b2) * g.w * g.w;
  p.x -= lr_t * m.x / (sqrtf(v.x) + eps); p.y -= lr_t * m.y / (sqrtf(v.y) + eps);
  p.z -= lr_t * m.z / (sqrtf(v.z) + eps); p.w -= lr_t * m.w / (sqrtf(v.w) + eps);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int n_partials,
            int64_t partial_stride, __nv_bfloat16* __restrict__ shadow, int64_t n, float lr_t, const float* __restrict__ scal, float b1,
            float b2, float eps) {
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    float4 gg = ld_stream_f4(g + 4 * i);
    for (int sp = 1; sp < n_partials; ++sp) {  // split-K partials of the producing GEMM
      const float4 o = ld_stream_f4(g + (size_t)sp * partial_stride + 4 * i);
      gg.x += o.x; gg.y += o.y; gg.z += o.z; gg.w += o.w;
    }
    adam_update4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    st_stream_f4(p + 4 * i, pp); st_stream_f4(m + 4 * i, mm); st_stream_f4(v + 4 * i, vv);
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(shadow + 4 * i) = s;
    }
  }
  // tail (n % 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float gg = g[i];
    for (int sp = 1; sp < n_partials; ++sp) gg += g[(size_t)sp * partial_stride + i];
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    const float pp = p[i] - lr_t * mm / (sqrtf(vv) + eps);
    m[i] = mm; v[i] = vv; p[i] = pp;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pp);
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
                float b1, float b2, float eps) {
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    const int slot = __ldg(slot_of_item + item);
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) gg = __ldg(reinterpret_cast<const float4*>(G + (size_t)slot * H) + c4);
    adam_update4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    st_stream_f4(p + 4 * i, pp); st_stream_f4(m + 4 * i, mm); st_stream_f4(v + 4 * i, vv);
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(shadow + 4 * i) = s;
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

int grid_for(int64_t work_items, int threads, int max_blocks) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

// 148 SMs x 8 resident 256-thread CTAs: one full wave, grid-stride over the rest
static const int kStreamBlocks = 148 * 8;

extern "C" int ltg_adam(float* p, float* m, float* v, const float* g, int n_partials, int64_t partial_stride, void* shadow_bf16, int64_t n,
                        float lr_t, const float* scal, float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(p && m && v && g && n_partials >= 1);
  LTG_REQUIRE(n_partials == 1 || (partial_stride % 4 == 0 && partial_stride >= n));
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(g)) & 15) == 0);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  if (n <= 0) return LTG_OK;
  adam_kernel<<<grid_for(n >> 2, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(
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
                            float lr_t, const float* scal, float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(p && m && v && slot_of_item && G);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  if (n_items <= 0) return LTG_OK;
  enc_adam_kernel<<<grid_for((int64_t)n_items * H4, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(
      p, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n_items, slot_of_item, G, lr_t, scal, beta1, beta2, eps);
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
