// a13: TensorFlow-1 semantics Adam (train.py:160-164), fused with the bf16 shadow write, and the
// encoder-weight variant that rebuilds its (sparse) gradient row on the fly from the batch CSC.
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
constexpr int H4 = H / 4;  // 150 float4 per encoder row

__device__ __forceinline__ void adam_update4(float4& p, float4& m, float4& v, const float4 g, float lr_t, float b1, float b2, float eps) {
  m.x = b1 * m.x + (1.f - b1) * g.x; m.y = b1 * m.y + (1.f - b1) * g.y;
  m.z = b1 * m.z + (1.f - b1) * g.z; m.w = b1 * m.w + (1.f - b1) * g.w;
  v.x = b2 * v.x + (1.f - b2) * g.x * g.x; v.y = b2 * v.y + (1.f - b2) * g.y * g.y;
  v.z = b2 * v.z + (1.f - b2) * g.z * g.z; v.w = b2 * v.w + (1.f - b2) * g.w * g.w;
  p.x -= lr_t * m.x / (sqrtf(v.x) + eps); p.y -= lr_t * m.y / (sqrtf(v.y) + eps);
  p.z -= lr_t * m.z / (sqrtf(v.z) + eps); p.w -= lr_t * m.w / (sqrtf(v.w) + eps);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            __nv_bfloat16* __restrict__ shadow, int64_t n, float lr_t, const float* __restrict__ scal, float b1, float b2, float eps) {
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    const float4 gg = ld_stream_f4(g + 4 * i);
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
    const float gg = g[i];
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    const float pp = p[i] - lr_t * mm / (sqrtf(vv) + eps);
    m[i] = mm; v[i] = vv; p[i] = pp;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pp);
  }
}

// gradient of W_q0 row `item`, 4 columns starting at c4*4: sum over the batch rows that contain the item
__device__ __forceinline__ float4 enc_row_grad(int item, int c4, const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_row,
                                               const int32_t* __restrict__ csc_pos, const float* __restrict__ coef,
                                               const float* __restrict__ dh1, int ld) {
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  const int e0 = __ldg(csc_ptr + item), e1 = __ldg(csc_ptr + item + 1);
  // hot items appear in hundreds of batch rows: keep 4 independent (index -> coef, index -> dh1 row) chains in flight
  for (int e = e0; e < e1; e += 4) {
    int pos[4], row[4];
    float c[4];
    float4 d[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = e + q < e1;
      pos[q] = ok ? __ldg(csc_pos + e + q) : -1;
      row[q] = ok ? __ldg(csc_row + e + q) : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      c[q] = pos[q] >= 0 ? __ldg(coef + pos[q]) : 0.f;
      d[q] = __ldg(reinterpret_cast<const float4*>(dh1 + (size_t)row[q] * ld) + c4);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      g.x = fmaf(c[q], d[q].x, g.x); g.y = fmaf(c[q], d[q].y, g.y); g.z = fmaf(c[q], d[q].z, g.z); g.w = fmaf(c[q], d[q].w, g.w);
    }
  }
  return g;
}

__global__ void __launch_bounds__(256)
enc_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int n_items,
                const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_row, const int32_t* __restrict__ csc_pos,
                const float* __restrict__ coef, const float* __restrict__ dh1, int ld, float lr_t, const float* __restrict__ scal,
                float b1, float b2, float eps) {
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    const float4 gg = enc_row_grad(item, c4, csc_ptr, csc_row, csc_pos, coef, dh1, ld);
    adam_update4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    st_stream_f4(p + 4 * i, pp); st_stream_f4(m + 4 * i, mm); st_stream_f4(v + 4 * i, vv);
    if (shadow != nullptr) {
      uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(shadow + 4 * i) = s;
    }
  }
}

__global__ void __launch_bounds__(256)
enc_wgrad_kernel(float* __restrict__ dW, int n_items, const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_row,
                 const int32_t* __restrict__ csc_pos, const float* __restrict__ coef, const float* __restrict__ dh1, int ld) {
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    st_stream_f4(dW + 4 * i, enc_row_grad(item, c4, csc_ptr, csc_row, csc_pos, coef, dh1, ld));
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

extern "C" int ltg_adam(float* p, float* m, float* v, const float* g, void* shadow_bf16, int64_t n, float lr_t, const float* scal,
                        float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(p && m && v && g);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(g)) & 15) == 0);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  if (n <= 0) return LTG_OK;
  adam_kernel<<<grid_for(n >> 2, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(p, m, v, g, reinterpret_cast<__nv_bfloat16*>(shadow_bf16),
                                                                                       n, lr_t, scal, beta1, beta2, eps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_adam(float* p, float* m, float* v, void* shadow_bf16, int n_items, const int32_t* csc_ptr, const int32_t* csc_row,
                            const int32_t* csc_pos, const float* coef, const float* dh1pre, int ld_dh1, float lr_t, const float* scal,
                            float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(p && m && v && csc_ptr && csc_row && csc_pos && coef && dh1pre);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(ld_dh1 % 4 == 0 && ld_dh1 >= H);
  if (n_items <= 0) return LTG_OK;
  enc_adam_kernel<<<grid_for((int64_t)n_items * H4, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(
      p, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n_items, csc_ptr, csc_row, csc_pos, coef, dh1pre, ld_dh1, lr_t, scal, beta1, beta2, eps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_wgrad(float* dW, int n_items, const int32_t* csc_ptr, const int32_t* csc_row, const int32_t* csc_pos,
                             const float* coef, const float* dh1pre, int ld_dh1, void* stream) {
  LTG_REQUIRE(dW && csc_ptr && csc_row && csc_pos && coef && dh1pre);
  LTG_REQUIRE(ld_dh1 % 4 == 0 && ld_dh1 >= H);
  if (n_items <= 0) return LTG_OK;
  enc_wgrad_kernel<<<grid_for((int64_t)n_items * H4, 256, kStreamBlocks), 256, 0, (cudaStream_t)stream>>>(dW, n_items, csc_ptr, csc_row, csc_pos,
                                                                                                         coef, dh1pre, ld_dh1);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
