// a11/a12: discriminator pieces that are not GEMMs (discriminator.py:14-55, train.py:142).
//   - frozen embedding gather (F5)
//   - head: fc2 (h3 -> 1) + sigmoid + both loss terms + backward seed
// The three dense layers run on the tcgen05 GEMM (gemm_ops.cu) with the bias+tanh+dropout epilogue.
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

// bf16 embedding rows are padded from h0=100 to 128 columns (256 B = 16 uint4 per row)

__global__ void disc_gather_kernel(const uint4* __restrict__ E, const int32_t* __restrict__ pop, const int32_t* __restrict__ niche, int P,
                                   uint4* __restrict__ Xp, uint4* __restrict__ Xn) {
  pdl_trigger();
  pdl_wait_cta();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte vector each; 16 per row
  const int64_t row = t >> 4;
  const int v = (int)(t & 15);
  if (row >= P) return;
  Xp[row * 16 + v] = __ldg(E + (size_t)pop[row] * 16 + v);
  Xn[row * 16 + v] = __ldg(E + (size_t)niche[row] * 16 + v);
}

// derivative of y = dropout(tanh(a)) w.r.t. a, recovered from the stored y: mask/keep * (1 - tanh^2)
__device__ __forceinline__ float dact_drop_tanh(float y, float keep, bool drop) {
  if (!drop) return 1.0f - y * y;
  if (y == 0.f) return 0.f;
  const float t = y * keep;
  return (1.0f - t * t) / keep;
}

constexpr int HEAD_THREADS = 256;  // 8 warps, each walks rows with a grid stride
constexpr int HEAD_MAXH3 = 320;    // 5 column pairs per lane
constexpr int HEAD_PPL = HEAD_MAXH3 / 64;

// One warp per row: lane l owns the column pairs (2l, 2l+1) + 64k (128-byte coalesced row segments, bf16x2 loads). The
// weight gradient dw4 = Y3^T ds is accumulated in registers over all rows a warp visits, then reduced once per CTA in
// shared memory and once per CTA in global memory (the per-element shared-memory atomics of the first version cost 45 us).
__global__ void __launch_bounds__(HEAD_THREADS)
disc_head_kernel(const __nv_bfloat16* __restrict__ Y3, int ld, int P, int h3, const float* __restrict__ w4, const float* __restrict__ b4,
                 const int32_t* __restrict__ label, float keep, float* __restrict__ y_out, float* __restrict__ scal,
                 __nv_bfloat16* __restrict__ dz3, float* __restrict__ dw4, float* __restrict__ db4) {
  __shared__ float s_dw4[HEAD_MAXH3];
  __shared__ float s_acc[4];  // d_loss, sum_y, db4, cnt
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool bwd = dz3 != nullptr;
  const bool drop = keep > 0.f && keep < 1.f;
  for (int j = tid; j < HEAD_MAXH3; j += HEAD_THREADS) s_dw4[j] = 0.f;
  if (tid < 4) s_acc[tid] = 0.f;
  __syncthreads();
  float2 wv[HEAD_PPL], gw[HEAD_PPL];
#pragma unroll
  for (int k = 0; k < HEAD_PPL; ++k) {
    const int j = 2 * lane + 64 * k;
    wv[k].x = j < h3 ? __ldg(w4 + j) : 0.f;
    wv[k].y = j + 1 < h3 ? __ldg(w4 + j + 1) : 0.f;
    gw[k] = make_float2(0.f, 0.f);
  }
  const float bias = b4[0];
  float loss = 0.f, sumy = 0.f, sds = 0.f, ngen = 0.f;
  const int warps_total = gridDim.x * (HEAD_THREADS / 32);
  // two rows per iteration: the loads of both rows are issued before either row's reduction (the kernel is latency-bound)
  for (int row0 = blockIdx.x * (HEAD_THREADS / 32) + warp; row0 < P; row0 += 2 * warps_total) {
    const int rows[2] = {row0, row0 + warps_total};
    float2 yv[2][HEAD_PPL];
    int lab[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const bool ok = rows[q] < P;
      lab[q] = ok ? label[rows[q]] : -1;
      const __nv_bfloat16* yr = Y3 + (size_t)(ok ? rows[q] : row0) * ld;   // ld is even and >= h3: pair loads stay inside the row
#pragma unroll
      for (int k = 0; k < HEAD_PPL; ++k) {
        const int j = 2 * lane + 64 * k;
        yv[q][k] = j < ld ? unpack_bf16x2(*reinterpret_cast<const uint32_t*>(yr + j)) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int row = rows[q];
      if (row >= P) break;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < HEAD_PPL; ++k) { s = fmaf(yv[q][k].x, wv[k].x, s); s = fmaf(yv[q][k].y, wv[k].y, s); }
      s = warp_sum(s) + bias;
      const float y = 1.0f / (1.0f + __expf(-s));
      if (lane == 0 && y_out != nullptr) y_out[row] = y;
      // -log(sigmoid(s)) = softplus(-s) ; -log(1 - sigmoid(s)) = softplus(s)     train.py:142
      // label < 0 (pair dropped by the validity filter, train.py:240-243): no loss, zero gradient row
      const float sp = (lab[q] == 0) ? -s : s;
      const float l = lab[q] < 0 ? 0.f : fmaxf(sp, 0.f) + log1pf(__expf(-fabsf(sp)));
      const float ds = lab[q] < 0 ? 0.f : ((lab[q] == 0) ? (y - 1.0f) : y);
      if (lane == 0) { loss += l; if (lab[q] == 1) { sumy += y; ngen += 1.f; } sds += ds; }
      if (bwd) {
        __nv_bfloat16* dr = dz3 + (size_t)row * ld;
#pragma unroll
        for (int k = 0; k < HEAD_PPL; ++k) {
          const int j = 2 * lane + 64 * k;
          if (j < h3) {  // h3 is even in every configuration this kernel accepts
            const float d0 = ds * wv[k].x * dact_drop_tanh(yv[q][k].x, keep, drop);
            const float d1 = ds * wv[k].y * dact_drop_tanh(yv[q][k].y, keep, drop);
            *reinterpret_cast<uint32_t*>(dr + j) = pack_bf16x2(d0, d1);
            gw[k].x = fmaf(yv[q][k].x, ds, gw[k].x);
            gw[k].y = fmaf(yv[q][k].y, ds, gw[k].y);
          }
        }
      }
    }
  }
  if (lane == 0) { atomicAdd(&s_acc[0], loss); atomicAdd(&s_acc[1], sumy); atomicAdd(&s_acc[2], sds); atomicAdd(&s_acc[3], ngen); }
  if (bwd) {
#pragma unroll
    for (int k = 0; k < HEAD_PPL; ++k) {
      atomicAdd(&s_dw4[2 * lane + 64 * k], gw[k].x);
      atomicAdd(&s_dw4[2 * lane + 64 * k + 1], gw[k].y);
    }
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(scal + LTG_S_D_LOSS, s_acc[0]);
    atomicAdd(scal + LTG_S_SUM_Y, s_acc[1]);
    atomicAdd(scal + LTG_S_CNT, s_acc[3]);
    if (bwd && db4 != nullptr) atomicAdd(db4, s_acc[2]);
  }
  if (bwd && dw4 != nullptr)
    for (int j = tid; j < h3; j += HEAD_THREADS) atomicAdd(dw4 + j, s_dw4[j]);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst, int64_t ld_dst,
                                 int64_t rows, int64_t cols) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = __float2bfloat16(src[r * ld_src + c]);
  }
}

}  // namespace

extern "C" int ltg_disc_gather(const void* E_bf16, const int32_t* pop_ids, const int32_t* niche_ids, int P, void* Xp, void* Xn, void* stream) {
  LTG_REQUIRE(E_bf16 && pop_ids && niche_ids && Xp && Xn);
  if (P <= 0) return LTG_OK;
  const int64_t n = (int64_t)P * 16;
  ltg_launch(disc_gather_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(E_bf16), pop_ids, niche_ids, P,
                                                                                     reinterpret_cast<uint4*>(Xp), reinterpret_cast<uint4*>(Xn));
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_disc_head(const void* Y3_bf16, int ld, int P, int h3, const float* w4, const float* b4, const int32_t* label,
                             float keep, float* y_out, float* scal, void* dz3_bf16, float* dw4, float* db4, void* stream) {
  LTG_REQUIRE(Y3_bf16 && w4 && b4 && label && scal);
  LTG_REQUIRE(h3 > 0 && h3 <= HEAD_MAXH3 && h3 % 2 == 0 && ld >= h3 && ld % 2 == 0);
  if (P <= 0) return LTG_OK;
  int blocks = (P + HEAD_THREADS / 32 - 1) / (HEAD_THREADS / 32);
  if (blocks > 148 * 4) blocks = 148 * 4;
  disc_head_kernel<<<blocks, HEAD_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(Y3_bf16), ld, P, h3, w4, b4, label, keep, y_out, scal, reinterpret_cast<__nv_bfloat16*>(dz3_bf16),
      dw4, db4);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols, void* stream) {
  LTG_REQUIRE(src && dst);
  if (rows <= 0 || cols <= 0) return LTG_OK;
  int64_t blocks = (rows * cols + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
