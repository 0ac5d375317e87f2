// a9/a10: niche-item sampling and (popular, generated) pair construction on the device.
//
// Reference: sample.py:40-67 draws n_u of the user's C_u candidates without replacement with
// probability proportional to the generator's softmax restricted to the candidates
// (np.random.choice(..., replace=False, p=...)), train.py:230 sorts the drawn ids, train.py:236-238
// pairs each with a uniformly random popular item of the user and train.py:240-243 drops pairs whose
// ids are not in ITEM_FEATURE_DICT.
//
// Successive draws without replacement proportional to p are Plackett-Luce, which is what the
// Gumbel-top-k trick samples: key_c = logit_c + Gumbel_c, keep the n_u largest keys. The softmax
// normaliser (and the reference's renormalisation over the candidates) cancels inside the arg-top-k,
// so the keys are built from the logits directly. The n_u-th largest key is found with a 4-pass
// 8-bit radix select over order-preserving uint32 keys in shared memory; winners are emitted in
// candidate order, which is ascending item id (data_processing.py:220), i.e. already "sorted".
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int SAMP_THREADS = 512;  // heavy users (thousands of candidates) set the kernel's duration: give each CTA 16 warps

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// exclusive prefix sum of one int per thread over the CTA; *total receives the CTA-wide sum. Two barriers.
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protects s_warp against the previous use
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SAMP_THREADS / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) base += c;
    tot += c;
  }
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(SAMP_THREADS)
sample_pairs_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int n_items, int64_t uid0,
                    const int32_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_items, const int32_t* __restrict__ samp_ptr,
                    const int32_t* __restrict__ pop_ptr, const int32_t* __restrict__ pop_items, const uint8_t* __restrict__ item_valid,
                    uint64_t seed, uint32_t step, const uint32_t* __restrict__ step_dev,
                    int32_t* __restrict__ samp_items, int32_t* __restrict__ samp_partner, int32_t* __restrict__ samp_valid,
                    int32_t* __restrict__ cnt_out, const int32_t* __restrict__ user_order, const float* __restrict__ cand_vals) {
  pdl_trigger();
  pdl_wait_cta();
  extern __shared__ uint32_t s_keys[];  // [max_cand]
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_sel[2];          // digit, remaining-k
  __shared__ int s_warp[SAMP_THREADS / 32];
  __shared__ int s_nvalid;

  // heavy users first (user_order sorts by candidate count, descending): the longest CTAs start at t = 0
  const int u = user_order != nullptr ? user_order[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int s0 = samp_ptr[u];
  const int c0 = cand_ptr[u];
  const int C = cand_ptr[u + 1] - c0;
  int n = samp_ptr[u + 1] - s0;
  if (n <= 0) return;
  if (step_dev != nullptr) step += *step_dev;
  const uint64_t ubase = (uint64_t)(uid0 + u) * (uint64_t)n_items;
  const int nslots = n;
  if (n > C) n = C;  // sample.py:51-61 would shrink the draw; cannot happen when own niche items are candidates

  // keys
  const __nv_bfloat16* row = logits != nullptr ? logits + (size_t)u * ld : nullptr;
  for (int c = tid; c < C; c += SAMP_THREADS) {
    const int item = cand_items[c0 + c];
    const uint32_t r = ltg_rand_u32(seed, LTG_STREAM_SAMPLE, step, ubase + (uint64_t)item);
    const float g = -__logf(-__logf(ltg_u01(r)));
    // cand_vals (catalog-sharded layout): the candidates' logits gathered across the item shards, aligned with cand_items
    const float lg = cand_vals != nullptr ? cand_vals[c0 + c] : __bfloat162float(row[item]);
    s_keys[c] = float_to_ordered(lg + g);
  }
  if (tid == 0) s_nvalid = 0;
  __syncthreads();

  // radix select: the n-th largest key (8 bits per pass; the digit search over the 256 bins is done by warp 0)
  uint32_t prefix = 0, pmask = 0;
  uint32_t k = (uint32_t)n;  // 1-based rank among keys matching the prefix, counted from the top
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int c = tid; c < C; c += SAMP_THREADS) {
      const uint32_t key = s_keys[c];
      if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns bins [8l, 8l+8); suffix sums over lanes give the count of keys in higher bins
      uint32_t h[8], mine = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { h[q] = s_hist[8 * lane + q]; mine += h[q]; }
      uint32_t above = 0;  // keys in bins owned by higher lanes
      uint32_t run = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, run, o);
        if (lane + o < 32) run += t;
      }
      above = run - mine;
      // the digit lies in this lane's bins iff above < k <= above + mine
      if (above < k && k <= above + mine) {
        uint32_t acc = above;
        int d = 7;
        for (; d > 0; --d) {
          if (acc + h[d] >= k) break;
          acc += h[d];
        }
        s_sel[0] = (uint32_t)(8 * lane + d);
        s_sel[1] = k - acc;
      }
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    pmask |= 255u << shift;
    k = s_sel[1];
  }
  const uint32_t T = prefix;   // threshold key; `k` of the keys equal to T are taken (first in candidate order)
  const int need_eq = (int)k;

  // ordered compaction: every thread owns a contiguous segment of the candidate list
  const int seg = (C + SAMP_THREADS - 1) / SAMP_THREADS;
  const int a0 = min(C, tid * seg), a1 = min(C, a0 + seg);
  int n_eq = 0;
  for (int c = a0; c < a1; ++c) n_eq += (s_keys[c] == T) ? 1 : 0;
  int tot_eq;
  int eq_rank = block_exclusive_scan(n_eq, s_warp, &tot_eq);
  int n_take = 0;
  {
    int er = eq_rank;
    for (int c = a0; c < a1; ++c) {
      const uint32_t key = s_keys[c];
      if (key > T) ++n_take;
      else if (key == T) { if (er < need_eq) ++n_take; ++er; }
    }
  }
  int tot_take;
  int off = block_exclusive_scan(n_take, s_warp, &tot_take);
  const int p0 = pop_ptr[u];
  const int np = pop_ptr[u + 1] - p0;
  int nvalid = 0;
  for (int c = a0; c < a1; ++c) {
    const uint32_t key = s_keys[c];
    bool take = key > T;
    if (key == T) { take = eq_rank < need_eq; ++eq_rank; }
    if (!take) continue;
    const int slot = s0 + off++;
    const int item = cand_items[c0 + c];
    int partner = -1;
    if (np > 0) {
      const uint32_t r = ltg_rand_u32(seed, LTG_STREAM_PARTNER, step, ubase + (uint64_t)item);
      partner = pop_items[p0 + (int)(((uint64_t)r * (uint64_t)np) >> 32)];  // train.py:236-238
    }
    const int ok = (partner >= 0 && item_valid[item] && item_valid[partner]) ? 1 : 0;  // train.py:240-243
    samp_items[slot] = item;
    samp_partner[slot] = partner >= 0 ? partner : 0;
    samp_valid[slot] = ok ? 1 : -1;
    nvalid += ok;
  }
  nvalid = (int)warp_sum((float)nvalid);
  if (lane == 0 && nvalid > 0) atomicAdd(&s_nvalid, nvalid);
  // slots that could not be filled (n < nslots)
  for (int j = n + tid; j < nslots; j += SAMP_THREADS) {
    samp_items[s0 + j] = 0; samp_partner[s0 + j] = 0; samp_valid[s0 + j] = -1;
  }
  __syncthreads();
  if (tid == 0 && s_nvalid > 0 && cnt_out != nullptr) atomicAdd(cnt_out, s_nvalid);
}

}  // namespace

extern "C" int ltg_sample_pairs(const void* logits_bf16, int ld_logits, int B, int n_items, int64_t uid0,
                                const int32_t* cand_ptr, const int32_t* cand_items, const int32_t* samp_ptr,
                                const int32_t* pop_ptr, const int32_t* pop_items, const uint8_t* item_valid,
                                uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                int32_t* samp_items, int32_t* samp_partner, int32_t* samp_valid, int32_t* cnt, int max_cand, const int32_t* user_order,
                                void* stream) {
  return ltg_sample_pairs_vals(logits_bf16, ld_logits, nullptr, B, n_items, uid0, cand_ptr, cand_items, samp_ptr, pop_ptr, pop_items, item_valid, seed,
                               step, step_dev, samp_items, samp_partner, samp_valid, cnt, max_cand, user_order, stream);
}

extern "C" int ltg_sample_pairs_vals(const void* logits_bf16, int ld_logits, const float* cand_vals, int B, int n_items, int64_t uid0,
                                     const int32_t* cand_ptr, const int32_t* cand_items, const int32_t* samp_ptr,
                                     const int32_t* pop_ptr, const int32_t* pop_items, const uint8_t* item_valid,
                                     uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                     int32_t* samp_items, int32_t* samp_partner, int32_t* samp_valid, int32_t* cnt, int max_cand,
                                     const int32_t* user_order, void* stream) {
  LTG_REQUIRE((logits_bf16 || cand_vals) && cand_ptr && cand_items && samp_ptr && pop_ptr && pop_items && item_valid);
  LTG_REQUIRE(samp_items && samp_partner && samp_valid);
  LTG_REQUIRE(max_cand >= 0 && (size_t)max_cand * 4 <= 200 * 1024);
  if (B <= 0) return LTG_OK;
  const size_t smem = (size_t)(max_cand > 0 ? max_cand : 1) * sizeof(uint32_t);
  static size_t smem_opted = 40 * 1024;  // static smem of the kernel takes ~1.2 KB of the default 48 KB
  if (smem > smem_opted) {
    cudaError_t e = cudaFuncSetAttribute(sample_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
    smem_opted = smem;
  }
  ltg_launch(sample_pairs_kernel, dim3(B), dim3(SAMP_THREADS), smem, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld_logits, n_items, uid0, cand_ptr, cand_items, samp_ptr, pop_ptr, pop_items,
      item_valid, seed, step, step_dev, samp_items, samp_partner, samp_valid, cnt, user_order, cand_vals);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
