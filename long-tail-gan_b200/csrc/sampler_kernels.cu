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

__global__ void __launch_bounds__(SAMP_THREADS)
sample_pairs_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int n_items, int64_t uid0,
                    const int32_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_items, const int32_t* __restrict__ samp_ptr,
                    const int32_t* __restrict__ pop_ptr, const int32_t* __restrict__ pop_items, const uint8_t* __restrict__ item_valid,
                    uint64_t seed, uint32_t step, const uint32_t* __restrict__ step_dev,
                    int32_t* __restrict__ samp_items, int32_t* __restrict__ samp_partner, int32_t* __restrict__ samp_valid,
                    int32_t* __restrict__ cnt_out, const int32_t* __restrict__ user_order) {
  extern __shared__ uint32_t s_keys[];  // [max_cand]
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_sel[2];          // digit, remaining-k
  __shared__ int s_warp[SAMP_THREADS / 32 + 1];
  __shared__ int s_base, s_eq_base, s_nvalid;

  // heavy users first (user_order sorts by candidate count, descending): the longest CTAs start at t = 0
  const int u = user_order != nullptr ? user_order[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x;
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
  const __nv_bfloat16* row = logits + (size_t)u * ld;
  for (int c = tid; c < C; c += SAMP_THREADS) {
    const int item = cand_items[c0 + c];
    const uint32_t r = ltg_rand_u32(seed, LTG_STREAM_SAMPLE, step, ubase + (uint64_t)item);
    const float g = -__logf(-__logf(ltg_u01(r)));
    s_keys[c] = float_to_ordered(__bfloat162float(row[item]) + g);
  }
  if (tid == 0) { s_nvalid = 0; }
  __syncthreads();

  // radix select: the n-th largest key
  uint32_t prefix = 0, pmask = 0;
  uint32_t k = (uint32_t)n;  // 1-based rank among keys matching the prefix, counted from the top
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += SAMP_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int c = tid; c < C; c += SAMP_THREADS) {
      const uint32_t key = s_keys[c];
      if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0; int d = 255;
      for (; d > 0; --d) {
        if (acc + s_hist[d] >= k) break;
        acc += s_hist[d];
      }
      s_sel[0] = (uint32_t)d;
      s_sel[1] = k - acc;
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    pmask |= 255u << shift;
    k = s_sel[1];
    __syncthreads();
  }
  const uint32_t T = prefix;   // threshold key; `k` of the keys equal to T are taken (first in candidate order)
  const int need_eq = (int)k;

  // ordered compaction
  if (tid == 0) { s_base = 0; s_eq_base = 0; }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int cb = 0; cb < C; cb += SAMP_THREADS) {
    const int c = cb + tid;
    const uint32_t key = c < C ? s_keys[c] : 0u;
    const bool gt = c < C && key > T;
    const bool eq = c < C && key == T;
    // rank of this thread's equal-key among equals (block order)
    const uint32_t eq_ballot = __ballot_sync(0xffffffffu, eq);
    const int eq_before_w = __popc(eq_ballot & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(eq_ballot);
    __syncthreads();
    int eq_off = s_eq_base;
    for (int w = 0; w < warp; ++w) eq_off += s_warp[w];
    const bool take = gt || (eq && (eq_off + eq_before_w) < need_eq);
    int eq_total = 0;
    for (int w = 0; w < SAMP_THREADS / 32; ++w) eq_total += s_warp[w];
    __syncthreads();
    const uint32_t tk_ballot = __ballot_sync(0xffffffffu, take);
    const int tk_before_w = __popc(tk_ballot & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(tk_ballot);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    int tk_total = 0;
    for (int w = 0; w < SAMP_THREADS / 32; ++w) tk_total += s_warp[w];
    if (take) {
      const int slot = s0 + off + tk_before_w;
      const int item = cand_items[c0 + c];
      const int p0 = pop_ptr[u];
      const int np = pop_ptr[u + 1] - p0;
      int partner = -1;
      if (np > 0) {
        const uint32_t r = ltg_rand_u32(seed, LTG_STREAM_PARTNER, step, ubase + (uint64_t)item);
        partner = pop_items[p0 + (int)(((uint64_t)r * (uint64_t)np) >> 32)];  // train.py:236-238
      }
      const int ok = (partner >= 0 && item_valid[item] && item_valid[partner]) ? 1 : 0;  // train.py:240-243
      samp_items[slot] = item;
      samp_partner[slot] = partner >= 0 ? partner : 0;
      samp_valid[slot] = ok ? 1 : -1;
      if (ok) atomicAdd(&s_nvalid, 1);
    }
    __syncthreads();
    if (tid == 0) { s_base += tk_total; s_eq_base += eq_total; }
    __syncthreads();
  }
  // slots that could not be filled (n < nslots)
  for (int j = n + tid; j < nslots; j += SAMP_THREADS) {
    samp_items[s0 + j] = 0; samp_partner[s0 + j] = 0; samp_valid[s0 + j] = -1;
  }
  if (tid == 0 && s_nvalid > 0 && cnt_out != nullptr) atomicAdd(cnt_out, s_nvalid);
}

}  // namespace

extern "C" int ltg_sample_pairs(const void* logits_bf16, int ld_logits, int B, int n_items, int64_t uid0,
                                const int32_t* cand_ptr, const int32_t* cand_items, const int32_t* samp_ptr,
                                const int32_t* pop_ptr, const int32_t* pop_items, const uint8_t* item_valid,
                                uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                int32_t* samp_items, int32_t* samp_partner, int32_t* samp_valid, int32_t* cnt, int max_cand, const int32_t* user_order,
                                void* stream) {
  LTG_REQUIRE(logits_bf16 && cand_ptr && cand_items && samp_ptr && pop_ptr && pop_items && item_valid);
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
  sample_pairs_kernel<<<B, SAMP_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld_logits, n_items, uid0, cand_ptr, cand_items, samp_ptr, pop_ptr, pop_items,
      item_valid, seed, step, step_dev, samp_items, samp_partner, samp_valid, cnt, user_order);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
