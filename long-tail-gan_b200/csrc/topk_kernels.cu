// a16/a17: exact per-row top-k and the ranking metrics built on it
// (eval_functions.py:11-38 NDCG@k, 40-62 Recall@k; the seen-item mask of train.py:341 / test.py:149).
//
// One CTA per user row. The row's seen items become a bitmap in shared memory (one bit per catalog
// item), scores are mapped to order-preserving uint32 keys, and the k-th largest key is found with a
// 4-pass 8-bit radix select that streams the row from L2/HBM (nothing but the bitmap is staged, so
// the same kernel serves a 1,000-item and a 1,000,000-item catalog). Ties at the threshold are
// resolved by lowest item index with a second radix select over the indices, which runs only when
// the threshold key is actually shared. The k winners are then sorted (bitonic, 128 lanes) by
// (score desc, index asc) and tested against the held-out CSR row by binary search.
#include <stdlib.h>
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 128;

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <bool BF16>
__device__ __forceinline__ uint32_t load_key(const void* row, int i, const uint32_t* bitmap) {
  if (bitmap[i >> 5] & (1u << (i & 31))) return f2ord(-INFINITY);
  float f;
  if (BF16) f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[i]);
  else f = reinterpret_cast<const float*>(row)[i];
  return f2ord(f);
}

// Finds, among elements whose `key(i)` passes `pred`, the value with 1-based rank k from the top of
// an arbitrary 32-bit quantity `val(i)`. Returns the value; *count_eq receives how many elements share it
// and *k_rem the rank remaining inside that group.
template <class ValFn>
__device__ uint32_t radix_select_desc(int n, uint32_t k, ValFn val, uint32_t* s_hist, uint32_t* s_sel, uint32_t* count_eq, uint32_t* k_rem) {
  uint32_t prefix = 0, pmask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += TK_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += TK_THREADS) {
      uint32_t v; const bool ok = val(i, v);
      if (ok && (v & pmask) == prefix) atomicAdd(&s_hist[(v >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0; int d = 255;
      for (; d > 0; --d) {
        if (acc + s_hist[d] >= k) break;
        acc += s_hist[d];
      }
      s_sel[0] = (uint32_t)d; s_sel[1] = k - acc; s_sel[2] = s_hist[d];
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    pmask |= 255u << shift;
    k = s_sel[1];
    *count_eq = s_sel[2];
    __syncthreads();
  }
  *k_rem = k;
  return prefix;
}

template <bool BF16>
__global__ void __launch_bounds__(TK_THREADS)
topk_metrics_kernel(const void* __restrict__ scores, int64_t ld, int n_items,
                    const int32_t* __restrict__ seen_ptr, const int32_t* __restrict__ seen_items,
                    const int32_t* __restrict__ held_ptr, const int32_t* __restrict__ held_items,
                    int k, int4 rk, int n_rk, int32_t* __restrict__ topk_idx, double* __restrict__ dcg, int32_t* __restrict__ hits) {
  extern __shared__ uint32_t s_bitmap[];  // [(n_items+31)/32]
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_sel[3];
  __shared__ unsigned long long s_win[TK_MAXK];
  __shared__ int s_nwin;
  __shared__ double s_dcg[TK_MAXK / 32];
  __shared__ int s_hits[4];

  const int u = blockIdx.x;
  const int tid = threadIdx.x;
  const void* row = BF16 ? (const void*)(reinterpret_cast<const __nv_bfloat16*>(scores) + (size_t)u * ld)
                         : (const void*)(reinterpret_cast<const float*>(scores) + (size_t)u * ld);
  const int words = (n_items + 31) >> 5;
  for (int w = tid; w < words; w += TK_THREADS) s_bitmap[w] = 0u;
  if (tid < TK_MAXK) s_win[tid] = 0ull;
  if (tid == 0) s_nwin = 0;
  if (tid < 4) s_hits[tid] = 0;
  __syncthreads();
  if (seen_ptr != nullptr)
    for (int j = seen_ptr[u] + tid; j < seen_ptr[u + 1]; j += TK_THREADS) {
      const int it = seen_items[j];
      atomicOr(&s_bitmap[it >> 5], 1u << (it & 31));
    }
  __syncthreads();

  const int kk = min(k, n_items);
  uint32_t cnt_eq = 0, k_rem = 0;
  const uint32_t T = radix_select_desc(n_items, (uint32_t)kk,
      [&](int i, uint32_t& v) { v = load_key<BF16>(row, i, s_bitmap); return true; }, s_hist, s_sel, &cnt_eq, &k_rem);
  // ties at the threshold: keep the k_rem lowest indices among the cnt_eq elements equal to T
  uint32_t idx_thr = 0xFFFFFFFFu;  // compare on ~idx: take eq elements with ~idx >= idx_thr_inv
  uint32_t inv_thr = 0u;
  if (cnt_eq != k_rem) {
    uint32_t c2 = 0, k2 = 0;
    inv_thr = radix_select_desc(n_items, k_rem,
        [&](int i, uint32_t& v) { v = ~(uint32_t)i; return load_key<BF16>(row, i, s_bitmap) == T; }, s_hist, s_sel, &c2, &k2);
  }
  (void)idx_thr;
  // collect winners (unordered), then sort
  for (int i = tid; i < n_items; i += TK_THREADS) {
    const uint32_t key = load_key<BF16>(row, i, s_bitmap);
    if (key > T || (key == T && ~(uint32_t)i >= inv_thr)) {
      const int slot = atomicAdd(&s_nwin, 1);
      if (slot < TK_MAXK) s_win[slot] = ((unsigned long long)key << 32) | (unsigned long long)(~(uint32_t)i);
    }
  }
  __syncthreads();
  // bitonic sort, descending, 128 entries (unused entries are 0 = smallest)
  for (int size = 2; size <= TK_MAXK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < TK_MAXK) {
        const int j = tid ^ stride;
        if (j > tid) {
          const unsigned long long a = s_win[tid], b = s_win[j];
          const bool desc = (tid & size) == 0;
          if ((a < b) == desc) { s_win[tid] = b; s_win[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  // metrics
  double term = 0.0;
  if (tid < kk) {
    const int idx = (int)(~(uint32_t)(s_win[tid] & 0xFFFFFFFFull));
    if (topk_idx != nullptr) topk_idx[(size_t)u * k + tid] = idx;
    int lo = held_ptr[u], hi = held_ptr[u + 1];
    bool hit = false;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = held_items[mid];
      if (v == idx) { hit = true; break; }
      if (v < idx) lo = mid + 1; else hi = mid;
    }
    if (hit) {
      term = 1.0 / log2((double)(tid + 2));  // eval_functions.py:25
      if (n_rk > 0 && tid < rk.x) atomicAdd(&s_hits[0], 1);
      if (n_rk > 1 && tid < rk.y) atomicAdd(&s_hits[1], 1);
      if (n_rk > 2 && tid < rk.z) atomicAdd(&s_hits[2], 1);
      if (n_rk > 3 && tid < rk.w) atomicAdd(&s_hits[3], 1);
    }
  } else if (tid < k && topk_idx != nullptr) {
    topk_idx[(size_t)u * k + tid] = -1;
  }
  if (tid < TK_MAXK) {
    term = warp_sum_d(term);
    if ((tid & 31) == 0) s_dcg[tid >> 5] = term;
  }
  __syncthreads();
  if (tid == 0) {
    dcg[u] = s_dcg[0] + s_dcg[1] + s_dcg[2] + s_dcg[3];
    for (int j = 0; j < n_rk; ++j) hits[(size_t)u * n_rk + j] = s_hits[j];
  }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Staged variant (catalogs whose row of keys fits in shared memory, <= 49,152 items): the row is read from HBM/L2 exactly ONCE with
// 16-byte loads, converted to order-preserving keys in shared memory, seen items are overwritten with -inf there, and the selection
// runs on shared memory: a 3-pass radix select with 11/11/10-bit digits whose 2048-bin histogram is scanned by the whole CTA (four
// bins per thread + a block-wide suffix sum), not by one thread. Round 1's streaming kernel above re-read the row 5-6 times with
// scalar loads and walked 256 bins serially four times: 0.06 of HBM peak. It remains the path for catalogs
// that do not fit (1 M items).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int TS_THREADS = 512;
constexpr int TS_BINS = 2048;
constexpr int TS_MAX_ITEMS = 49152;
constexpr int TS_LD = 5;               // 16-byte loads in flight per thread while the row is staged
constexpr int TS_CAND = 2048;          // candidate list of the pre-filter (key << 32 | ~index)

// block-wide: given per-thread `mine`, returns the sum over all threads with a HIGHER thread index (exclusive suffix sum)
__device__ __forceinline__ uint32_t block_suffix_excl(uint32_t mine, uint32_t* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t run = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_down_sync(0xffffffffu, run, o);
    if (lane + o < 32) run += t;
  }
  __syncthreads();                       // s_warp may still be read from the previous use
  if (lane == 0) s_warp[warp] = run;     // warp total
  __syncthreads();
  uint32_t higher = 0;
#pragma unroll
  for (int w = 0; w < TS_THREADS / 32; ++w) higher += (w > warp) ? s_warp[w] : 0u;
  return higher + run - mine;
}

// k-th largest (1-based) of val(i) over the elements for which val returns true; digits of 11, 11 and 10 bits
template <class ValFn>
__device__ uint32_t radix_select11(int n, uint32_t k, ValFn val, uint32_t* s_hist, uint32_t* s_warp, uint32_t* s_sel, uint32_t* count_eq,
                                   uint32_t* k_rem) {
  uint32_t prefix = 0, pmask = 0;
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const int shift = shifts[p];
    const uint32_t dmask = (1u << bits[p]) - 1u;
    for (int i = threadIdx.x; i < TS_BINS; i += TS_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += TS_THREADS) {
      uint32_t v; const bool ok = val(i, v);
      if (ok && (v & pmask) == prefix) atomicAdd(&s_hist[(v >> shift) & dmask], 1u);
    }
    __syncthreads();
    // thread t owns bins [4t, 4t+4); the digit is in its range iff  above < k <= above + mine
    const uint4 h = *reinterpret_cast<const uint4*>(s_hist + 4 * threadIdx.x);
    const uint32_t mine = h.x + h.y + h.z + h.w;
    const uint32_t above = block_suffix_excl(mine, s_warp);
    if (above < k && k <= above + mine) {
      const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
      uint32_t acc = above;
      int d = 3;
      for (; d > 0; --d) {
        if (acc + hh[d] >= k) break;
        acc += hh[d];
      }
      s_sel[0] = (uint32_t)(4 * threadIdx.x + d); s_sel[1] = k - acc; s_sel[2] = hh[d];
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    pmask |= dmask << shift;
    k = s_sel[1];
    *count_eq = s_sel[2];
    __syncthreads();
  }
  *k_rem = k;
  return prefix;
}

template <bool BF16>
__global__ void __launch_bounds__(TS_THREADS)
topk_staged_kernel(const void* __restrict__ scores, int64_t ld, int n_items,
                   const int32_t* __restrict__ seen_ptr, const int32_t* __restrict__ seen_items,
                   const int32_t* __restrict__ held_ptr, const int32_t* __restrict__ held_items,
                   int k, int4 rk, int n_rk, int32_t* __restrict__ topk_idx, double* __restrict__ dcg, int32_t* __restrict__ hits,
                   bool use_prefilter) {
  extern __shared__ __align__(16) uint32_t s_keys[];   // [n_items rounded up to 8]
  __shared__ __align__(16) uint32_t s_hist[TS_BINS];
  __shared__ uint32_t s_warp[TS_THREADS / 32];
  __shared__ uint32_t s_sel[3];
  __shared__ uint32_t s_lmax[TS_THREADS];
  __shared__ unsigned long long s_cand[TS_CAND];
  __shared__ int s_ncand;
  __shared__ unsigned long long s_win[TK_MAXK];
  __shared__ int s_nwin;
  __shared__ double s_dcg[TK_MAXK / 32];
  __shared__ int s_hits[4];

  const int u = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < TK_MAXK) s_win[tid] = 0ull;
  if (tid == 0) { s_nwin = 0; s_ncand = 0; }
  if (tid < 4) s_hits[tid] = 0;
  // ---- the row, once, 16 bytes per load
  if (BF16) {
    const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(scores) + (size_t)u * ld;
    const int n8 = n_items >> 3;
    for (int v = tid; v < n8; v += TS_THREADS) {
      const uint4 x = ld_nc_v4(reinterpret_cast<const uint4*>(row) + v);
      const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(w[q]);
        s_keys[8 * v + 2 * q] = f2ord(f.x); s_keys[8 * v + 2 * q + 1] = f2ord(f.y);
      }
    }
    for (int i = (n8 << 3) + tid; i < n_items; i += TS_THREADS) s_keys[i] = f2ord(__bfloat162float(row[i]));
  } else {
    const float* row = reinterpret_cast<const float*>(scores) + (size_t)u * ld;
    const int n4 = n_items >> 2;
    // rounds of TS_LD predicated 16-byte loads per thread, all issued before the first use: two memory round trips for a 20 k-item row
    // (an unrolled-by-4 loop with a scalar remainder took four to five)
    for (int v0 = 0; v0 < n4; v0 += TS_LD * TS_THREADS) {
      uint4 x[TS_LD];
#pragma unroll
      for (int q = 0; q < TS_LD; ++q) {
        const int v = v0 + q * TS_THREADS + tid;
        if (v < n4) x[q] = ld_nc_v4(reinterpret_cast<const uint4*>(row) + v);
      }
#pragma unroll
      for (int q = 0; q < TS_LD; ++q) {
        const int v = v0 + q * TS_THREADS + tid;
        if (v < n4) {
          uint4 o;
          o.x = f2ord(__uint_as_float(x[q].x)); o.y = f2ord(__uint_as_float(x[q].y)); o.z = f2ord(__uint_as_float(x[q].z));
          o.w = f2ord(__uint_as_float(x[q].w));
          *reinterpret_cast<uint4*>(s_keys + 4 * v) = o;
        }
      }
    }
    for (int i = (n4 << 2) + tid; i < n_items; i += TS_THREADS) s_keys[i] = f2ord(row[i]);
  }
  __syncthreads();
  // seen items -> -inf (train.py:341: pred[X.nonzero()] = -inf)
  if (seen_ptr != nullptr) {
    const uint32_t ninf = f2ord(-INFINITY);
    for (int j = seen_ptr[u] + tid; j < seen_ptr[u + 1]; j += TS_THREADS) s_keys[seen_items[j]] = ninf;
  }
  __syncthreads();

  const int kk = min(k, n_items);
  // ---- candidate pre-filter. Every thread takes the maximum of its own (strided) elements: 512 DISTINCT elements of the row, so the
  // kk-th largest of them is a lower bound of the row's kk-th largest key and only the elements >= that bound can be in the top kk
  // (about 110 of 20 k for scores without structure). The exact selection below then runs over that short list instead of the row:
  // the 3 histogram passes over all keys -- fp32 logits of one row share sign, exponent and leading mantissa bits, so their first
  // digit lands in a handful of bins and the shared-memory atomics serialise -- shrink to one compare pass. Rows whose list overflows
  // (massive ties, e.g. one score for the whole catalog) take the full-row path: same result either way.
  if (use_prefilter) {
    uint32_t lmax = 0u;   // (key 0 = a negative NaN: below every real score)
    for (int i = tid; i < n_items; i += TS_THREADS) lmax = max(lmax, s_keys[i]);
    s_lmax[tid] = lmax;
    __syncthreads();
    uint32_t c0 = 0, k0 = 0;
    const uint32_t bound = radix_select11(TS_THREADS, (uint32_t)kk, [&](int i, uint32_t& v) { v = s_lmax[i]; return true; }, s_hist, s_warp,
                                          s_sel, &c0, &k0);
    for (int i = tid; i < n_items; i += TS_THREADS) {
      const uint32_t key = s_keys[i];
      if (key >= bound) {
        const int slot = atomicAdd(&s_ncand, 1);
        if (slot < TS_CAND) s_cand[slot] = ((unsigned long long)key << 32) | (unsigned long long)(~(uint32_t)i);
      }
    }
    __syncthreads();
  }
  const bool use_cand = use_prefilter && s_ncand <= TS_CAND;   // uniform over the CTA
  const int n_el = use_cand ? s_ncand : n_items;
  auto key_of = [&](int i) -> uint32_t { return use_cand ? (uint32_t)(s_cand[i] >> 32) : s_keys[i]; };
  auto inv_of = [&](int i) -> uint32_t { return use_cand ? (uint32_t)(s_cand[i] & 0xFFFFFFFFull) : ~(uint32_t)i; };

  uint32_t cnt_eq = 0, k_rem = 0;
  const uint32_t T = radix_select11(n_el, (uint32_t)kk, [&](int i, uint32_t& v) { v = key_of(i); return true; }, s_hist, s_warp, s_sel,
                                    &cnt_eq, &k_rem);
  // ties at the threshold: keep the k_rem lowest indices among the cnt_eq elements equal to T
  uint32_t inv_thr = 0u;
  if (cnt_eq != k_rem) {
    uint32_t c2 = 0, k2 = 0;
    inv_thr = radix_select11(n_el, k_rem, [&](int i, uint32_t& v) { v = inv_of(i); return key_of(i) == T; }, s_hist, s_warp, s_sel, &c2, &k2);
  }
  for (int i = tid; i < n_el; i += TS_THREADS) {
    const uint32_t key = key_of(i), inv = inv_of(i);
    if (key > T || (key == T && inv >= inv_thr)) {
      const int slot = atomicAdd(&s_nwin, 1);
      if (slot < TK_MAXK) s_win[slot] = ((unsigned long long)key << 32) | (unsigned long long)inv;
    }
  }
  __syncthreads();
  // bitonic sort, descending, 128 entries (unused entries are 0 = smallest)
  for (int size = 2; size <= TK_MAXK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < TK_MAXK) {
        const int j = tid ^ stride;
        if (j > tid) {
          const unsigned long long a = s_win[tid], b = s_win[j];
          const bool desc = (tid & size) == 0;
          if ((a < b) == desc) { s_win[tid] = b; s_win[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  // metrics (eval_functions.py:25-30, 47-52)
  double term = 0.0;
  if (tid < kk) {
    const int idx = (int)(~(uint32_t)(s_win[tid] & 0xFFFFFFFFull));
    if (topk_idx != nullptr) topk_idx[(size_t)u * k + tid] = idx;
    int lo = held_ptr[u], hi = held_ptr[u + 1];
    bool hit = false;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = held_items[mid];
      if (v == idx) { hit = true; break; }
      if (v < idx) lo = mid + 1; else hi = mid;
    }
    if (hit) {
      term = 1.0 / log2((double)(tid + 2));
      if (n_rk > 0 && tid < rk.x) atomicAdd(&s_hits[0], 1);
      if (n_rk > 1 && tid < rk.y) atomicAdd(&s_hits[1], 1);
      if (n_rk > 2 && tid < rk.z) atomicAdd(&s_hits[2], 1);
      if (n_rk > 3 && tid < rk.w) atomicAdd(&s_hits[3], 1);
    }
  } else if (tid < k && topk_idx != nullptr) {
    topk_idx[(size_t)u * k + tid] = -1;
  }
  if (tid < TK_MAXK) {
    term = warp_sum_d(term);
    if ((tid & 31) == 0) s_dcg[tid >> 5] = term;
  }
  __syncthreads();
  if (tid == 0) {
    dcg[u] = s_dcg[0] + s_dcg[1] + s_dcg[2] + s_dcg[3];
    for (int j = 0; j < n_rk; ++j) hits[(size_t)u * n_rk + j] = s_hits[j];
  }
}

}  // namespace

extern "C" int ltg_topk_metrics(const void* scores, int is_bf16, int64_t ld, int n_rows, int n_items,
                                const int32_t* seen_ptr, const int32_t* seen_items, const int32_t* held_ptr, const int32_t* held_items,
                                int k, const int32_t* rk_host, int n_rk, int32_t* topk_idx, double* dcg, int32_t* hits, void* stream) {
  LTG_REQUIRE(scores && held_ptr && held_items && dcg);
  LTG_REQUIRE(k >= 1 && k <= TK_MAXK && n_rk >= 0 && n_rk <= 4 && (n_rk == 0 || (rk_host != nullptr && hits != nullptr)));
  LTG_REQUIRE(seen_ptr == nullptr || seen_items != nullptr);
  if (n_rows <= 0) return LTG_OK;
  int4 rk = make_int4(0, 0, 0, 0);
  int* rkp = reinterpret_cast<int*>(&rk);
  for (int j = 0; j < n_rk; ++j) { LTG_REQUIRE(rk_host[j] >= 1 && rk_host[j] <= k); rkp[j] = rk_host[j]; }
  // staged variant: keys of the row in shared memory (16-byte aligned rows required for the vector loads)
  static int use_staged = -1;
  if (use_staged < 0) { const char* e = getenv("LTG_TOPK_STAGED"); use_staged = (e == nullptr || atoi(e) != 0) ? 1 : 0; }
  const bool aligned = (reinterpret_cast<uintptr_t>(scores) & 15) == 0 && ((ld * (is_bf16 ? 2 : 4)) & 15) == 0;
  if (use_staged && n_items <= TS_MAX_ITEMS && aligned) {
    const size_t sm = (size_t)((n_items + 7) / 8 * 8) * 4;
    static size_t opted_s[2] = {16 * 1024, 16 * 1024};   // the kernel has ~28 KB of static shared memory
    static int prefilter = -1;   // LTG_TOPK_PREFILTER=0: select over the whole row (A/B switch)
    if (prefilter < 0) { const char* e = getenv("LTG_TOPK_PREFILTER"); prefilter = (e != nullptr && e[0] == '0') ? 0 : 1; }
    if (sm > opted_s[is_bf16 ? 1 : 0]) {
      cudaError_t e = is_bf16 ? cudaFuncSetAttribute(topk_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
                              : cudaFuncSetAttribute(topk_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
      opted_s[is_bf16 ? 1 : 0] = sm;
    }
    if (is_bf16)
      topk_staged_kernel<true><<<n_rows, TS_THREADS, sm, (cudaStream_t)stream>>>(scores, ld, n_items, seen_ptr, seen_items, held_ptr, held_items, k, rk,
                                                                                 n_rk, topk_idx, dcg, hits, prefilter != 0);
    else
      topk_staged_kernel<false><<<n_rows, TS_THREADS, sm, (cudaStream_t)stream>>>(scores, ld, n_items, seen_ptr, seen_items, held_ptr, held_items, k, rk,
                                                                                  n_rk, topk_idx, dcg, hits, prefilter != 0);
    LTG_CHECK_LAUNCH();
    return LTG_OK;
  }
  const size_t smem = (size_t)((n_items + 31) / 32) * 4;
  LTG_REQUIRE(smem <= 200 * 1024);
  static size_t opted[2] = {40 * 1024, 40 * 1024};
  if (smem > opted[is_bf16 ? 1 : 0]) {
    cudaError_t e = is_bf16 ? cudaFuncSetAttribute(topk_metrics_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(topk_metrics_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
    opted[is_bf16 ? 1 : 0] = smem;
  }
  if (is_bf16)
    topk_metrics_kernel<true><<<n_rows, TK_THREADS, smem, (cudaStream_t)stream>>>(scores, ld, n_items, seen_ptr, seen_items, held_ptr, held_items, k,
                                                                                 rk, n_rk, topk_idx, dcg, hits);
  else
    topk_metrics_kernel<false><<<n_rows, TK_THREADS, smem, (cudaStream_t)stream>>>(scores, ld, n_items, seen_ptr, seen_items, held_ptr, held_items, k,
                                                                                  rk, n_rk, topk_idx, dcg, hits);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
