// f1: the GAN side tables of Codes/data_processing.py:170-271 on the host cores.
//
//   load_items_to_sample (170-224)  per user: own niche items + the top max(2n, 10-n) OTHER niche items ranked by their best
//                                   overlap coefficient with any of the user's niche items
//   load_vectors (227-271)          per user and niche item: the user's popular item with the highest overlap coefficient
//
// overlap(a, b) = |U_a AND U_b| / min(|U_a|, |U_b|) (data_processing.py:100-107) is read from the sparse co-occurrence counts
// C = X^T X (CSR, sorted columns) and the item degrees deg = diag(C): the same float64 division the reference performs, so every
// comparison below sees the reference's numbers bit for bit. The reference walks Python dicts of dicts (O(U * n_niche * n_u)); the
// package's NumPy formulation of the same loops takes 12.6 ms + 2.0 ms per user at the ML-20M shape (33 minutes for 136,677 users,
// one thread). Users are independent: here every thread takes blocks of users through one shared cursor and keeps a dense
// best[] row of its own that it resets by the list of touched columns.
//
// Host-only code (no kernels), exported through the same C-ABI as the device path.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

struct CoocView {
  const int64_t* indptr;      // [n_items + 1]
  const int32_t* idx32;       // column ids, sorted within a row (one of the two is set)
  const int64_t* idx64;
  const int64_t* counts;      // co-occurrence counts
  const double* deg;          // [n_items] = diag(C)
  int n_items;
  int64_t col(int64_t e) const { return idx32 != nullptr ? (int64_t)idx32[e] : idx64[e]; }
};

inline double coef(const CoocView& C, int64_t a, int64_t j, int64_t count) {
  const double da = C.deg[a], dj = C.deg[j];
  const double denom = da < dj ? da : dj;
  return denom > 0.0 ? (double)count / denom : 0.0;       // (0 where an item never occurs: the reference has no entry for it)
}

// C[a, j] by binary search in row a (0 if absent)
inline int64_t count_at(const CoocView& C, int64_t a, int64_t j) {
  int64_t lo = C.indptr[a], hi = C.indptr[a + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t c = C.col(mid);
    if (c == j) return C.counts[mid];
    if (c < j) lo = mid + 1; else hi = mid;
  }
  return 0;
}

int n_workers(int n_threads, int64_t n_users) {
  int n = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (n <= 0) n = 1;
  if (n > 256) n = 256;
  if ((int64_t)n > n_users) n = (int)(n_users > 0 ? n_users : 1);
  return n;
}

template <class F>
void for_user_blocks(int n_threads, int64_t n_users, F body) {
  const int T = n_workers(n_threads, n_users);
  std::atomic<int64_t> next(0);
  const int64_t blk = 64;
  auto work = [&](int t) {
    for (;;) {
      const int64_t u0 = next.fetch_add(blk, std::memory_order_relaxed);
      if (u0 >= n_users) break;
      body(t, u0, std::min(n_users, u0 + blk));
    }
  };
  if (T == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" int ltg_cand_sets(const int64_t* C_indptr_host, const void* C_indices_host, int indices_are_64, const int64_t* C_counts_host,
                             const double* deg_host, int n_items, const int32_t* niche_sorted_host, int n_niche,
                             const int64_t* un_ptr_host, const int32_t* un_items_host, const uint8_t* eligible_host, int64_t n_users,
                             int n_threads, const int64_t* out_ptr_host, int32_t* out_items_host, int32_t* out_count_host) {
  LTG_REQUIRE(C_indptr_host && C_indices_host && C_counts_host && deg_host && niche_sorted_host && un_ptr_host && un_items_host);
  LTG_REQUIRE(eligible_host && out_ptr_host && out_items_host && out_count_host && n_items > 0 && n_niche >= 0 && n_users >= 0);
  for (int i = 0; i < n_niche; ++i) {
    LTG_REQUIRE(niche_sorted_host[i] >= 0 && niche_sorted_host[i] < n_items);
    LTG_REQUIRE(i == 0 || niche_sorted_host[i] > niche_sorted_host[i - 1]);
  }
  CoocView C = {C_indptr_host, indices_are_64 ? nullptr : static_cast<const int32_t*>(C_indices_host),
                indices_are_64 ? static_cast<const int64_t*>(C_indices_host) : nullptr, C_counts_host, deg_host, n_items};
  std::atomic<int> bad(0);
  for_user_blocks(n_threads, n_users, [&](int, int64_t u0, int64_t u1) {
    std::vector<double> best((size_t)n_items, 0.0);
    std::vector<uint8_t> mine((size_t)n_items, 0);
    std::vector<int32_t> touched, others;
    for (int64_t u = u0; u < u1; ++u) {
      out_count_host[u] = 0;
      if (!eligible_host[u]) continue;
      const int32_t* cur = un_items_host + un_ptr_host[u];
      const int64_t n = un_ptr_host[u + 1] - un_ptr_host[u];
      int32_t* out = out_items_host + out_ptr_host[u];
      const int64_t cap = out_ptr_host[u + 1] - out_ptr_host[u];
      touched.clear();
      bool ok = true;
      for (int64_t k = 0; k < n; ++k) {
        const int64_t a = cur[k];
        if (a < 0 || a >= n_items) { ok = false; break; }
        mine[(size_t)a] = 1;
        for (int64_t e = C.indptr[a]; e < C.indptr[a + 1]; ++e) {        // best[j] = max over the user's niche items a of overlap(a, j)
          const int64_t j = C.col(e);
          const double v = coef(C, a, j, C.counts[e]);
          if (v > best[(size_t)j]) {
            if (best[(size_t)j] == 0.0) touched.push_back((int32_t)j);
            best[(size_t)j] = v;
          }
        }
      }
      if (!ok) { bad.store(1); for (int64_t k = 0; k < n; ++k) if (cur[k] >= 0 && cur[k] < n_items) mine[(size_t)cur[k]] = 0; continue; }
      others.clear();
      for (int i = 0; i < n_niche; ++i) if (!mine[(size_t)niche_sorted_host[i]]) others.push_back(niche_sorted_host[i]);
      const int64_t want = std::max<int64_t>(2 * n, 10 - n);                // data_processing.py:182
      const int64_t take = std::min<int64_t>(want, (int64_t)others.size());
      // descending coefficient, ties by ascending item id: the stable argsort of -best over the ascending `others`
      auto before = [&](int32_t x, int32_t y) { return best[(size_t)x] > best[(size_t)y] || (best[(size_t)x] == best[(size_t)y] && x < y); };
      if (take > 0 && take < (int64_t)others.size()) std::partial_sort(others.begin(), others.begin() + take, others.end(), before);
      if (n + take > cap) { bad.store(2); } else {
        for (int64_t k = 0; k < n; ++k) out[k] = cur[k];
        for (int64_t k = 0; k < take; ++k) out[n + k] = others[(size_t)k];
        std::sort(out, out + n + take);
        out_count_host[u] = (int32_t)(n + take);
      }
      for (int32_t j : touched) best[(size_t)j] = 0.0;
      for (int64_t k = 0; k < n; ++k) mine[(size_t)cur[k]] = 0;
    }
  });
  if (bad.load() == 1) { ltg_set_last_error("a user's niche item id is outside [0, n_items)", __FILE__, __LINE__); return LTG_ERR_ARG; }
  if (bad.load() == 2) { ltg_set_last_error("candidate buffer too small for a user (out_ptr must reserve n + max(2n, 10-n))", __FILE__, __LINE__); return LTG_ERR_ARG; }
  return LTG_OK;
}

extern "C" int ltg_real_pairs(const int64_t* C_indptr_host, const void* C_indices_host, int indices_are_64, const int64_t* C_counts_host,
                              const double* deg_host, int n_items, const uint8_t* item_valid_host, const int64_t* un_ptr_host,
                              const int32_t* un_items_host, const int64_t* up_ptr_host, const int32_t* up_items_host,
                              const uint8_t* eligible_host, int64_t n_users, int n_threads, int32_t* out_niche_host, int32_t* out_pop_host,
                              int32_t* out_count_host) {
  LTG_REQUIRE(C_indptr_host && C_indices_host && C_counts_host && deg_host && item_valid_host && un_ptr_host && un_items_host);
  LTG_REQUIRE(up_ptr_host && up_items_host && eligible_host && out_niche_host && out_pop_host && out_count_host && n_items > 0 && n_users >= 0);
  CoocView C = {C_indptr_host, indices_are_64 ? nullptr : static_cast<const int32_t*>(C_indices_host),
                indices_are_64 ? static_cast<const int64_t*>(C_indices_host) : nullptr, C_counts_host, deg_host, n_items};
  std::atomic<int> bad(0);
  for_user_blocks(n_threads, n_users, [&](int, int64_t u0, int64_t u1) {
    for (int64_t u = u0; u < u1; ++u) {
      out_count_host[u] = 0;
      if (!eligible_host[u]) continue;
      const int32_t* niches = un_items_host + un_ptr_host[u];
      const int64_t nn = un_ptr_host[u + 1] - un_ptr_host[u];
      const int32_t* pops = up_items_host + up_ptr_host[u];
      const int64_t np_ = up_ptr_host[u + 1] - up_ptr_host[u];
      int32_t* on = out_niche_host + un_ptr_host[u];
      int32_t* op = out_pop_host + un_ptr_host[u];
      if (np_ <= 0) { bad.store(1); continue; }
      int32_t m = 0;
      for (int64_t k = 0; k < nn; ++k) {
        const int64_t a = niches[k];
        if (a < 0 || a >= n_items) { bad.store(1); break; }
        // first maximum in list order (np.argmax over the row of the coefficient block)
        double bv = -1.0; int32_t bp = pops[0];
        for (int64_t q = 0; q < np_; ++q) {
          const int64_t p = pops[q];
          if (p < 0 || p >= n_items) { bad.store(1); break; }
          const double v = coef(C, a, p, count_at(C, a, p));
          if (v > bv) { bv = v; bp = (int32_t)p; }
        }
        if (item_valid_host[a] && item_valid_host[bp]) { on[m] = (int32_t)a; op[m] = bp; ++m; }   // data_processing.py:258-262
      }
      out_count_host[u] = m;
    }
  });
  if (bad.load() != 0) { ltg_set_last_error("an eligible user has no popular item or an item id outside [0, n_items)", __FILE__, __LINE__); return LTG_ERR_ARG; }
  return LTG_OK;
}
