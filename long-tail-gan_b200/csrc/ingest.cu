// a1 / f3: dataset ingestion on the host cores -- the `pandas.read_csv` + `scipy.sparse.csr_matrix((ones, (rows, cols)))` pair of
// Codes/data_processing.py:6-37 as one multi-threaded pass: the file is memory-mapped, cut at line boundaries into one byte range
// per thread, every thread parses its `row,col` lines into its own pair list, and the CSR is built by a parallel counting sort over
// the rows (atomic row counters -> prefix sum -> scatter -> per-row sort + duplicate merge -> compaction). Duplicated (row, col)
// pairs are summed into the value, exactly what csr_matrix does with repeated coordinates; column ids come out sorted
// within each row (scipy's sort_indices order), so the arrays are bit-identical to the reference's.
//
// Host-only code (no kernels): it lives in libltgan.so because the C-ABI is the package's one native boundary. 57 M-interaction
// files (Netflix shape) are what it is for; the bundled 179 k-row sample parses in a few milliseconds either way.
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

struct CsvPairs {
  std::vector<int64_t> rows, cols;   // file order
  int64_t row_min = 0, row_max = -1, col_min = 0, col_max = -1;
};

struct ChunkOut {
  std::vector<int64_t> rows, cols;
  int64_t row_min = INT64_MAX, row_max = INT64_MIN, col_min = INT64_MAX, col_max = INT64_MIN;
  int64_t bad_line = -1;             // byte offset of the first line that did not parse
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// One field as a non-negative or negative decimal integer ("12", "-3", "12.0" as pandas would read a float column holding integers).
// Returns false if the field is empty or holds anything else.
bool parse_int_field(const char* b, const char* e, int64_t* out) {
  while (b < e && is_space(*b)) ++b;
  while (e > b && is_space(e[-1])) --e;
  if (b < e && *b == '"' && e - b >= 2 && e[-1] == '"') { ++b; --e; }
  if (b >= e) return false;
  bool neg = false;
  if (*b == '-' || *b == '+') { neg = *b == '-'; ++b; }
  if (b >= e || *b < '0' || *b > '9') return false;
  int64_t v = 0;
  while (b < e && *b >= '0' && *b <= '9') { v = v * 10 + (*b - '0'); ++b; }
  if (b < e && *b == '.') {          // "12.0": only zeros may follow
    ++b;
    while (b < e && *b == '0') ++b;
  }
  if (b != e) return false;
  *out = neg ? -v : v;
  return true;
}

// Parses the lines of [b, e) (b at a line start); fields are comma-separated, the row id is field `fr`, the column id field `fc`.
void parse_chunk(const char* base, const char* b, const char* e, int fr, int fc, ChunkOut* out) {
  const int last = fr > fc ? fr : fc;
  out->rows.reserve((size_t)(e - b) / 8 + 16);
  out->cols.reserve((size_t)(e - b) / 8 + 16);
  while (b < e) {
    const char* nl = static_cast<const char*>(memchr(b, '\n', (size_t)(e - b)));
    const char* le = nl != nullptr ? nl : e;
    const char* p = b;
    bool blank = true;
    for (const char* q = b; q < le; ++q) if (!is_space(*q)) { blank = false; break; }
    if (!blank) {                     // (pandas skips blank lines)
      int64_t r = 0, c = 0;
      bool ok_r = false, ok_c = false;
      for (int f = 0; f <= last; ++f) {
        const char* comma = static_cast<const char*>(memchr(p, ',', (size_t)(le - p)));
        const char* fe = comma != nullptr ? comma : le;
        if (f == fr) ok_r = parse_int_field(p, fe, &r);
        if (f == fc) ok_c = parse_int_field(p, fe, &c);
        if (comma == nullptr) break;
        p = comma + 1;
      }
      if (!ok_r || !ok_c) { if (out->bad_line < 0) out->bad_line = (int64_t)(b - base); return; }
      out->rows.push_back(r); out->cols.push_back(c);
      out->row_min = std::min(out->row_min, r); out->row_max = std::max(out->row_max, r);
      out->col_min = std::min(out->col_min, c); out->col_max = std::max(out->col_max, c);
    }
    b = le + 1;
  }
}

int clamp_threads(int n, int64_t work, int64_t per_thread) {
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  if (n <= 0) n = 1;
  const int64_t useful = work / per_thread + 1;
  if (n > useful) n = (int)useful;
  if (n > 256) n = 256;
  return n < 1 ? 1 : n;
}

template <class F>
void run_threads(int n, F f) {
  if (n == 1) { f(0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t)n);
  for (int t = 0; t < n; ++t) th.emplace_back(f, t);
  for (auto& x : th) x.join();
}

std::string trim_field(const char* b, const char* e) {
  while (b < e && is_space(*b)) ++b;
  while (e > b && is_space(e[-1])) --e;
  if (b < e && *b == '"' && e - b >= 2 && e[-1] == '"') { ++b; --e; }
  return std::string(b, e);
}

}  // namespace

extern "C" int ltg_csv_open(const char* path_host, const char* row_name_host, const char* col_name_host, int n_threads,
                            void** handle_host, int64_t* stats_host) {
  LTG_REQUIRE(path_host != nullptr && row_name_host != nullptr && col_name_host != nullptr && handle_host != nullptr);
  *handle_host = nullptr;
  const int fd = open(path_host, O_RDONLY);
  if (fd < 0) { ltg_set_last_error((std::string("cannot open ") + path_host).c_str(), __FILE__, __LINE__); return LTG_ERR_ARG; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); ltg_set_last_error("fstat failed", __FILE__, __LINE__); return LTG_ERR_ARG; }
  const size_t size = (size_t)st.st_size;
  CsvPairs* h = new CsvPairs();
  if (size == 0) { close(fd); delete h; ltg_set_last_error("empty CSV file (no header line)", __FILE__, __LINE__); return LTG_ERR_ARG; }
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) { delete h; ltg_set_last_error("mmap failed", __FILE__, __LINE__); return LTG_ERR_ARG; }
  const char* base = static_cast<const char*>(map);
  const char* end = base + size;
  // header: the positions of the two named columns (tp['uid'], tp['sid'] in data_processing.py:8-15); a UTF-8 byte-order mark in front
  // of it is skipped like pandas does
  const char* hb = (size >= 3 && (unsigned char)base[0] == 0xEF && (unsigned char)base[1] == 0xBB && (unsigned char)base[2] == 0xBF) ? base + 3 : base;
  const char* hnl = static_cast<const char*>(memchr(hb, '\n', (size_t)(end - hb)));
  const char* he = hnl != nullptr ? hnl : end;
  int fr = -1, fc = -1, f = 0;
  for (const char* p = hb; p <= he; ++f) {
    const char* comma = static_cast<const char*>(memchr(p, ',', (size_t)(he - p)));
    const char* fe = comma != nullptr ? comma : he;
    const std::string name = trim_field(p, fe);
    if (name == row_name_host && fr < 0) fr = f;
    if (name == col_name_host && fc < 0) fc = f;
    if (comma == nullptr) break;
    p = comma + 1;
  }
  if (fr < 0 || fc < 0) {
    munmap(map, size); delete h;
    ltg_set_last_error((std::string("CSV header lacks column ") + (fr < 0 ? row_name_host : col_name_host)).c_str(), __FILE__, __LINE__);
    return LTG_ERR_ARG;
  }
  const char* body = hnl != nullptr ? hnl + 1 : end;
  const int64_t nbytes = (int64_t)(end - body);
  const int T = clamp_threads(n_threads, nbytes, 1 << 20);
  // byte ranges that start at line starts
  std::vector<const char*> cut((size_t)T + 1);
  cut[0] = body; cut[(size_t)T] = end;
  for (int t = 1; t < T; ++t) {
    const char* p = body + nbytes * t / T;
    if (p < cut[(size_t)t - 1]) p = cut[(size_t)t - 1];
    const char* nl = p < end ? static_cast<const char*>(memchr(p, '\n', (size_t)(end - p))) : nullptr;
    cut[(size_t)t] = nl != nullptr ? nl + 1 : end;
  }
  std::vector<ChunkOut> outs((size_t)T);
  run_threads(T, [&](int t) { parse_chunk(base, cut[(size_t)t], cut[(size_t)t + 1], fr, fc, &outs[(size_t)t]); });
  for (int t = 0; t < T; ++t)
    if (outs[(size_t)t].bad_line >= 0) {
      char msg[256];
      snprintf(msg, sizeof(msg), "CSV line at byte %lld of %s does not hold two integer ids", (long long)outs[(size_t)t].bad_line, path_host);
      munmap(map, size); delete h;
      ltg_set_last_error(msg, __FILE__, __LINE__);
      return LTG_ERR_ARG;
    }
  size_t total = 0;
  std::vector<size_t> off((size_t)T + 1, 0);
  for (int t = 0; t < T; ++t) { total += outs[(size_t)t].rows.size(); off[(size_t)t + 1] = total; }
  h->rows.resize(total); h->cols.resize(total);
  run_threads(T, [&](int t) {
    const ChunkOut& o = outs[(size_t)t];
    if (!o.rows.empty()) {
      memcpy(h->rows.data() + off[(size_t)t], o.rows.data(), o.rows.size() * sizeof(int64_t));
      memcpy(h->cols.data() + off[(size_t)t], o.cols.data(), o.cols.size() * sizeof(int64_t));
    }
  });
  munmap(map, size);
  if (total > 0) {
    h->row_min = INT64_MAX; h->row_max = INT64_MIN; h->col_min = INT64_MAX; h->col_max = INT64_MIN;
    for (int t = 0; t < T; ++t) {
      const ChunkOut& o = outs[(size_t)t];
      if (o.rows.empty()) continue;
      h->row_min = std::min(h->row_min, o.row_min); h->row_max = std::max(h->row_max, o.row_max);
      h->col_min = std::min(h->col_min, o.col_min); h->col_max = std::max(h->col_max, o.col_max);
    }
  }
  if (stats_host != nullptr) {
    stats_host[0] = (int64_t)total; stats_host[1] = h->row_min; stats_host[2] = h->row_max; stats_host[3] = h->col_min; stats_host[4] = h->col_max;
  }
  *handle_host = h;
  return LTG_OK;
}

extern "C" int ltg_csv_pairs(void* handle_host, int64_t* rows_host, int64_t* cols_host) {
  LTG_REQUIRE(handle_host != nullptr && rows_host != nullptr && cols_host != nullptr);
  const CsvPairs* h = static_cast<const CsvPairs*>(handle_host);
  if (!h->rows.empty()) {
    memcpy(rows_host, h->rows.data(), h->rows.size() * sizeof(int64_t));
    memcpy(cols_host, h->cols.data(), h->cols.size() * sizeof(int64_t));
  }
  return LTG_OK;
}

extern "C" int ltg_csv_to_csr(void* handle_host, int64_t row_offset, int64_t n_rows, int64_t n_cols, int n_threads, int32_t* indptr_host,
                              int32_t* indices_host, float* counts_host, int64_t* nnz_host) {
  LTG_REQUIRE(handle_host != nullptr && indptr_host != nullptr && nnz_host != nullptr && n_rows >= 0 && n_cols >= 0);
  const CsvPairs* h = static_cast<const CsvPairs*>(handle_host);
  const int64_t n = (int64_t)h->rows.size();
  LTG_REQUIRE(n == 0 || (indices_host != nullptr && counts_host != nullptr));
  LTG_REQUIRE(n < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31));
  if (n > 0) {
    LTG_REQUIRE(h->row_min - row_offset >= 0 && h->row_max - row_offset < n_rows);   // every id inside the matrix (scipy raises too)
    LTG_REQUIRE(h->col_min >= 0 && h->col_max < n_cols);
  }
  const int T = clamp_threads(n_threads, n, 1 << 16);
  // 1. pairs per row
  std::vector<std::atomic<int32_t>> cnt((size_t)n_rows + 1);
  run_threads(T, [&](int t) {
    for (int64_t r = n_rows * t / T; r < n_rows * (t + 1) / T; ++r) cnt[(size_t)r].store(0, std::memory_order_relaxed);
  });
  cnt[(size_t)n_rows].store(0, std::memory_order_relaxed);
  run_threads(T, [&](int t) {
    for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) cnt[(size_t)(h->rows[(size_t)i] - row_offset)].fetch_add(1, std::memory_order_relaxed);
  });
  // 2. row starts of the unmerged lists; the counters become scatter cursors
  std::vector<int64_t> start((size_t)n_rows + 1);
  int64_t acc = 0;
  for (int64_t r = 0; r < n_rows; ++r) { start[(size_t)r] = acc; acc += cnt[(size_t)r].load(std::memory_order_relaxed); cnt[(size_t)r].store(0, std::memory_order_relaxed); }
  start[(size_t)n_rows] = acc;
  // 3. scatter the column ids (order inside a row is fixed by the sort below)
  std::vector<int32_t> tmp((size_t)n);
  run_threads(T, [&](int t) {
    for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) {
      const size_t r = (size_t)(h->rows[(size_t)i] - row_offset);
      const int32_t k = cnt[r].fetch_add(1, std::memory_order_relaxed);
      tmp[(size_t)(start[r] + k)] = (int32_t)h->cols[(size_t)i];
    }
  });
  // 4. per row: sort, merge duplicates (count -> value); rows are handed out in blocks through one shared cursor
  std::vector<int32_t> uniq((size_t)n_rows + 1, 0);
  std::vector<float> val((size_t)n);
  std::atomic<int64_t> next(0);
  const int64_t blk = 256;
  run_threads(T, [&](int) {
    for (;;) {
      const int64_t r0 = next.fetch_add(blk, std::memory_order_relaxed);
      if (r0 >= n_rows) break;
      const int64_t r1 = std::min(n_rows, r0 + blk);
      for (int64_t r = r0; r < r1; ++r) {
        int32_t* b = tmp.data() + start[(size_t)r];
        int32_t* e = tmp.data() + start[(size_t)r + 1];
        std::sort(b, e);
        float* v = val.data() + start[(size_t)r];
        int32_t m = 0;
        for (int32_t* p = b; p < e;) {
          int32_t* q = p + 1;
          while (q < e && *q == *p) ++q;
          b[m] = *p; v[m] = (float)(q - p); ++m;
          p = q;
        }
        uniq[(size_t)r] = m;
      }
    }
  });
  // 5. final row pointers and compaction
  std::vector<int64_t> fin((size_t)n_rows + 1);
  acc = 0;
  for (int64_t r = 0; r < n_rows; ++r) { fin[(size_t)r] = acc; indptr_host[r] = (int32_t)acc; acc += uniq[(size_t)r]; }
  fin[(size_t)n_rows] = acc; indptr_host[n_rows] = (int32_t)acc;
  *nnz_host = acc;
  run_threads(T, [&](int t) {
    for (int64_t r = n_rows * t / T; r < n_rows * (t + 1) / T; ++r) {
      const int32_t m = uniq[(size_t)r];
      if (m == 0) continue;
      memcpy(indices_host + fin[(size_t)r], tmp.data() + start[(size_t)r], (size_t)m * sizeof(int32_t));
      memcpy(counts_host + fin[(size_t)r], val.data() + start[(size_t)r], (size_t)m * sizeof(float));
    }
  });
  return LTG_OK;
}

extern "C" int ltg_csv_close(void* handle_host) {
  delete static_cast<CsvPairs*>(handle_host);
  return LTG_OK;
}
