"""Data layer with the reference's loader names and return conventions (Codes/data_processing.py), rewritten on
NumPy/SciPy sparse algebra so that it scales past the toy catalog: the O(I^2) Python set intersections of
`load_overlap_coeff` (data_processing.py:110-167) become one sparse X^T X, and the per-user Python loops of
`load_items_to_sample` / `load_vectors` (170-271) become row-wise max / arg-max over slices of that matrix.

Every loader returns what the reference returns (dicts keyed by user / item id, scipy CSR matrices) so `train.py` reads
like the reference; `build_train_tables` converts them to the flat int32 CSR-style arrays the device engine consumes.
"""
import codecs

import numpy as np
import pandas as pd
from scipy import sparse


# ---------------------------------------------------------------------------------------------------------------------
# interaction matrices (data_processing.py:6-37)
# ---------------------------------------------------------------------------------------------------------------------
def csr_from_pairs(rows, cols, n_rows, n_cols, dtype):
    """CSR of a binary interaction list: counting sort by (row, col), duplicates summed -- the arrays scipy's
    csr_matrix((ones, (rows, cols))) + sort_indices produce (bit-exact, tests/test_data_processing.py)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    key = rows * n_cols + cols
    uniq, counts = np.unique(key, return_counts=True)
    r = uniq // n_cols
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    return np.cumsum(indptr).astype(np.int32), (uniq % n_cols).astype(np.int32), counts.astype(dtype)


class CsvPairs(object):
    """A `uid,sid` interaction file parsed by the native multi-threaded reader (ltg_csv_open, csrc/ingest.cu): the stand-in for
    `pd.read_csv(path)` followed by `csr_matrix((ones, (rows, cols)))` (data_processing.py:7-15, 21-35) on files with tens of
    millions of rows. Context manager; the parsed pairs live in the library until close()."""

    def __init__(self, path, row_name="uid", col_name="sid", n_threads=0):
        import ctypes
        from . import _lib
        self._lib = _lib
        self._h = ctypes.c_void_p()
        stats = np.zeros(5, dtype=np.int64)
        _lib.check(_lib.load().ltg_csv_open(str(path).encode(), row_name.encode(), col_name.encode(), int(n_threads), ctypes.byref(self._h),
                                            stats.ctypes.data))
        self.n_pairs, self.row_min, self.row_max, self.col_min, self.col_max = (int(x) for x in stats)
        self.n_threads = int(n_threads)

    def pairs(self):
        """(rows, cols) int64 in file order."""
        r = np.empty(self.n_pairs, dtype=np.int64); c = np.empty(self.n_pairs, dtype=np.int64)
        self._lib.check(self._lib.load().ltg_csv_pairs(self._h, r.ctypes.data, c.ctypes.data))
        return r, c

    def to_csr(self, n_rows, n_cols, dtype, row_offset=0):
        """scipy CSR [n_rows, n_cols] of the pairs (rows shifted by -row_offset), duplicates summed, indices sorted."""
        indptr = np.empty(n_rows + 1, dtype=np.int32)
        indices = np.empty(max(1, self.n_pairs), dtype=np.int32)
        counts = np.empty(max(1, self.n_pairs), dtype=np.float32)
        nnz = np.zeros(1, dtype=np.int64)
        self._lib.check(self._lib.load().ltg_csv_to_csr(self._h, int(row_offset), int(n_rows), int(n_cols), self.n_threads, indptr.ctypes.data,
                                                        indices.ctypes.data, counts.ctypes.data, nnz.ctypes.data))
        k = int(nnz[0])
        return sparse.csr_matrix((counts[:k].astype(dtype), indices[:k].copy(), indptr), shape=(n_rows, n_cols), dtype=dtype)

    def close(self):
        if self._h:
            self._lib.load().ltg_csv_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def _native_csv():
    """LTG_NATIVE_CSV=0 keeps the pandas + NumPy readers below (the comparison arm of tests/test_data_processing.py)."""
    import os
    return os.environ.get("LTG_NATIVE_CSV", "1") != "0"


def load_train_data(csv_file, n_items):
    """data_processing.py:6-17: returns (CSR float32 [uid.max()+1, n_items], uid.min())."""
    if _native_csv():
        with CsvPairs(csv_file) as cp:
            if cp.n_pairs == 0:
                raise ValueError("%s holds no interactions" % csv_file)
            return cp.to_csr(cp.row_max + 1, n_items, np.float32), cp.row_min
    tp = pd.read_csv(csv_file)
    n_users = int(tp["uid"].max()) + 1
    indptr, indices, data = csr_from_pairs(tp["uid"].to_numpy(), tp["sid"].to_numpy(), n_users, n_items, np.float32)
    return sparse.csr_matrix((data, indices, indptr), shape=(n_users, n_items), dtype="float32"), int(tp["uid"].min())


def load_tr_te_data(csv_file_tr, csv_file_te, n_items):
    """data_processing.py:20-37: fold-in / held-out CSR float64 with the uid offset removed."""
    if _native_csv():
        with CsvPairs(csv_file_tr) as tr, CsvPairs(csv_file_te) as te:
            if tr.n_pairs == 0 or te.n_pairs == 0:
                raise ValueError("%s / %s: a split holds no interactions" % (csv_file_tr, csv_file_te))
            start_idx = min(tr.row_min, te.row_min)
            n = max(tr.row_max, te.row_max) - start_idx + 1
            return tr.to_csr(n, n_items, np.float64, start_idx), te.to_csr(n, n_items, np.float64, start_idx), start_idx
    tp_tr = pd.read_csv(csv_file_tr)
    tp_te = pd.read_csv(csv_file_te)
    start_idx = int(min(tp_tr["uid"].min(), tp_te["uid"].min()))
    end_idx = int(max(tp_tr["uid"].max(), tp_te["uid"].max()))
    n = end_idx - start_idx + 1
    out = []
    for tp in (tp_tr, tp_te):
        indptr, indices, data = csr_from_pairs(tp["uid"].to_numpy() - start_idx, tp["sid"].to_numpy(), n, n_items, np.float64)
        out.append(sparse.csr_matrix((data, indices, indptr), shape=(n, n_items), dtype="float64"))
    return out[0], out[1], start_idx


# ---------------------------------------------------------------------------------------------------------------------
# GAN side tables (data_processing.py:40-340)
# ---------------------------------------------------------------------------------------------------------------------
def _read_show2id(show2id_path):
    SHOW2ID = {}
    with codecs.open(show2id_path, "r", "utf-8") as f:
        for row in f:
            s = row.strip().split("\t")
            SHOW2ID[s[0]] = s[1]
    return SHOW2ID


def load_item_one_hot_features(item_list_path, SHOW2ID, n_items):
    """data_processing.py:40-70. The reference materialises an n_items-long one-hot list per item; only membership
    (`id in ITEM_FEATURE_DICT`, train.py:240) and FEATURE_LEN are ever used, so the dict maps id -> id and the dense
    array (fed to a placeholder nobody reads, train.py:300) is not built."""
    ITEM_OH_DICT = {}
    FEATURE_LEN = 0
    with codecs.open(item_list_path, "r", "utf-8") as f:
        for row in f:
            s = row.strip()
            if s in SHOW2ID:
                ITEM_OH_DICT[int(SHOW2ID[s])] = int(SHOW2ID[s])
                FEATURE_LEN = n_items
    return ITEM_OH_DICT, FEATURE_LEN, None


def load_user_items(csv_file_path):
    """data_processing.py:72-96: uid -> list of sids in file order (the first two columns of the file, whatever their names)."""
    if _native_csv():
        with open(csv_file_path, "r", encoding="utf-8-sig") as f:
            names = [x.strip().strip('"') for x in f.readline().rstrip("\r\n").split(",")]
        if len(names) >= 2 and names[0] != names[1]:
            with CsvPairs(csv_file_path, names[0], names[1]) as cp:
                u, s_ = cp.pairs()
            order = np.argsort(u, kind="stable")             # users in ascending id, each user's items in file order
            us, ss = u[order], s_[order]
            cut = np.nonzero(np.diff(us))[0] + 1
            starts = np.concatenate([[0], cut]).astype(np.int64)
            keys = us[starts].tolist() if len(us) else []
            vals = np.split(ss, cut) if len(us) else []
            by_user = {k: v.tolist() for k, v in zip(keys, vals)}
            # the reference's dict is filled in order of first appearance; keep that iteration order
            first_seen = u[np.sort(np.unique(u, return_index=True)[1])].tolist() if len(u) else []
            return {k: by_user[k] for k in first_seen}
    tp = pd.read_csv(csv_file_path)
    u = tp.iloc[:, 0].to_numpy()
    s = tp.iloc[:, 1].to_numpy()
    out = {}
    for uid, sid in zip(u.tolist(), s.tolist()):
        out.setdefault(uid, []).append(sid)
    return out


class OverlapCoeffs(object):
    """OVERLAP_COEFFS of the reference (a dict of dicts, data_processing.py:110-167) held SPARSE: the co-occurrence counts
    C = X^T X as CSR plus the item degrees; a coefficient C[a,b] / min(C[a,a], C[b,b]) is formed only for the (niche x niche) and
    (niche x popular) blocks the two consumers read. (Round 1 densified C into an I x I float64 matrix and ~4 temporaries of the same
    size: 13-16 GB at the ML-20M catalog, impossible at 1 M items.) `OV[a]` still yields the dense row a, `OV[a][b]` a coefficient."""

    DENSE_LIMIT = 30000   # `.matrix` (dense I x I float64, compatibility / tests) is refused above this catalog size

    def __init__(self, C, deg):
        self.C = C.tocsr()
        self.C.sort_indices()
        self.deg = np.asarray(deg, dtype=np.float64)
        self.present = self.deg > 0
        self.n_items = self.C.shape[0]

    def block(self, rows, cols):
        """Dense float64 block M[rows][:, cols] (the same division the reference performs; 0 where an item never occurs)."""
        rows = np.asarray(rows, dtype=np.int64); cols = np.asarray(cols, dtype=np.int64)
        c = np.asarray(self.C[rows][:, cols].todense(), dtype=np.float64)
        denom = np.minimum(self.deg[rows][:, None], self.deg[cols][None, :])
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(denom > 0, c / denom, 0.0)

    def colmax(self, rows, n_cols_mask=None):
        """max over `rows` of M[rows, j] for every item j (length n_items), touching only the non-zeros of those rows."""
        rows = np.asarray(rows, dtype=np.int64)
        sub = self.C[rows]
        best = np.zeros(self.n_items, dtype=np.float64)
        if sub.nnz:
            r_of = np.repeat(rows, np.diff(sub.indptr))
            denom = np.minimum(self.deg[r_of], self.deg[sub.indices])
            with np.errstate(divide="ignore", invalid="ignore"):
                val = np.where(denom > 0, sub.data.astype(np.float64) / denom, 0.0)
            np.maximum.at(best, sub.indices, val)
        return best

    @property
    def matrix(self):
        if self.n_items > self.DENSE_LIMIT:
            raise MemoryError("OverlapCoeffs.matrix would be a dense %d x %d float64 array; use block()/colmax() (catalogs above %d items "
                              "are served sparse)" % (self.n_items, self.n_items, self.DENSE_LIMIT))
        allr = np.arange(self.n_items)
        return self.block(allr, allr)

    def __getitem__(self, a):
        return self.block([int(a)], np.arange(self.n_items))[0]

    def __contains__(self, a):
        return bool(self.present[a])

    def __len__(self):
        return int(self.present.sum())


def load_overlap_coeff(show2id_path, user_tag_matrix_path):
    """data_processing.py:110-167 as one sparse product: C = X^T X on the binary user x item matrix of item_counts.csv, kept sparse;
    coefficient = C[a,b] / min(C[a,a], C[b,b]) in float64 (the same division the reference performs), formed on demand."""
    SHOW2ID = _read_show2id(show2id_path)
    n_items = int(max(int(v) for v in SHOW2ID.values())) + 1
    if _native_csv():
        # integer ids (every bundled / benchmark dataset): native parse of the first two columns + a lookup table for SHOW2ID
        try:
            keys = np.asarray([int(k) for k in SHOW2ID.keys()], dtype=np.int64)
            vals = np.asarray([int(v) for v in SHOW2ID.values()], dtype=np.int64)
            with open(user_tag_matrix_path, "r", encoding="utf-8-sig") as f:
                names = [x.strip().strip('"') for x in f.readline().rstrip("\r\n").split(",")]
            ok = len(keys) > 0 and keys.min() >= 0 and keys.max() < (1 << 26) and len(names) >= 2 and names[0] != names[1] and \
                all(str(int(k)) == k for k in list(SHOW2ID.keys())[:1000])
        except ValueError:
            ok = False
        if ok:
            from . import _lib
            try:
                with CsvPairs(user_tag_matrix_path, names[0], names[1]) as cp:
                    users, tags = cp.pairs()
            except _lib.LtgError:
                users = None      # ids that are not integers: the string path below
            if users is not None:
                lut = np.full(int(keys.max()) + 1, -1, dtype=np.int64)
                lut[keys] = vals
                inside = (tags >= 0) & (tags < len(lut))
                item = np.where(inside, lut[np.clip(tags, 0, len(lut) - 1)], -1)
                keep = item >= 0                      # tags without an entry in item2id.txt are skipped (data_processing.py:141-145)
                _, uidx = np.unique(users[keep], return_inverse=True)
                return overlap_from_interactions(uidx, item[keep], n_items)
    tp = pd.read_csv(user_tag_matrix_path, dtype=str)
    users = tp.iloc[:, 0].to_numpy()
    tags = tp.iloc[:, 1].to_numpy()
    keep = np.asarray([t in SHOW2ID for t in tags])
    users, tags = users[keep], tags[keep]
    item = np.asarray([int(SHOW2ID[t]) for t in tags], dtype=np.int64)
    _, uidx = np.unique(users, return_inverse=True)
    return overlap_from_interactions(uidx, item, n_items)


def overlap_from_interactions(uidx, item, n_items):
    """Sparse OverlapCoeffs from (user index, item id) interaction pairs (a pair counts once: the reference builds sets)."""
    X = sparse.csr_matrix((np.ones(len(item), dtype=np.int64), (uidx, item)), shape=(int(np.max(uidx)) + 1 if len(uidx) else 1, n_items))
    X.sum_duplicates()
    X.data[:] = 1
    C = (X.T @ X).tocsr()
    return OverlapCoeffs(C, C.diagonal())


def _native_tables():
    """LTG_NATIVE_TABLES=0 keeps the NumPy loops of load_items_to_sample / load_vectors (the comparison arm of the tests)."""
    import os
    return os.environ.get("LTG_NATIVE_TABLES", "1") != "0"


def _ragged_lists(d, N, eligible):
    """dict user -> list of ids  ->  (ptr int64 [N+1], items int32) over the eligible users, list order kept."""
    ptr = np.zeros(N + 1, dtype=np.int64)
    chunks = []
    for u in np.nonzero(eligible)[0].tolist():
        v = np.asarray(d[u], dtype=np.int64)
        chunks.append(v)
        ptr[u + 1] = len(v)
    items = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int64)
    return np.cumsum(ptr), np.ascontiguousarray(items, dtype=np.int32)


def _cooc_args(OV):
    """(indptr int64, indices, indices_are_64, counts int64, deg float64) of the sparse co-occurrence matrix for the C-ABI."""
    C = OV.C
    indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
    is64 = C.indices.dtype == np.int64
    indices = np.ascontiguousarray(C.indices, dtype=np.int64 if is64 else np.int32)
    counts = np.ascontiguousarray(C.data, dtype=np.int64)
    return indptr, indices, int(is64), counts, np.ascontiguousarray(OV.deg, dtype=np.float64)


def _eligible_users(user_popular_data, user_niche_data, N):
    el = np.zeros(N, dtype=np.uint8)
    for u in user_niche_data:
        if isinstance(u, (int, np.integer)) and 0 <= u < N and u in user_popular_data:
            el[u] = 1
    return el


def load_items_to_sample(user_popular_data, user_niche_data, NICHE_TAGS, OVERLAP_COEFFS, N, n_threads=0):
    """data_processing.py:170-224: candidates = the user's niche items + the top max(2n, 10-n) other niche items ranked by
    their best overlap with any of the user's niche items (stable: ties keep ascending item id, which is the iteration
    order of the reference's `NICHE_TAGS - curr_niche_tags` set of small ints). Runs on all host cores through ltg_cand_sets
    (csrc/tables.cu); the NumPy loop below is the same computation (LTG_NATIVE_TABLES=0)."""
    if _native_tables() and isinstance(OVERLAP_COEFFS, OverlapCoeffs):
        from . import _lib
        el = _eligible_users(user_popular_data, user_niche_data, N)
        un_ptr, un_items = _ragged_lists(user_niche_data, N, el)
        n_u = np.diff(un_ptr)
        out_ptr = np.concatenate([[0], np.cumsum(np.where(el > 0, n_u + np.maximum(2 * n_u, 10 - n_u), 0))]).astype(np.int64)
        out_items = np.empty(max(1, int(out_ptr[-1])), dtype=np.int32)
        out_count = np.zeros(max(1, N), dtype=np.int32)
        niche_sorted = np.ascontiguousarray(sorted(NICHE_TAGS), dtype=np.int32)
        indptr, indices, is64, counts, deg = _cooc_args(OVERLAP_COEFFS)
        _lib.check(_lib.load().ltg_cand_sets(indptr.ctypes.data, indices.ctypes.data, is64, counts.ctypes.data, deg.ctypes.data,
                                             OVERLAP_COEFFS.n_items, niche_sorted.ctypes.data, len(niche_sorted), un_ptr.ctypes.data,
                                             un_items.ctypes.data, el.ctypes.data, N, int(n_threads), out_ptr.ctypes.data,
                                             out_items.ctypes.data, out_count.ctypes.data))
        return {u: out_items[out_ptr[u]: out_ptr[u] + out_count[u]].astype(np.int64) for u in np.nonzero(el)[0].tolist()}
    niche_sorted = np.asarray(sorted(NICHE_TAGS), dtype=np.int64)
    out = {}
    for user_idx in range(N):
        if user_idx not in user_popular_data or user_idx not in user_niche_data:
            continue
        cur = np.asarray(user_niche_data[user_idx], dtype=np.int64)
        n = len(cur)
        num_sample = max(2 * n, 10 - n)
        others = niche_sorted[~np.isin(niche_sorted, cur)]
        best = OVERLAP_COEFFS.colmax(cur)[others] if len(others) else np.zeros(0)
        order = np.argsort(-best, kind="stable")[: min(num_sample, len(others))]
        out[user_idx] = np.sort(np.concatenate([cur, others[order]]))
    return out


def load_vectors(user_popular_data, user_niche_data, OVERLAP_COEFFS, ITEM_FEATURE_DICT, N, n_threads=0):
    """data_processing.py:227-271: for each niche item of the user the popular item of the user with the highest overlap
    (first maximum in list order), dropped when either id is not in ITEM_FEATURE_DICT. Runs on all host cores through
    ltg_real_pairs (csrc/tables.cu); the NumPy loop below is the same computation (LTG_NATIVE_TABLES=0)."""
    if _native_tables() and isinstance(OVERLAP_COEFFS, OverlapCoeffs):
        from . import _lib
        el = _eligible_users(user_popular_data, user_niche_data, N)
        un_ptr, un_items = _ragged_lists(user_niche_data, N, el)
        up_ptr, up_items = _ragged_lists(user_popular_data, N, el)
        n_items = OVERLAP_COEFFS.n_items
        valid = np.zeros(n_items, dtype=np.uint8)
        keys = np.asarray([k for k in ITEM_FEATURE_DICT.keys() if 0 <= k < n_items], dtype=np.int64)
        valid[keys] = 1
        out_n = np.empty(max(1, len(un_items)), dtype=np.int32); out_p = np.empty(max(1, len(un_items)), dtype=np.int32)
        out_count = np.zeros(max(1, N), dtype=np.int32)
        indptr, indices, is64, counts, deg = _cooc_args(OVERLAP_COEFFS)
        _lib.check(_lib.load().ltg_real_pairs(indptr.ctypes.data, indices.ctypes.data, is64, counts.ctypes.data, deg.ctypes.data, n_items,
                                              valid.ctypes.data, un_ptr.ctypes.data, un_items.ctypes.data, up_ptr.ctypes.data,
                                              up_items.ctypes.data, el.ctypes.data, N, int(n_threads), out_n.ctypes.data, out_p.ctypes.data,
                                              out_count.ctypes.data))
        x_niche, x_pop = {}, {}
        for u in np.nonzero(el)[0].tolist():
            a, b = int(un_ptr[u]), int(un_ptr[u]) + int(out_count[u])
            x_niche[u] = out_n[a:b].tolist()
            x_pop[u] = out_p[a:b].tolist()
        return x_niche, x_pop
    x_niche, x_pop = {}, {}
    for user_idx in range(N):
        if user_idx not in user_popular_data or user_idx not in user_niche_data:
            continue
        pops = np.asarray(user_popular_data[user_idx], dtype=np.int64)
        niches = np.asarray(user_niche_data[user_idx], dtype=np.int64)
        best = pops[np.argmax(OVERLAP_COEFFS.block(niches, pops), axis=1)]
        cn, cp = [], []
        for a, b in zip(niches.tolist(), best.tolist()):
            if a in ITEM_FEATURE_DICT and b in ITEM_FEATURE_DICT:
                cn.append(a)
                cp.append(b)
        x_niche[user_idx] = cn
        x_pop[user_idx] = cp
    return x_niche, x_pop


def load_pop_niche_tags(show2id_path, item_list_path, niche_tags_path, n_items):
    """data_processing.py:275-340."""
    SHOW2ID = _read_show2id(show2id_path)
    IDs_present = set()
    with codecs.open(item_list_path, "r", "utf-8") as f:
        for row in f:
            s = row.strip()
            if s in SHOW2ID:
                IDs_present.add(SHOW2ID[s])
    NICHE_TAGS = set()
    with codecs.open(niche_tags_path, "r", "utf-8") as f:
        for row in f:
            s = row.strip()
            if s in SHOW2ID and SHOW2ID[s] in IDs_present:
                NICHE_TAGS.add(int(SHOW2ID[s]))
    ALL_TAGS = list(range(n_items))
    OTHER_TAGS = np.asarray(sorted(set(ALL_TAGS) - NICHE_TAGS))
    return SHOW2ID, IDs_present, NICHE_TAGS, ALL_TAGS, OTHER_TAGS


# ---------------------------------------------------------------------------------------------------------------------
# flat tables for the device engine
# ---------------------------------------------------------------------------------------------------------------------
def _ragged(d, n):
    ptr = np.zeros(n + 1, dtype=np.int64)
    chunks = []
    for u in range(n):
        v = d.get(u)
        if v is not None and len(v):
            chunks.append(np.asarray(v, dtype=np.int64))
            ptr[u + 1] = len(chunks[-1])
    items = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int64)
    return np.cumsum(ptr).astype(np.int32), items.astype(np.int32)


def build_train_tables(train_data, user_popular_data, user_niche_data, user_x_niche_vectors, user_x_popular_n_vectors,
                       USER_TAGS_TO_SAMPLE, ITEM_FEATURE_DICT, n_items):
    """Reference-shaped side tables -> the keyword arguments of engine.TrainData."""
    N = train_data.shape[0]
    csr = train_data.tocsr()
    csr.sort_indices()
    if csr.nnz and float(csr.data.max()) != 1.0:
        # csr_matrix((ones,(rows,cols))) sums duplicate (uid, sid) rows (data_processing.py:13-15): the device tables are binary
        raise ValueError("train_GAN.csv holds repeated (uid, sid) rows (max count %g); the device tables carry binary interactions -- "
                         "deduplicate the file or pass the counts through TrainData/ltg_enc_gather_fwd `values`" % float(csr.data.max()))
    pop_ptr, pop_items = _ragged(user_popular_data, N)
    cand_ptr, cand_items = _ragged(USER_TAGS_TO_SAMPLE, N)
    real_ptr, real_niche = _ragged(user_x_niche_vectors, N)
    _, real_pop = _ragged(user_x_popular_n_vectors, N)
    n_niche = np.asarray([len(user_niche_data.get(u, ())) for u in range(N)], dtype=np.int32)
    eligible = np.asarray([(u in user_popular_data) and (u in user_niche_data) for u in range(N)], dtype=bool)
    item_valid = np.zeros(n_items, dtype=np.uint8)
    item_valid[np.asarray(sorted(ITEM_FEATURE_DICT.keys()), dtype=np.int64)] = 1
    return dict(n_items=n_items, indptr=csr.indptr.astype(np.int32), indices=csr.indices.astype(np.int32), pop_ptr=pop_ptr,
                pop_items=pop_items, n_niche=n_niche, cand_ptr=cand_ptr, cand_items=cand_items, real_ptr=real_ptr,
                real_niche=real_niche, real_pop=real_pop, eligible=eligible, item_valid=item_valid)


def tables_from_golden(npz):
    """The same tables from the committed fixture tests/golden/askubuntu_sample.npz (generated by the reference loaders)."""
    g = npz
    n_items = int(g["n_items"])
    item_valid = np.zeros(n_items, dtype=np.uint8)
    item_valid[g["valid_items"]] = 1
    n_niche = np.diff(g["niche_ptr"]).astype(np.int32)
    eligible = np.asarray(g["has_pop"], dtype=bool) & np.asarray(g["has_niche"], dtype=bool)
    return dict(n_items=n_items, indptr=g["train_indptr"].astype(np.int32), indices=g["train_indices"].astype(np.int32),
                pop_ptr=g["pop_ptr"], pop_items=g["pop_items"], n_niche=n_niche, cand_ptr=g["cand_ptr"], cand_items=g["cand_items"],
                real_ptr=g["real_ptr"], real_niche=g["real_niche"], real_pop=g["real_pop"], eligible=eligible, item_valid=item_valid)
