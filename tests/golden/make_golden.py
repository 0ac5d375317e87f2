"""Generates tests/golden/askubuntu_sample.npz by IMPORTING THE REFERENCE'S OWN MODULES from /root/reference/Codes
(data_processing.py, sample.py, eval_functions.py -- plain Python/NumPy, they run unchanged under Python 3.12) on the
bundled dataset /root/reference/Dataset/Askubuntu_Sample. Run in the build container only; the GPU box has no
/root/reference, it uses the committed .npz.

    python tests/golden/make_golden.py

`eval_functions.py` imports tensorflow and bottleneck at module scope; both are stubbed (an empty module, and
bottleneck.argpartition -> numpy.argpartition, which has the same contract) exactly as SURVEY.md 8c describes.
"""
import os
import sys
import types

import numpy as np
import pandas  # noqa: F401  (import before the stubs)

REF = "/root/reference"
CODES = os.path.join(REF, "Codes")
DATA = os.path.join(REF, "Dataset", "Askubuntu_Sample")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "askubuntu_sample.npz")


def main():
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    bn = types.ModuleType("bottleneck")
    bn.argpartition = np.argpartition
    bn.__version__ = "stub"
    sys.modules.setdefault("bottleneck", bn)
    sys.path.insert(0, CODES)
    import data_processing as dp
    import eval_functions as ef
    import sample as smp

    n_items = sum(1 for _ in open(os.path.join(DATA, "unique_item_id.txt")))
    show2id_path = os.path.join(DATA, "item2id.txt")
    item_list_path = os.path.join(DATA, "item_list.txt")
    SHOW2ID, IDs_present, NICHE_TAGS, ALL_TAGS, OTHER_TAGS = dp.load_pop_niche_tags(show2id_path, item_list_path,
                                                                                    os.path.join(DATA, "niche_items.txt"), n_items)
    ITEM_FEATURE_DICT, FEATURE_LEN, _ = dp.load_item_one_hot_features(item_list_path, SHOW2ID, n_items)
    train, uid_start = dp.load_train_data(os.path.join(DATA, "train_GAN.csv"), n_items)
    vad_tr, vad_te, vad_start = dp.load_tr_te_data(os.path.join(DATA, "validation_tr.csv"), os.path.join(DATA, "validation_te.csv"), n_items)
    tst_tr, tst_te, tst_start = dp.load_tr_te_data(os.path.join(DATA, "test_tr.csv"), os.path.join(DATA, "test_te.csv"), n_items)
    user_pop = dp.load_user_items(os.path.join(DATA, "train_GAN_popular.csv"))
    user_niche = dp.load_user_items(os.path.join(DATA, "train_GAN_niche.csv"))
    OVERLAP = dp.load_overlap_coeff(show2id_path, os.path.join(DATA, "item_counts.csv"))
    N = train.shape[0]
    real_niche, real_pop = dp.load_vectors(user_pop, user_niche, OVERLAP, ITEM_FEATURE_DICT, N)
    cand = dp.load_items_to_sample(user_pop, user_niche, NICHE_TAGS, OVERLAP, N)

    def ragged(d, n):
        ptr = np.zeros(n + 1, dtype=np.int64)
        items = []
        for u in range(n):
            v = list(d.get(u, []))
            items += [int(x) for x in v]
            ptr[u + 1] = len(items)
        return ptr.astype(np.int32), np.asarray(items, dtype=np.int32)

    out = dict(n_items=np.int32(n_items), feature_len=np.int32(FEATURE_LEN), uid_start=np.int64(uid_start), vad_start=np.int64(vad_start),
               tst_start=np.int64(tst_start), niche_tags=np.asarray(sorted(NICHE_TAGS), dtype=np.int32),
               valid_items=np.asarray(sorted(ITEM_FEATURE_DICT.keys()), dtype=np.int32))
    for name, m in (("train", train), ("vad_tr", vad_tr), ("vad_te", vad_te), ("tst_tr", tst_tr), ("tst_te", tst_te)):
        m = m.tocsr()
        m.sort_indices()
        out[name + "_indptr"] = m.indptr.astype(np.int32)
        out[name + "_indices"] = m.indices.astype(np.int16)
        out[name + "_data_is_one"] = np.bool_(bool((m.data == 1).all()))
        out[name + "_shape"] = np.asarray(m.shape, dtype=np.int64)
    out["pop_ptr"], out["pop_items"] = ragged(user_pop, N)
    out["niche_ptr"], out["niche_items"] = ragged(user_niche, N)
    out["has_pop"] = np.asarray([u in user_pop for u in range(N)])
    out["has_niche"] = np.asarray([u in user_niche for u in range(N)])
    out["cand_ptr"], out["cand_items"] = ragged({u: v.tolist() for u, v in cand.items()}, N)
    out["real_ptr"], out["real_niche"] = ragged(real_niche, N)
    _, out["real_pop"] = ragged(real_pop, N)
    # a slice of the overlap table (full table is 8 MB): rows of the first 40 items
    out["overlap_rows_0_40"] = np.asarray([[OVERLAP[a][b] for b in range(n_items)] for a in range(40)], dtype=np.float64)

    # ---- known answers of the verbatim metric code ----
    pop_scores = np.asarray(train.sum(axis=0)).ravel().astype(np.float32)

    def metrics(tr, te, pred):
        pred = pred.copy()
        pred[tr.nonzero()] = -np.inf
        nd = ef.NDCG_binary_at_k_batch(pred, te, k=100)
        r20, _ = ef.Recall_at_k_batch(pred, te, k=20)
        r50, _ = ef.Recall_at_k_batch(pred, te, k=50)
        return np.asarray([np.mean(nd), np.mean(r20), np.mean(r50), len(nd)], dtype=np.float64)

    out["ka_pop_vad"] = metrics(vad_tr, vad_te, np.tile(pop_scores, (vad_tr.shape[0], 1)))
    out["ka_pop_tst"] = metrics(tst_tr, tst_te, np.tile(pop_scores, (tst_tr.shape[0], 1)))
    rnd = np.random.RandomState(0).rand(vad_tr.shape[0], n_items).astype(np.float32)
    out["ka_rand_vad"] = metrics(vad_tr, vad_te, rnd)
    # per-user values on a small tie-free block, for exact comparison
    blk = rnd[:64]
    blk_m = blk.copy()
    blk_m[vad_tr[:64].nonzero()] = -np.inf
    out["ka_block_ndcg"] = np.asarray(ef.NDCG_binary_at_k_batch(blk_m, vad_te[:64], k=100), dtype=np.float64)
    out["ka_block_r20"] = np.asarray(ef.Recall_at_k_batch(blk_m, vad_te[:64], k=20)[0], dtype=np.float64)

    # ---- sampler: draws of the verbatim sample.py under a fixed NumPy seed ----
    u = int(np.nonzero(np.diff(out["cand_ptr"]) > 12)[0][0])
    c = out["cand_items"][out["cand_ptr"][u]:out["cand_ptr"][u + 1]]
    p = np.random.RandomState(5).rand(len(c)).astype(np.float32)
    np.random.seed(1234)
    draws = []
    for _ in range(200):
        m, ids = smp.sample_from_generator_new(c, p, 4, n_items)
        draws.append(np.sort(ids))
    out["samp_cand"] = c.astype(np.int32)
    out["samp_p"] = p
    out["samp_draws"] = np.asarray(draws, dtype=np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "N", N, "nnz", train.nnz, "real pairs", len(out["real_niche"]), "cand", len(out["cand_items"]))
    print("known answers:", out["ka_pop_vad"], out["ka_pop_tst"], out["ka_rand_vad"])


if __name__ == "__main__":
    main()
