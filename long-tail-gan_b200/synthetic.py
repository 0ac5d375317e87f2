"""Seeded synthetic inputs with the shapes of BASELINE.json's configurations.

There is no network for ML-20M / Netflix / MSD, and the reference's own side-table builders are O(I^2) Python
(data_processing.py:110-224), so the large shapes use: Zipf(1.0) item popularity (permuted, seed 1234); log-normal user
degrees with the dataset's mean, clipped to [5, min(I/4, 2000)] (seed 1235); items drawn per user proportionally to
popularity; niche set = least popular items that together hold 50% of the interactions; candidate set
C_u = own niche items + max(2 n_u, 10 - n_u) other niche items (the size rule of data_processing.py:182, with uniformly
random niche items standing in for the overlap-coefficient ranking); real pairs = (each niche item of the user, a random
popular item of the user) (stand-in for the arg-max overlap partner of data_processing.py:242-263); all items valid.
Everything is vectorised NumPy so the 136,677 x 20,108 configuration builds in seconds.
"""
import numpy as np

CONFIGS = {
    # name: (n_users, n_items, mean interactions per user)
    "ml20m": (136677, 20108, 73.0),
    "netflix": (463435, 17769, 123.0),
    "msd": (571355, 41140, 59.0),
    "askubuntu_shape": (10001, 1000, 17.9),
}


def _group_rank(keys_sorted_group):
    """rank of each element inside its group, for an array of group ids that is sorted by group."""
    n = len(keys_sorted_group)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    start = np.r_[True, keys_sorted_group[1:] != keys_sorted_group[:-1]]
    idx = np.arange(n)
    first = np.maximum.accumulate(np.where(start, idx, 0))
    return idx - first


def _draw_unique(rng, owner_counts, want, sampler, n_items, oversample=1.7):
    """For every owner u draw `want[u]` distinct items using `sampler(n) -> item ids`; returns (owner, item) sorted by
    (owner, item). Owners may end up with fewer items than wanted when the distribution is very peaked."""
    n_owner = len(want)
    draws = np.ceil(want * oversample).astype(np.int64) + 8
    owner = np.repeat(np.arange(n_owner, dtype=np.int64), draws)
    items = sampler(int(draws.sum())).astype(np.int64)
    key = owner * n_items + items
    # first occurrence of each (owner, item) in draw order
    uniq, first = np.unique(key, return_index=True)
    first.sort()
    owner, items = owner[first], items[first]          # still grouped by owner, in draw order
    rank = _group_rank(owner)
    keep = rank < want[owner]
    owner, items = owner[keep], items[keep]
    order = np.lexsort((items, owner))
    return owner[order], items[order]


def make_interactions(n_users, n_items, mean_deg, seed=1234):
    rng_pop = np.random.RandomState(seed)
    rng_deg = np.random.RandomState(seed + 1)
    pop = 1.0 / np.arange(1, n_items + 1, dtype=np.float64)
    pop = pop[rng_pop.permutation(n_items)]
    cdf = np.cumsum(pop / pop.sum())
    sigma = 1.0
    deg = rng_deg.lognormal(np.log(mean_deg) - 0.5 * sigma * sigma, sigma, size=n_users)
    deg = np.clip(np.rint(deg), 5, min(n_items // 4, 2000)).astype(np.int64)
    rng = np.random.RandomState(seed + 2)
    owner, items = _draw_unique(rng, None, deg, lambda n: np.minimum(np.searchsorted(cdf, rng.rand(n)), n_items - 1), n_items)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(indptr, owner + 1, 1)
    return np.cumsum(indptr), items.astype(np.int32)


def make_side_tables(indptr, indices, n_items, seed=1234, niche_share=0.5):
    """GAN side tables in the CSR form engine.TrainData expects."""
    n_users = len(indptr) - 1
    rng = np.random.RandomState(seed + 3)
    counts = np.bincount(indices, minlength=n_items).astype(np.int64)
    order = np.argsort(counts, kind="stable")                    # least popular first
    csum = np.cumsum(counts[order])
    n_niche_items = int(np.searchsorted(csum, niche_share * csum[-1], side="right"))
    is_niche = np.zeros(n_items, dtype=bool)
    is_niche[order[:n_niche_items]] = True
    niche_ids = np.nonzero(is_niche)[0]
    owner = np.repeat(np.arange(n_users, dtype=np.int64), np.diff(indptr))
    nz_niche = is_niche[indices]
    n_niche = np.bincount(owner[nz_niche], minlength=n_users).astype(np.int64)
    n_pop = np.bincount(owner[~nz_niche], minlength=n_users).astype(np.int64)
    eligible = (n_niche > 0) & (n_pop > 0)
    pop_ptr = np.concatenate([[0], np.cumsum(n_pop)])
    pop_items = indices[~nz_niche]
    # candidates: own niche items first (always kept), then random other niche items
    extra = np.where(eligible, np.maximum(2 * n_niche, 10 - n_niche), 0)
    extra = np.minimum(extra, len(niche_ids) - n_niche)
    own_owner, own_items = owner[nz_niche & eligible[owner]], indices[nz_niche & eligible[owner]].astype(np.int64)
    draws = np.ceil(extra * 1.5).astype(np.int64) + np.where(extra > 0, 8, 0)
    d_owner = np.repeat(np.arange(n_users, dtype=np.int64), draws)
    d_items = niche_ids[rng.randint(0, len(niche_ids), size=int(draws.sum()))].astype(np.int64)
    all_owner = np.concatenate([own_owner, d_owner])
    all_items = np.concatenate([own_items, d_items])
    is_own = np.concatenate([np.ones(len(own_owner), dtype=bool), np.zeros(len(d_owner), dtype=bool)])
    key = all_owner * n_items + all_items
    _, first = np.unique(key, return_index=True)                 # first occurrence: own items precede the random draws
    first.sort()
    all_owner, all_items, is_own = all_owner[first], all_items[first], is_own[first]
    order2 = np.lexsort((~is_own, all_owner))                    # per owner: own first, then draws in draw order (stable)
    all_owner, all_items, is_own = all_owner[order2], all_items[order2], is_own[order2]
    rank = _group_rank(all_owner)
    keep = rank < (n_niche + extra)[all_owner]
    all_owner, all_items = all_owner[keep], all_items[keep]
    order3 = np.lexsort((all_items, all_owner))
    cand_owner, cand_items = all_owner[order3], all_items[order3]
    cand_ptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(cand_ptr, cand_owner + 1, 1)
    cand_ptr = np.cumsum(cand_ptr)
    # real pairs
    r_owner = own_owner
    r_niche = own_items
    pick = pop_ptr[r_owner] + np.floor(rng.rand(len(r_owner)) * n_pop[r_owner]).astype(np.int64)
    r_pop = pop_items[pick]
    real_ptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(real_ptr, r_owner + 1, 1)
    real_ptr = np.cumsum(real_ptr)
    return dict(n_items=n_items, indptr=indptr.astype(np.int32), indices=indices.astype(np.int32), pop_ptr=pop_ptr.astype(np.int32),
                pop_items=pop_items.astype(np.int32), n_niche=n_niche.astype(np.int32), cand_ptr=cand_ptr.astype(np.int32),
                cand_items=cand_items.astype(np.int32), real_ptr=real_ptr.astype(np.int32), real_niche=r_niche.astype(np.int32),
                real_pop=r_pop.astype(np.int32), eligible=eligible, item_valid=np.ones(n_items, dtype=np.uint8))


def make_config(name, n_users=None, seed=1234):
    """Training-side tables for one of CONFIGS (optionally truncated to the first n_users users)."""
    N, I, deg = CONFIGS[name]
    if n_users is not None:
        N = min(N, int(n_users))
    indptr, indices = make_interactions(N, I, deg, seed)
    return make_side_tables(indptr, indices, I, seed)


def make_eval_split(n_users, n_items, mean_deg, seed=4321, heldout=0.2):
    """Held-out users with an 80/20 fold-in / held-out split. Returns (tr_indptr, tr_indices, te_indptr, te_indices)."""
    indptr, indices = make_interactions(n_users, n_items, mean_deg, seed)
    rng = np.random.RandomState(seed + 9)
    is_te = rng.rand(len(indices)) < heldout
    owner = np.repeat(np.arange(n_users, dtype=np.int64), np.diff(indptr))

    def part(sel):
        p = np.zeros(n_users + 1, dtype=np.int64)
        np.add.at(p, owner[sel] + 1, 1)
        return np.cumsum(p).astype(np.int32), indices[sel].astype(np.int32)

    tr = part(~is_te)
    te = part(is_te)
    return tr[0], tr[1], te[0], te[1]
