"""CPU tests of the host logic that prepares the device inputs of a step: engine.TrainData (per-batch index structures), the
data-parallel shard tables (engine.build_dp_shard_tables), the catalog-shard tables (vocab_parallel.shard_tables / ShardData) and the
encoder work list -- built on device="cpu" from the reference-generated fixture (tests/golden/askubuntu_sample.npz) and checked
against plain NumPy restatements of what the kernels expect (train.py:192-251 for the pair layout)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "askubuntu_sample.npz")
B = 100     # config.ini BATCH_SIZE


@pytest.fixture(scope="module")
def mods():
    return dict(eng=importlib.import_module("long-tail-gan_b200.engine"), dp=importlib.import_module("long-tail-gan_b200.data_processing"),
                ops=importlib.import_module("long-tail-gan_b200.ops"), vp=importlib.import_module("long-tail-gan_b200.vocab_parallel"))


@pytest.fixture(scope="module")
def tabs(mods):
    return mods["dp"].tables_from_golden(np.load(GOLD))


@pytest.fixture(scope="module")
def data(mods, tabs):
    return mods["eng"].TrainData(batch_size=B, device="cpu", **tabs)


def test_batches_cover_users_and_pairs(data, tabs):
    N = len(tabs["indptr"]) - 1
    assert data.N == N == 10001 and len(data.batches) == 101                      # ceil(10001 / 100), train.py:192
    assert [bt["b0"] for bt in data.batches] == list(range(0, N, B)) and data.batches[-1]["B"] == 1
    assert sum(bt["B"] for bt in data.batches) == N
    assert sum(bt["Pr"] for bt in data.batches) == 92814                           # every precomputed real pair, once
    assert sum(bt["nnz"] for bt in data.batches) == 179368
    elig = np.asarray(tabs["eligible"], dtype=bool)
    cand_len = np.diff(np.asarray(tabs["cand_ptr"], dtype=np.int64))
    n_draw = np.minimum(np.where(elig, tabs["n_niche"], 0), cand_len)
    assert sum(bt["K"] for bt in data.batches) == int(n_draw.sum())                # one slot per draw of train.py:227
    assert data.max_B == B and data.max_P == max(bt["P"] for bt in data.batches)
    for bt in data.batches[:5] + data.batches[-2:]:
        b0, nB, Pr, K = bt["b0"], bt["B"], bt["Pr"], bt["K"]
        r0, r1 = int(tabs["real_ptr"][b0]), int(tabs["real_ptr"][b0 + nB])
        # real pairs first (label 0), then K slots for the generated pairs (label -1 until the sampler fills them), user-major
        assert np.array_equal(bt["pair_niche"][:Pr].numpy(), tabs["real_niche"][r0:r1])
        assert np.array_equal(bt["pair_pop"][:Pr].numpy(), tabs["real_pop"][r0:r1])
        assert (bt["label"][:Pr] == 0).all() and (bt["label"][Pr:Pr + K] == -1).all() and bt["P"] == Pr + K
        assert np.array_equal(np.diff(bt["samp_ptr"].numpy()), n_draw[b0:b0 + nB])
        assert bt["max_cand"] == int(cand_len[b0:b0 + nB].max())
        so = bt["samp_order"].numpy()
        assert sorted(so.tolist()) == list(range(nB)) and (np.diff(cand_len[b0:b0 + nB][so]) <= 0).all()   # longest candidate list first


def test_active_item_structures_match_the_batch_csr(data, tabs):
    indptr = np.asarray(tabs["indptr"], dtype=np.int64); indices = np.asarray(tabs["indices"], dtype=np.int64)
    for bt in (data.batches[0], data.batches[37], data.batches[-1]):
        b0, nB = bt["b0"], bt["B"]
        e0, e1 = int(indptr[b0]), int(indptr[b0 + nB])
        items = indices[e0:e1]
        active = np.unique(items)
        slot = bt["slot_of_item"].numpy()
        assert bt["n_active"] == len(active)
        assert np.array_equal(np.nonzero(slot >= 0)[0], active) and np.array_equal(slot[active], np.arange(len(active)))
        # compact CSC over the active items: for slot s, entries act_ptr[s]..act_ptr[s+1] list the batch rows (ascending) that hold
        # the item and the position of that interaction in the global CSR (what the coefficient array is indexed by)
        act_ptr, csc_row, csc_pos = bt["act_ptr"].numpy(), bt["csc_row"].numpy(), bt["csc_pos"].numpy()
        assert act_ptr[-1] == e1 - e0
        rows = np.repeat(np.arange(nB), np.diff(indptr[b0:b0 + nB + 1]))
        for s in (0, len(active) // 2, len(active) - 1):
            sl = slice(act_ptr[s], act_ptr[s + 1])
            assert (indices[csc_pos[sl]] == active[s]).all()
            assert np.array_equal(csc_row[sl], np.sort(rows[items == active[s]]))
            assert np.array_equal(rows[csc_pos[sl] - e0], csc_row[sl])


def test_encoder_work_list_covers_every_chunk_once(mods):
    ops = mods["ops"]
    rng = np.random.RandomState(0)
    deg = np.concatenate([[0, 1, 127, 128, 129, 941], rng.randint(0, 400, size=60)])
    indptr = np.concatenate([[0], np.cumsum(deg)])
    w = ops.enc_work_list(indptr).astype(np.int64)
    rows, chunks = w & ((1 << 20) - 1), w >> 20
    want = sorted((r, c) for r, d in enumerate(deg) for c in range(max(1, -(-int(d) // ops.ENC_CHUNK))))
    assert sorted(zip(rows.tolist(), chunks.tolist())) == want                     # empty rows keep their chunk 0 (bias + tanh)
    size = np.minimum(ops.ENC_CHUNK, deg[rows] - chunks * ops.ENC_CHUNK)
    assert (np.diff(size) <= 0).all()                                              # full chunks first (longest work first)


@pytest.mark.parametrize("world", [2, 4])
def test_dp_shard_tables_partition_the_global_batch(mods, tabs, world):
    """engine.build_dp_shard_tables: for local batch bi, rank r lists exactly the interactions of the GLOBAL batch (the ranks' bi-th
    batches side by side) whose item falls in its row shard; over the ranks every interaction appears once."""
    eng = mods["eng"]
    Bl = B // world if world == 2 else 25
    n_local_total = int(np.ceil((len(tabs["indptr"]) - 1) / Bl))
    nb = n_local_total // world
    I = int(tabs["n_items"])
    R = (I + world - 1) // world
    indptr = np.asarray(tabs["indptr"], dtype=np.int64); indices = np.asarray(tabs["indices"], dtype=np.int64)
    per_rank = []
    for r in range(world):
        d = eng.TrainData(batch_size=Bl, device="cpu", first_batch=r * nb, max_batches=3, **tabs)
        per_rank.append(eng.build_dp_shard_tables(d, tabs["indptr"], tabs["indices"], world, r, nb, R))
    for bi in range(3):
        seen = []
        for r in range(world):
            tb = per_rank[r][bi]
            n = tb["n_entries"]
            rows, items, slots = tb["e_row"].numpy()[:n], tb["e_item"].numpy()[:n], tb["e_slot"].numpy()[:n]
            assert ((items >= r * R) & (items < (r + 1) * R)).all()
            sl = tb["slot_local"].numpy()
            assert np.array_equal(sl[items - r * R], slots) and tb["n_active"] == len(np.unique(items))
            # row q*Bl + j of the global batch is user (q*nb + bi)*Bl + j; its uid and 1/sqrt(nnz) ride along
            q, j = rows // Bl, rows % Bl
            users = (q * nb + bi) * Bl + j
            assert np.array_equal(tb["row_uid"].numpy()[rows], users)
            deg = np.diff(indptr)[users]
            assert np.allclose(tb["row_rnorm"].numpy()[rows], 1.0 / np.sqrt(deg), rtol=1e-6)
            seen += list(zip(users.tolist(), items.tolist()))
        want = []
        for q in range(world):
            u0 = (q * nb + bi) * Bl
            for u in range(u0, u0 + Bl):
                want += [(u, int(i)) for i in indices[indptr[u]: indptr[u + 1]]]
        assert sorted(seen) == sorted(want)


@pytest.mark.parametrize("world", [2, 3])
def test_catalog_shard_tables_partition_items_and_candidates(mods, tabs, world):
    vp = mods["vp"]
    I = int(tabs["n_items"])
    bounds = vp.shard_bounds(I, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == I and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    assert all((hi - lo) % 8 == 0 for lo, hi in bounds[:-1])                       # TMA pitch / vector loads
    indptr = np.asarray(tabs["indptr"], dtype=np.int64); indices = np.asarray(tabs["indices"], dtype=np.int64)
    nnz = 0
    own_total = np.zeros(len(tabs["cand_items"]), dtype=np.int64)
    for lo, hi in bounds:
        st = vp.shard_tables(tabs, lo, hi)
        nnz += len(st["indices"])
        u = 4321
        got = st["indices"][st["indptr"][u]: st["indptr"][u + 1]] + lo
        row = indices[indptr[u]: indptr[u + 1]]
        assert np.array_equal(got, row[(row >= lo) & (row < hi)])
        assert np.allclose(st["row_rnorm"], 1.0 / np.sqrt(np.maximum(np.diff(indptr), 1e-12)))   # the norm of the WHOLE row
        sd = vp.ShardData(st, lo, hi, B, device="cpu", max_batches=2)
        for bt in sd.batches:
            pos, lid, rw = bt["cand_own_pos"].numpy(), bt["cand_own_lid"].numpy(), bt["cand_own_row"].numpy()
            assert np.array_equal(np.asarray(tabs["cand_items"])[pos], lid + lo)
            c0 = int(tabs["cand_ptr"][bt["b0"]])
            assert (pos >= c0).all() and (pos < int(tabs["cand_ptr"][bt["b0"] + bt["B"]])).all()
            assert np.array_equal(np.searchsorted(np.asarray(tabs["cand_ptr"]), pos, side="right") - 1 - bt["b0"], rw)
            own_total[pos] += 1
    assert nnz == len(indices)
    c_end = int(tabs["cand_ptr"][2 * B])
    assert (own_total[:c_end] == 1).all()                                          # every candidate of the first two batches has one owner
