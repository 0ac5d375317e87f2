"""World-size-2 gloo test (CPU) of the data-parallel formulation the device path implements (engine.run_g_step with
world_size > 1): each rank takes half of the global batch, normalises its mean losses by B_global, all-reduces the three
adversarial scalars (sum p, sum y, cnt) BEFORE the backward pass (the GAN term multiplies global sums), then
SUM-all-reduces the gradients. The result must equal the single-process gradient of the whole batch (oracle)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _problem():
    from oracle import ltgan_oracle as orc
    import helpers
    rng = np.random.RandomState(3)
    I, B = 200, 16
    tabs = helpers.synth_side_tables(rng, B, I, mean_nnz=10)
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], 0, B, I))
    params = orc.init_vae_params(I, seed=5)
    params[3] = params[3] * 3
    E, dparams = orc.init_disc_params(I, 20, 30, 40, 50, seed=6)
    keep = torch.from_numpy(rng.rand(B, I) < 0.75)
    eps = torch.from_numpy(rng.randn(B, 200).astype(np.float32))
    # sampled masks / generated pairs per user
    mask = torch.zeros(B, I)
    pg, xg, owner = [], [], []
    for u in range(B):
        if not tabs["eligible"][u]:
            continue
        c = tabs["cand_items"][tabs["cand_ptr"][u]:tabs["cand_ptr"][u + 1]]
        pick = rng.choice(c, size=min(3, len(c)), replace=False)
        pops = tabs["pop_items"][tabs["pop_ptr"][u]:tabs["pop_ptr"][u + 1]]
        for it in np.sort(pick):
            mask[u, it] = 1.0
            xg.append(int(it)); pg.append(int(pops[rng.randint(len(pops))])); owner.append(u)
    return dict(orc=orc, X=X, params=params, E=E, dparams=dparams, keep=keep, eps=eps, mask=mask, xg=np.asarray(xg), pg=np.asarray(pg),
                owner=np.asarray(owner), B=B, I=I)


def _local_grads(P, rows, B_global, sums=None):
    """Gradient of this rank's share of g_loss. `sums` = (sum_p, sum_y, cnt) global values (None: compute local and return)."""
    orc = P["orc"]
    ps = [p.clone().requires_grad_(True) for p in P["params"]]
    out = orc.vae_forward(ps, P["X"][rows], P["keep"][rows], 0.75, P["eps"][rows], 1.0, 0.1)
    sel = np.isin(P["owner"], rows)
    y = orc.disc_forward(P["E"], P["dparams"], torch.from_numpy(P["pg"][sel]), torch.from_numpy(P["xg"][sel]), None, 1.0).detach()
    s_local = (out["probs"] * P["mask"][rows]).sum()
    if sums is None:
        return float(s_local.detach()), float(y.sum()), float(sel.sum())
    sum_p, sum_y, cnt = sums
    ybar = sum_y / cnt
    # local share of: mean losses (normalised by B_global) + gan = -(lam/cnt) * SUM_P * SUM_Y, whose gradient w.r.t. the
    # local probabilities is -(lam * ybar) * d(local sum p)
    n_loc = len(rows)
    loss = (out["neg_ll"] + 0.1 * out["KL"]) * (n_loc / B_global) - 1.0 * ybar * s_local
    return [g.detach() for g in torch.autograd.grad(loss, ps)]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dmod = importlib.import_module("long-tail-gan_b200.dist")
    P = _problem()
    first, count = dmod.shard_range(P["B"], rank, world)   # shard users the way batches are sharded
    rows = np.arange(first, first + count)
    lay = dmod.global_step_layout(count, world)
    assert lay["B_global"] == P["B"]
    sums = torch.tensor(_local_grads(P, rows, lay["B_global"]), dtype=torch.float64)
    dist.all_reduce(sums)                                    # F3: global sum p, sum y, cnt before the backward
    grads = _local_grads(P, rows, lay["B_global"], tuple(float(x) for x in sums))
    dmod.allreduce_sum_(grads)
    if rank == 0:
        ret["grads"] = [g.numpy() for g in grads]
        ret["sums"] = sums.numpy()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process():
    P = _problem()
    orc = P["orc"]
    pairs = dict(x_popular_g=torch.from_numpy(P["pg"]), x_generated=torch.from_numpy(P["xg"]))
    ps = [p.clone().requires_grad_(True) for p in P["params"]]
    out = orc.vae_forward(ps, P["X"], P["keep"], 0.75, P["eps"], 1.0, 0.1)
    y = orc.disc_forward(P["E"], P["dparams"], pairs["x_popular_g"], pairs["x_generated"], None, 1.0).detach()
    cnt = len(P["xg"])
    loss = out["neg_ELBO"] + orc.gan_term(out["probs"], P["mask"], y, 1.0, cnt)
    want = [g.detach().numpy() for g in torch.autograd.grad(loss, ps)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert abs(ret["sums"][2] - cnt) < 1e-9
    for g, w in zip(ret["grads"], want):
        assert np.abs(g - w).max() <= 1e-5 * (np.abs(w).max() + 1e-12) + 1e-8


def test_shard_range_partitions_batches():
    dmod = importlib.import_module("long-tail-gan_b200.dist")
    seen = []
    for r in range(4):
        f, c = dmod.shard_range(274, r, 4)
        seen += list(range(f, f + c))
    assert seen == list(range(272)) and len(set(seen)) == 272


# ----------------------------------------------------------------------------------------------------------------------------------
# catalog-sharded (vocab-parallel) formulation, world size 2 on gloo: the exchanges vocab_parallel.CatalogShardedEngine performs
# (SURVEY 8e), with the package's own host functions (shard_bounds, shard_tables, combine_softmax_stats, merge_topk), against the
# single-process oracle on the whole catalog
# ----------------------------------------------------------------------------------------------------------------------------------
def _vp_problem():
    from oracle import ltgan_oracle as orc
    import helpers
    rng = np.random.RandomState(8)
    I, B = 203, 12          # 203 items: the two shards are unequal (104 + 99)
    tabs = helpers.synth_side_tables(rng, B, I, mean_nnz=9)
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], 0, B, I))
    params = orc.init_vae_params(I, seed=9)
    params[3] = params[3] * 4
    keep = torch.from_numpy(rng.rand(B, I) < 0.75)
    # sampled items per user (any items will do for the sum of sampled softmax probabilities)
    samp = [np.sort(rng.choice(I, size=3, replace=False)) for _ in range(B)]
    # ranking scores with ties inside and ACROSS the shards
    scores = np.round(rng.randn(B, I) * 2) / 2
    return dict(orc=orc, tabs=tabs, X=X, params=params, keep=keep, samp=samp, scores=scores.astype(np.float32), I=I, B=B)


def _vp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vp = importlib.import_module("long-tail-gan_b200.vocab_parallel")
    P = _vp_problem()
    I, B = P["I"], P["B"]
    lo, hi = vp.shard_bounds(I, world)[rank]
    st = vp.shard_tables(P["tabs"], lo, hi)
    W_q0, W_q1, W_p0, W_p1, b_q0, b_q1, b_p0, b_p1 = P["params"]
    # encoder: the shard's CSR rows (shard-local ids) x the shard's rows of W_enc, normalised by the WHOLE row's norm -> all-reduce
    Xs = torch.zeros(B, hi - lo)
    for u in range(B):
        Xs[u, st["indices"][st["indptr"][u]: st["indptr"][u + 1]]] = 1.0
    assert torch.equal(Xs, P["X"][:, lo:hi])                                  # shard_tables keeps exactly the shard's interactions
    h = Xs * torch.from_numpy(st["row_rnorm"][:B])[:, None] * P["keep"][:, lo:hi].float() / 0.75
    pre = h @ W_q0[lo:hi]
    dist.all_reduce(pre)
    h1 = torch.tanh(pre + b_q0)
    ml = h1 @ W_q1 + b_q1                                                     # middle: replicated
    h2 = torch.tanh(ml[:, :200] @ W_p0 + b_p0)
    logits = h2 @ W_p1[:, lo:hi] + b_p1[lo:hi]                                # the shard's logits
    lse_r = torch.logsumexp(logits, dim=1)
    xw_r = (logits * Xs).sum(1) - Xs.sum(1) * lse_r                           # local pass: sum x (logit - lse_r)
    su_r = torch.zeros(B)
    for u in range(B):
        own = [i - lo for i in P["samp"][u] if lo <= i < hi]
        su_r[u] = torch.exp(logits[u, own] - lse_r[u]).sum() if own else 0.0
    mine = torch.stack([lse_r, Xs.sum(1), su_r, torch.zeros(B)], dim=1)
    sa = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(sa, mine)
    lse, nx, su = vp.combine_softmax_stats(torch.stack(sa))
    # NLL: local part + sum x_r (lse_r - lse), summed over the shards (vocab_parallel.g_step)
    nll_part = -(xw_r + Xs.sum(1) * (lse_r - lse)).sum()
    nll = nll_part.clone()
    dist.all_reduce(nll)
    # ranking: local top-k of the shard (score desc, local id asc) -> all-gather -> merge
    k = 20
    loc = torch.from_numpy(P["orc"].topk_indices(P["scores"][:, lo:hi], k).copy())
    vals = torch.gather(torch.from_numpy(P["scores"][:, lo:hi].copy()), 1, loc)
    gids = loc + lo
    av = [torch.empty_like(vals) for _ in range(world)]; ag = [torch.empty_like(gids) for _ in range(world)]
    dist.all_gather(av, vals.contiguous()); dist.all_gather(ag, gids.contiguous())
    top = vp.merge_topk(torch.cat(av, 1), torch.cat(ag, 1), k)
    if rank == 0:
        ret["lse"] = lse.numpy(); ret["nx"] = nx.numpy(); ret["su"] = su.numpy(); ret["nll"] = float(nll) / B; ret["top"] = top.numpy()
        ret["bounds"] = vp.shard_bounds(I, world)
    dist.destroy_process_group()


def test_two_shard_catalog_formulation_equals_single_process():
    P = _vp_problem()
    orc = P["orc"]
    out = orc.vae_forward(P["params"], P["X"], P["keep"], 0.75, None, 0.0, 0.0)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_vp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["bounds"] == [(0, 104), (104, 203)]
    want_lse = torch.logsumexp(out["logits"], dim=1).numpy()
    assert np.abs(ret["lse"] - want_lse).max() < 1e-5
    assert np.array_equal(ret["nx"], P["X"].sum(1).numpy())
    want_su = np.asarray([float(out["probs"][u, P["samp"][u]].sum()) for u in range(P["B"])])
    assert np.abs(ret["su"] - want_su).max() < 1e-6
    assert abs(ret["nll"] - float(out["neg_ll"])) < 1e-5 * abs(float(out["neg_ll"]))
    assert np.array_equal(ret["top"], orc.topk_indices(P["scores"], 20))      # ties across the shard boundary: lowest global id first
