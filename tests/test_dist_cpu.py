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
