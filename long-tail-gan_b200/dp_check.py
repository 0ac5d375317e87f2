"""Correctness check of the data-parallel step (user-sharded batches, row-sharded optimizer, peer-memory or NCCL exchange) against
the single-GPU engine on the SAME global batch -- run inside an initialised process group (bench.py at N > 1, tests/test_dp_gpu.py,
tools/dp_check.py). The reference has no multi-GPU path; what must hold is that splitting train.py:192-329's batch over
N ranks changes nothing but float summation order:

  * every rank's bf16 weight shadows (what its next forward reads) are bit-identical to rank 0's and to the bf16 rounding of the
    gathered fp32 masters -- a missing barrier or a torn all-gather shows up here;
  * after `steps` A+D+G steps from identical initial weights the losses of the last step (NLL, KL, sum of sampled probabilities,
    pair count, d_loss, sum y) equal the single-GPU replay, and the weight DISPLACEMENT (w - w0) of W_dec, W_enc and the discriminator
    agrees within bf16/atomic-order noise.

Encoder dropout, eps and the niche sampling are keyed by the GLOBAL user id, so they are identical in both runs; the discriminator's
dropout is keyed by the local pair row, so the check runs with keep_d = 1 (dropout off in the discriminator)."""
import importlib

import numpy as np
import torch

H0, H1, H2, H3 = 100, 150, 250, 300


def _mods():
    pkg = __name__.rsplit(".", 1)[0]
    return (importlib.import_module(pkg + ".generator"), importlib.import_module(pkg + ".discriminator"),
            importlib.import_module(pkg + ".engine"), importlib.import_module(pkg + ".ops"))


def run_check(tabs, n_items, batch_per_rank, rank, world, steps=2, lr=1e-3, seed=11, init_seed=98765, tol=0.05, use_graphs=False):
    """tabs: side tables whose first batch_per_rank*world users form the global batch. Collective: every rank calls it.
    Returns a dict on every rank (the comparison fields are filled on rank 0)."""
    import torch.distributed as dist
    gen, dis, eng, ops = _mods()
    B, I = int(batch_per_rank), int(n_items)
    dev = torch.device("cuda", torch.cuda.current_device())

    def init_weights():
        vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=init_seed)
        vae.init_weights(init_seed)
        with torch.no_grad():   # a sharper decoder so that the sampled probabilities and the adversarial term are not ~1/I
            vae.WdT.mul_(3.0); vae.refresh_shadows()
        disc = dis.Discriminator(I, I, H0, H1, H2, H3, seed=init_seed + 1)
        return vae, disc

    def build(world_size, batch, first, rk):
        vae, disc = init_weights()
        data = eng.TrainData(batch_size=batch, first_batch=first, max_batches=1, **tabs)
        e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=seed, lr=lr, lam=1.0, keep_d=1.0, use_graphs=use_graphs,
                          world_size=world_size, B_global=B * world, max_active=data.max_active, rank=rk)
        if world_size > 1:
            e.attach_dp_tables(eng.build_dp_shard_tables(data, tabs["indptr"], tabs["indices"], world_size, rk, 1, e.R))
        return vae, disc, data, e

    def losses(e):
        sa = e.scal_all.detach().clone().double()
        return dict(nll=sa[0, ops.S_NLL_SUM], kl=sa[0, ops.S_KL_SUM], sum_p=sa[0, ops.S_SUM_P], sum_y=sa[0, ops.S_SUM_Y], cnt=sa[0, ops.S_CNT],
                    d_loss=sa[1, ops.S_D_LOSS])

    vae, disc, data, e = build(world, B, rank, rank)
    w0 = (vae.WdT.clone(), vae.W_q0.clone(), disc.arena.clone())
    for _ in range(steps):
        e.run_phase_a(data, 0); e.run_d_step(data, 0); e.run_g_step(data, 0)
    torch.cuda.synchronize()
    L = losses(e)
    # local sums -> global (sum_y / cnt are already global after the step's own exchange)
    loc = torch.stack([L["nll"], L["kl"], L["sum_p"], L["d_loss"]]).to(dev)
    dist.all_reduce(loc)
    out = dict(world=world, steps=steps, batch_per_rank=B, n_items=I,
               exchange=("peer memory (%s)" % ("NVLS multicast" if e.peer["dWdT_mc"] else "unicast")) if e.peer is not None else "NCCL collectives")
    # ---- shadows: identical on every rank, and equal to the rounding of the gathered masters
    ref_dec, ref_enc = vae.WdT_b.clone(), vae.W_q0_b.clone()
    dist.broadcast(ref_dec, 0); dist.broadcast(ref_enc, 0)
    diff = torch.tensor([float((vae.WdT_b.float() - ref_dec.float()).abs().max()), float((vae.W_q0_b.float() - ref_enc.float()).abs().max())],
                        device=dev, dtype=torch.float64)
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    e.gather_master()
    ok_round = torch.tensor([float(torch.equal(vae.WdT_b, vae.WdT.bfloat16()) and torch.equal(vae.W_q0_b, vae.W_q0.bfloat16()))], device=dev)
    dist.all_reduce(ok_round, op=dist.ReduceOp.MIN)
    out["shadow_max_abs_diff_vs_rank0"] = dict(W_dec=float(diff[0]), W_enc=float(diff[1]))
    out["shadows_equal_rounded_masters"] = bool(ok_round.item() > 0.5)
    dist.barrier()
    if rank == 0:
        vae1, disc1, data1, e1 = build(1, B * world, 0, 0)
        for _ in range(steps):
            e1.run_phase_a(data1, 0); e1.run_d_step(data1, 0); e1.run_g_step(data1, 0)
        torch.cuda.synchronize()
        L1 = losses(e1)
        glob = dict(nll=float(loc[0]), kl=float(loc[1]), sum_p=float(loc[2]), d_loss=float(loc[3]), sum_y=float(L["sum_y"]), cnt=float(L["cnt"]))
        single = {k: float(v) for k, v in L1.items()}
        rel = {k: abs(glob[k] - single[k]) / max(abs(single[k]), 1e-12) for k in glob}

        def disp(a, b, a0):
            return float(((a - a0) - (b - a0)).norm() / ((b - a0).norm() + 1e-30))
        mism = dict(W_dec=disp(vae.WdT, vae1.WdT, w0[0]), W_enc=disp(vae.W_q0, vae1.W_q0, w0[1]), disc=disp(disc.arena, disc1.arena, w0[2]))
        out.update(losses_dp=glob, losses_single_gpu=single, loss_rel_diff=rel, displacement_mismatch=mism, tol=tol,
                   ok=bool(out["shadows_equal_rounded_masters"] and max(out["shadow_max_abs_diff_vs_rank0"].values()) == 0.0 and
                           glob["cnt"] == single["cnt"] and max(v for k, v in rel.items() if k != "cnt") < 2e-2 and max(mism.values()) < tol))
        del e1, vae1, disc1, data1
    dist.barrier()
    del e, vae, disc, data
    torch.cuda.empty_cache()
    return out


def small_problem(world, batch_per_rank=50, n_items=1000, seed=5):
    """Seeded small side tables (structure of the bundled dataset) for tests/tools: batch_per_rank*world users."""
    rng = np.random.RandomState(seed)
    N, I = batch_per_rank * world, n_items
    n_pop = max(4, I // 10)
    niche_all = np.arange(n_pop, I)
    indptr, indices, pop_ptr, pop_items, n_niche, cand_ptr, cand_items, real_ptr, real_niche, real_pop, eligible = [0], [], [0], [], [], [0], [], [0], [], [], []
    for u in range(N):
        n = int(np.clip(rng.poisson(18), 2, I // 2))
        k_pop = int(np.clip(rng.binomial(n, 0.5), 1, min(n - 1, n_pop)))
        items = np.sort(np.concatenate([rng.choice(n_pop, k_pop, replace=False), rng.choice(niche_all, n - k_pop, replace=False)]))
        indices.append(items); indptr.append(indptr[-1] + len(items))
        pops, niches = rng.permutation(items[items < n_pop]), items[items >= n_pop]
        pop_items.append(pops); pop_ptr.append(pop_ptr[-1] + len(pops))
        eligible.append(True); n_niche.append(len(niches))
        others = np.setdiff1d(niche_all, niches)
        extra = rng.choice(others, min(len(others), max(2 * len(niches), 10 - len(niches))), replace=False)
        c = np.sort(np.concatenate([niches, extra]))
        cand_items.append(c); cand_ptr.append(cand_ptr[-1] + len(c))
        for g in niches:
            real_niche.append(int(g)); real_pop.append(int(pops[rng.randint(len(pops))]))
        real_ptr.append(real_ptr[-1] + len(niches))
    cat = lambda xs: np.concatenate(xs).astype(np.int32)  # noqa: E731
    i32 = lambda a: np.asarray(a, dtype=np.int32)  # noqa: E731
    return dict(n_items=I, indptr=i32(indptr), indices=cat(indices), pop_ptr=i32(pop_ptr), pop_items=cat(pop_items), n_niche=i32(n_niche),
                cand_ptr=i32(cand_ptr), cand_items=cat(cand_items), real_ptr=i32(real_ptr), real_niche=i32(real_niche), real_pop=i32(real_pop),
                eligible=np.asarray(eligible, dtype=bool), item_valid=np.ones(I, dtype=np.uint8))
