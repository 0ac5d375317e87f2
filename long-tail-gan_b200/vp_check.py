"""Correctness check of the catalog-sharded engine (vocab_parallel.py) against the single-GPU engine on the same batch and the same
initial weights: losses of one A + D + G step, weight displacement of the gathered shards, and the merged top-k evaluation.
Runs with any world size (1 = one shard holding the whole catalog: exercises the sharded code path on a single GPU)."""
import importlib

import numpy as np
import torch

H0, H1, H2, H3 = 100, 150, 250, 300


def run_check(n_items, batch, rank, world, steps=2, seed=11, lr=1e-3, n_eval=96, use_graphs=False):
    import torch.distributed as dist
    pkg = __name__.rsplit(".", 1)[0]
    gen = importlib.import_module(pkg + ".generator"); dis = importlib.import_module(pkg + ".discriminator")
    eng = importlib.import_module(pkg + ".engine"); vp = importlib.import_module(pkg + ".vocab_parallel")
    dpc = importlib.import_module(pkg + ".dp_check")
    I, B = int(n_items), int(batch)
    tabs = dpc.small_problem(1, batch_per_rank=B, n_items=I, seed=7)
    g = torch.Generator().manual_seed(123)
    lim = float(np.sqrt(6.0 / (I + 600)))
    params = [(torch.rand(I, 600, generator=g) * 2 - 1) * lim, (torch.rand(600, 400, generator=g) * 2 - 1) * 0.077, (torch.rand(200, 600, generator=g) * 2 - 1) * 0.087,
              (torch.rand(600, I, generator=g) * 2 - 1) * lim * 3.0, torch.randn(600, generator=g) * 0.001, torch.randn(400, generator=g) * 0.001,
              torch.randn(600, generator=g) * 0.001, torch.randn(I, generator=g) * 0.001]

    def new_disc():
        return dis.Discriminator(I, I, H0, H1, H2, H3, seed=4242)

    # ---- sharded run
    data, vae, lo, hi = vp.build_shard(tabs, I, rank, world, B, vae_params=params)
    disc = new_disc()
    e = vp.CatalogShardedEngine(vae, disc, data.max_B, data.max_P, I, lo, rank, world, seed=seed, lr=lr, lam=1.0, keep_d=1.0, max_active=data.max_active,
                                use_graphs=use_graphs)
    for _ in range(steps):
        e.run_phase_a(data, 0); e.run_d_step(data, 0); e.run_g_step(data, 0)
    torch.cuda.synchronize()
    L = e.last_losses(B)
    # evaluation on held-out style users: reuse the training rows split 80/20
    ip = np.asarray(tabs["indptr"], dtype=np.int64); idx = np.asarray(tabs["indices"], dtype=np.int64)
    n_eval = min(n_eval, len(ip) - 1)
    held = np.zeros(len(idx), dtype=bool); held[4::5] = True
    row = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
    m = row < n_eval
    trp = np.concatenate([[0], np.cumsum(np.bincount(row[m & ~held], minlength=n_eval))]); tep = np.concatenate([[0], np.cumsum(np.bincount(row[m & held], minlength=n_eval))])
    ev = (trp, idx[m & ~held], tep, idx[m & held])
    met = e.evaluate(*ev, k=100, recall_ks=(20, 50), keep=1.0)
    # gather the shards (fp32 masters) on every rank
    def gather(t_local, dim_len):
        if world == 1:
            return t_local.clone()
        R = vp.shard_bounds(I, world)[0][1]
        buf = torch.zeros(R, t_local.shape[1], device=t_local.device); buf[: t_local.shape[0]] = t_local
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return torch.cat(outs, 0)[:dim_len]
    WdT = gather(vae.WdT, I); Wq0 = gather(vae.W_q0, I)
    out = dict(world=world, n_items=I, batch=B, steps=steps, graphs=bool(use_graphs), losses_sharded={k: float(v) for k, v in L.items()})
    if rank == 0:
        vae1 = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=1); vae1.set_params(params); vae1.reset_optimizer()
        disc1 = new_disc()
        data1 = eng.TrainData(batch_size=B, **tabs)
        e1 = eng.GanEngine(vae1, disc1, data1.max_B, data1.max_P, seed=seed, lr=lr, lam=1.0, keep_d=1.0, use_graphs=False, max_active=data1.max_active)
        for _ in range(steps):
            e1.run_phase_a(data1, 0); e1.run_d_step(data1, 0); e1.run_g_step(data1, 0)
        torch.cuda.synchronize()
        L1 = e1.last_losses(B)
        met1 = e1.evaluate(*ev, k=100, recall_ks=(20, 50), keep=1.0)
        rel = {k: abs(L[k] - L1[k]) / max(abs(L1[k]), 1e-12) for k in ("neg_ll", "KL", "vae_loss", "gan_loss", "d_loss", "sum_p", "sum_y")}
        W0 = params[3].t().cuda(); Q0 = params[0].cuda(); d0 = new_disc()

        def disp(a, b, a0):
            return float(((a - a0) - (b - a0)).norm() / ((b - a0).norm() + 1e-30))
        mism = dict(W_dec=disp(WdT, vae1.WdT, W0), W_enc=disp(Wq0, vae1.W_q0, Q0), disc=disp(disc.arena, disc1.arena, d0.arena),
                    small=disp(vae.small[: vae._small_off["b_p1"][0]], vae1.small[: vae1._small_off["b_p1"][0]], torch.zeros(1, device="cuda")))
        nd, nd1 = np.asarray(met["ndcg@100"]), np.asarray(met1["ndcg@100"])
        out.update(losses_single_gpu={k: float(v) for k, v in L1.items()}, loss_rel_diff=rel, displacement_mismatch=mism,
                   cnt_equal=bool(L["cnt"] == L1["cnt"]), eval_ndcg=(float(nd.mean()), float(nd1.mean())), eval_users=(len(nd), len(nd1)),
                   eval_ndcg_max_abs_diff=float(np.abs(nd - nd1).max()) if len(nd) == len(nd1) else None)
        out["ok"] = bool(out["cnt_equal"] and max(rel.values()) < 2e-2 and max(v for k, v in mism.items() if k != "small") < 0.05 and mism["small"] < 1e-3
                         and len(nd) == len(nd1) and abs(nd.mean() - nd1.mean()) < 5e-3)
    if world > 1:
        dist.barrier()
    return out
