"""Single-GPU shape check at the 1,000,000-item catalog of BASELINE.json `configs[4]` (the catalog-sharded multi-GPU layout is
future work; this verifies that every kernel of the path handles a 1 M-item catalog: 64-bit offsets, TMA extents, the
125 KB seen-item bitmap of the top-k kernel, split-K over 15,625 k-blocks). One GAN step on 32 users, generator loss vs the
CPU oracle, plus an evaluation pass.   python tools/check_1m_items.py [n_items]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    I = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    B, seed = 32, 99
    from oracle import ltgan_oracle as orc
    from oracle import philox
    import helpers
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    t0 = time.time()
    indptr, indices = syn.make_interactions(2 * B, I, 50.0, seed=7)
    tabs = syn.make_side_tables(indptr, indices, I, seed=7)
    params = orc.init_vae_params(I, seed=1)
    params[3] = params[3] * 20.0
    E, dparams = orc.init_disc_params(I, 100, 150, 250, 300, seed=2)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=1)
    vae.set_params(params); vae.reset_optimizer()
    disc = dis.Discriminator(I, I, 100, 150, 250, 300, seed=1)
    disc.set_params(E, dparams)
    data = eng.TrainData(batch_size=B, **tabs)
    engine = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=seed, lr=1e-4, lam=1.0, use_graphs=False, max_active=data.max_active)
    print("setup %.1fs, HBM in use %.1f GB" % (time.time() - t0, torch.cuda.memory_allocated() / 2 ** 30))
    bi = 0
    bt = data.batches[bi]
    engine.phase_a(data, bi)
    engine.d_step(data, bi)
    d_loss = engine.last_losses(B)["d_loss"]
    eps = torch.randn(B, 200, generator=torch.Generator().manual_seed(0))
    engine.eps_inject = eps.cuda()
    engine.g_step(data, bi)
    torch.cuda.synchronize()
    got = engine.last_losses(B)
    step = int(engine.words[0].item())
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], 0, B, I))
    idx = np.arange(B, dtype=np.uint64)[:, None] * np.uint64(I) + np.arange(I, dtype=np.uint64)[None, :]
    keep = torch.from_numpy(philox.keep_mask(seed, philox.STREAM_ENC_DROPOUT, step, idx, 0.75))
    ref = orc.vae_forward(params, X, keep, 0.75, eps, 1.0, got["anneal"])
    rel = abs(got["vae_loss"] - float(ref["neg_ELBO"])) / abs(float(ref["neg_ELBO"]))
    print("I=%d: vae_loss %.4f vs oracle %.4f (rel %.2e), d_loss %.2f, generated pairs %d" % (I, got["vae_loss"], float(ref["neg_ELBO"]), rel, d_loss, int(got["cnt"])))
    assert rel < 1e-3 and np.isfinite(d_loss) and got["cnt"] > 0
    tr_p, tr_i, te_p, te_i = syn.make_eval_split(64, I, 50.0)
    m = engine.evaluate(tr_p, tr_i, te_p, te_i, k=100, recall_ks=(20, 50))
    print("eval ok: NDCG@100 %.4f over %d users" % (float(np.mean(m["ndcg@100"])), len(m["ndcg@100"])))
    print("check_1m_items ok")


if __name__ == "__main__":
    main()
