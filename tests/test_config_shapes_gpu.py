"""The other BASELINE.json configurations as parity cases (not bench lines): one GAN step at the Netflix-, MSD- and
ML-20M-shaped catalogs on a few hundred synthetic users, generator loss checked against the CPU oracle with the same
injected randomness, sampled pairs checked for validity, evaluation smoke-checked."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ltgan_oracle as orc  # noqa: E402
from oracle import philox  # noqa: E402
import helpers  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["netflix", "msd", "ml20m"])
def test_one_step_at_config_shape(name):
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    N, I, deg = syn.CONFIGS[name]
    B, seed = 96, 4711
    tabs = syn.make_config(name, n_users=2 * B)
    assert tabs["indices"].max() < I and (np.diff(tabs["indptr"]) >= 1).all()
    params = orc.init_vae_params(I, seed=1)
    params[3] = params[3] * 4.0
    E, dparams = orc.init_disc_params(I, 100, 150, 250, 300, seed=2)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=1)
    vae.set_params(params); vae.reset_optimizer()
    disc = dis.Discriminator(I, I, 100, 150, 250, 300, seed=1)
    disc.set_params(E, dparams)
    data = eng.TrainData(batch_size=B, **tabs)
    engine = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=seed, lr=1e-4, lam=1.0, use_graphs=False, max_active=data.max_active)
    bi = 1
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
    b0 = bt["b0"]
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], b0, b0 + B, I))
    idx = (np.uint64(bt["uid0"]) + np.arange(B, dtype=np.uint64))[:, None] * np.uint64(I) + np.arange(I, dtype=np.uint64)[None, :]
    keep = torch.from_numpy(philox.keep_mask(seed, philox.STREAM_ENC_DROPOUT, step, idx, 0.75))
    ref = orc.vae_forward(params, X, keep, 0.75, eps, 1.0, got["anneal"])
    assert abs(got["neg_ll"] - float(ref["neg_ll"])) < 1e-3 * abs(float(ref["neg_ll"])), (got["neg_ll"], float(ref["neg_ll"]))
    assert abs(got["vae_loss"] - float(ref["neg_ELBO"])) < 1e-3 * abs(float(ref["neg_ELBO"]))
    assert np.isfinite(d_loss) and got["cnt"] > 0 and np.isfinite(got["gan_loss"])
    # generated pairs: niche item from the user's candidate set, partner from the user's popular items
    Pr = bt["Pr"]
    sp = bt["samp_ptr"].cpu().numpy(); niche = bt["pair_niche"].cpu().numpy()[Pr:]; pop = bt["pair_pop"].cpu().numpy()[Pr:]
    for u in range(0, B, 7):
        g = b0 + u
        cand = tabs["cand_items"][tabs["cand_ptr"][g]:tabs["cand_ptr"][g + 1]]
        pops = tabs["pop_items"][tabs["pop_ptr"][g]:tabs["pop_ptr"][g + 1]]
        assert np.isin(niche[sp[u]:sp[u + 1]], cand).all() and np.isin(pop[sp[u]:sp[u + 1]], pops).all()
    tr_p, tr_i, te_p, te_i = syn.make_eval_split(64, I, deg)
    m = engine.evaluate(tr_p, tr_i, te_p, te_i, k=100, recall_ks=(20, 50))
    assert 0.0 <= np.mean(m["ndcg@100"]) <= 1.0 and len(m["recall@50"]) == int((np.diff(te_p) > 0).sum())
