"""The BASELINE.json configurations as parity cases: one GAN step at the Netflix-, MSD- and ML-20M-shaped catalogs (96 users, and
the headline batch of 500 users at the ML-20M shape), every loss of the step -- d_loss, neg_ll, neg_ELBO, the adversarial term and
g_loss -- checked against the CPU oracle with the same injected randomness (north_star: 1e-3 relative; 2e-2 for the small
adversarial term), sampled pairs checked for validity, evaluation smoke-checked."""
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


def _disc_masks(seed, step, n_rows, disc):
    return [torch.from_numpy(philox.hash_keep_mask(seed, philox.STREAM_DISC_DROPOUT + layer, step, n_rows, n, ld, 0.7))
            for layer, (n, ld) in enumerate(((disc.h1, disc.ld1), (disc.h2, disc.ld2), (disc.h3, disc.ld3)))]


@pytest.mark.parametrize("name,B", [("netflix", 96), ("msd", 96), ("ml20m", 96), ("ml20m", 500)])
def test_one_step_at_config_shape(name, B):
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    N, I, deg = syn.CONFIGS[name]
    seed = 4711
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
    dparams_before = [p.clone().cpu() for p in disc.d_params]
    engine.d_step(data, bi)
    torch.cuda.synchronize()
    d_loss = engine.last_losses(B)["d_loss"]
    step_d = int(engine.words[0].item()); t_d = int(engine.words[1].item())
    # oracle D update on the device's pairs with the mirrored dropout bits (train.py:142, 300)
    Pr, K, P = bt["Pr"], bt["K"], bt["P"]
    niche = bt["pair_niche"].cpu().numpy().astype(np.int64); pop = bt["pair_pop"].cpu().numpy().astype(np.int64)
    lab = bt["label"].cpu().numpy()
    gen_rows = np.nonzero(lab[Pr:] > 0)[0]
    pairs = dict(x_popular_n=torch.from_numpy(pop[:Pr]), x_niche=torch.from_numpy(niche[:Pr]),
                 x_popular_g=torch.from_numpy(pop[Pr:][gen_rows]), x_generated=torch.from_numpy(niche[Pr:][gen_rows]))
    masks = _disc_masks(seed, step_d, P, disc)
    dps = [p.clone() for p in dparams_before]
    ref_d, _ = orc.d_step(E, dps, [torch.zeros_like(p) for p in dps], [torch.zeros_like(p) for p in dps], pairs, [m[:Pr] for m in masks],
                          [m[Pr:][gen_rows] for m in masks], 0.7, orc.tf_adam_lr_t(1e-4, t_d))
    assert abs(d_loss - ref_d) < 1e-3 * abs(ref_d), (d_loss, ref_d)
    dparams_after = [p.clone().cpu() for p in disc.d_params]      # the G update sees the UPDATED discriminator (train.py:326)
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
    sp = bt["samp_ptr"].cpu().numpy()
    rows = np.repeat(np.arange(B), np.diff(sp))
    mask = torch.zeros(B, I)
    mask[rows[lab[Pr:] > 0], niche[Pr:][lab[Pr:] > 0]] = 1.0
    cnt = int((lab[Pr:] > 0).sum())
    m_gen = [m[gen_rows] for m in _disc_masks(seed, step, K, disc)]
    ps = [torch.as_tensor(p).clone() for p in params]
    ref = orc.g_step(ps, [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], E, dparams_after, X, keep, 0.75, eps,
                     got["anneal"], mask, pairs, m_gen, 0.7, 1.0, cnt, 1e-4)
    assert got["cnt"] == cnt and cnt > 0
    for key, tol in (("neg_ll", 1e-3), ("vae_loss", 1e-3), ("g_loss", 1e-3), ("gan_loss", 2e-2)):
        assert abs(got[key] - ref[key]) <= tol * abs(ref[key]) + 1e-9, (key, got[key], ref[key])
    # generated pairs: niche item from the user's candidate set, partner from the user's popular items
    niche = niche[Pr:]; pop = pop[Pr:]
    for u in range(0, B, 7):
        g = b0 + u
        cand = tabs["cand_items"][tabs["cand_ptr"][g]:tabs["cand_ptr"][g + 1]]
        pops = tabs["pop_items"][tabs["pop_ptr"][g]:tabs["pop_ptr"][g + 1]]
        assert np.isin(niche[sp[u]:sp[u + 1]], cand).all() and np.isin(pop[sp[u]:sp[u + 1]], pops).all()
    tr_p, tr_i, te_p, te_i = syn.make_eval_split(64, I, deg)
    m = engine.evaluate(tr_p, tr_i, te_p, te_i, k=100, recall_ks=(20, 50))
    assert 0.0 <= np.mean(m["ndcg@100"]) <= 1.0 and len(m["recall@50"]) == int((np.diff(te_p) > 0).sum())
