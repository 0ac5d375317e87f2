"""MultiDAE (Codes/Base_Recommender/MultiVAE.py:11-92), the reference's other base recommender, through the same engine: one GAN
step (phase A, D update, G update) against the CPU oracle with the same injected randomness -- losses within 1e-3 relative (2e-2 for
the small adversarial term), gradients within 5e-2 relative Frobenius (bf16 operands) --, the ranking evaluation, and the
generator.py-style wrapper contract."""
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

I, N, BATCH, SEED = 1000, 200, 100, 777


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu().reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_dae_gan_step_matches_oracle():
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    ops = importlib.import_module("long-tail-gan_b200.ops")
    tabs = helpers.synth_side_tables(np.random.RandomState(9), N, I)
    params = orc.init_dae_params(I, seed=4321)
    params[3] = params[3] * 3.0
    E, dparams = orc.init_disc_params(I, 100, 150, 250, 300, seed=5)
    dae = gen.MultiDAE([200, 600, I], lam=0.0, random_seed=1)
    assert dae.is_dae and dae.view("W_q1").shape == (600, 200) and [tuple(p.shape) for p in dae.params][1] == (600, 200)
    dae.set_params(params); dae.reset_optimizer()
    disc = dis.Discriminator(I, I, 100, 150, 250, 300, seed=1)
    disc.set_params(E, dparams)
    data = eng.TrainData(batch_size=BATCH, **tabs)
    e = eng.GanEngine(dae, disc, data.max_B, data.max_P, seed=SEED, lr=1e-4, lam=1.0, use_graphs=False, max_active=data.max_active)
    bi = 1
    bt = data.batches[bi]
    b0, B, Pr, K, P = bt["b0"], bt["B"], bt["Pr"], bt["K"], bt["P"]
    e.phase_a(data, bi)
    e.d_step(data, bi)
    torch.cuda.synchronize()
    dparams_after = [p.clone().cpu() for p in disc.d_params]
    e.g_step(data, bi, update=False)
    torch.cuda.synchronize()
    step = int(e.words[0].item())
    got = e.last_losses(B)
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], b0, b0 + B, I))
    idx = (np.uint64(bt["uid0"]) + np.arange(B, dtype=np.uint64))[:, None] * np.uint64(I) + np.arange(I, dtype=np.uint64)[None, :]
    keep = torch.from_numpy(philox.keep_mask(SEED, philox.STREAM_ENC_DROPOUT, step, idx, 0.75))
    niche = bt["pair_niche"].cpu().numpy().astype(np.int64); pop = bt["pair_pop"].cpu().numpy().astype(np.int64)
    lab = bt["label"].cpu().numpy()
    gen_rows = np.nonzero(lab[Pr:] > 0)[0]
    pairs = dict(x_popular_n=torch.from_numpy(pop[:Pr]), x_niche=torch.from_numpy(niche[:Pr]),
                 x_popular_g=torch.from_numpy(pop[Pr:][gen_rows]), x_generated=torch.from_numpy(niche[Pr:][gen_rows]))
    m_gen = [torch.from_numpy(philox.hash_keep_mask(SEED, philox.STREAM_DISC_DROPOUT + layer, step, K, n, ld, 0.7))[gen_rows]
             for layer, (n, ld) in enumerate(((disc.h1, disc.ld1), (disc.h2, disc.ld2), (disc.h3, disc.ld3)))]
    sp = bt["samp_ptr"].cpu().numpy()
    rows = np.repeat(np.arange(B), np.diff(sp))
    mask = torch.zeros(B, I)
    mask[rows[lab[Pr:] > 0], niche[Pr:][lab[Pr:] > 0]] = 1.0
    cnt = int((lab[Pr:] > 0).sum())
    ps = [torch.as_tensor(p).clone() for p in params]
    ref = orc.g_step(ps, [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], E, dparams_after, X, keep, 0.75, None, 0.0,
                     mask, pairs, m_gen, 0.7, 1.0, cnt, 1e-4, dae=True)
    assert got["cnt"] == cnt and cnt > 0 and got["KL"] == 0.0
    for key, tol in (("neg_ll", 1e-3), ("vae_loss", 1e-3), ("g_loss", 1e-3), ("gan_loss", 2e-2)):
        assert abs(got[key] - ref[key]) <= tol * abs(ref[key]) + 1e-9, (key, got[key], ref[key])
    dW0 = torch.zeros(I, 600, device="cuda")
    ops.enc_wgrad_expand(dW0, I, bt["slot_of_item"], e.G_enc)
    torch.cuda.synchronize()
    dev_grads = [dW0, dae.view("W_q1", "g"), dae.view("W_p0", "g"), e.dWdT.t(), dae.view("b_q0", "g"), dae.view("b_q1", "g"),
                 dae.view("b_p0", "g"), dae.view("b_p1", "g")]
    for name, g_dev, g_ref in zip("W0 W1 W2 W3 b0 b1 b2 b3".split(), dev_grads, ref["grads"]):
        assert rel(g_dev, g_ref) < 5e-2, (name, rel(g_dev, g_ref))
    # a full update + graph replay run, and the evaluation path
    e2 = eng.GanEngine(dae, disc, data.max_B, data.max_P, seed=SEED, lr=1e-4, lam=1.0, use_graphs=True, max_active=data.max_active)
    for _ in range(2):
        for b in range(len(data.batches)):
            e2.run_step(data, b)
    torch.cuda.synchronize()
    assert torch.isfinite(dae.WdT).all() and torch.isfinite(dae.W_q0).all() and torch.isfinite(dae.small).all()
    rng = np.random.RandomState(3)
    tr_ptr, tr_idx, te_ptr, te_idx = [0], [], [0], []
    for u in range(60):
        items = rng.choice(I, rng.randint(6, 30), replace=False)
        cut = max(1, len(items) // 5)
        te_idx.append(np.sort(items[:cut])); tr_idx.append(np.sort(items[cut:]))
        te_ptr.append(te_ptr[-1] + cut); tr_ptr.append(tr_ptr[-1] + len(items) - cut)
    m = e2.evaluate(np.asarray(tr_ptr), np.concatenate(tr_idx), np.asarray(te_ptr), np.concatenate(te_idx), k=100, recall_ks=(20, 50), keep=1.0)
    pd = [p.detach().cpu().float().contiguous() for p in dae.params]
    pd[:4] = [p.bfloat16().float() for p in pd[:4]]
    Xe = torch.from_numpy(helpers.dense_rows(np.asarray(tr_ptr), np.concatenate(tr_idx), 0, 60, I))
    pred = orc.dae_forward(pd, Xe, None, 1.0)["probs"].numpy().copy()
    pred[Xe.numpy().nonzero()] = -np.inf
    from scipy import sparse
    held = sparse.csr_matrix((np.ones(len(np.concatenate(te_idx))), np.concatenate(te_idx), np.asarray(te_ptr)), shape=(60, I))
    assert abs(np.mean(m["ndcg@100"]) - np.mean(orc.ndcg_binary_at_k_batch(pred, held, k=100))) < 5e-3


def test_dae_wrapper_contract(tmp_path):
    gen = importlib.import_module("long-tail-gan_b200.generator")
    (tmp_path / "unique_item_id.txt").write_text("".join("%d\n" % i for i in range(321)))
    model, out, loss, params, p_dims, total_anneal_steps, anneal_cap = gen.generator_DAECF(str(tmp_path))
    assert p_dims == [200, 600, 321] and len(params) == 8 and tuple(params[1].shape) == (600, 200) and tuple(params[3].shape) == (600, 321)
    assert out.name == "item_prob_dist" and loss.name == "neg_ELBO" and total_anneal_steps == 0 and anneal_cap == 0.0
