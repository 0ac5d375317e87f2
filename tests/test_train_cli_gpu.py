"""GPU end-to-end checks of the drop-in CLI path on the bundled-dataset fixture (tests/golden/askubuntu_sample.npz, produced
by the reference's loaders): one short training run through train.train_GAN, checkpoint, test.test_GAN, and the compat
metric functions against the reference's known answers."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "askubuntu_sample.npz")
pytestmark = pytest.mark.gpu


def test_eval_functions_match_reference_known_answers():
    """eval_functions.py known answers computed with the VERBATIM reference module (make_golden.py), tie-light scores."""
    from scipy import sparse
    ef = importlib.import_module("long-tail-gan_b200.eval_functions")
    g = np.load(GOLD)
    n_items = 1000
    ip, idx = g["vad_tr_indptr"].astype(np.int64), g["vad_tr_indices"].astype(np.int64)
    vtr = sparse.csr_matrix((np.ones(len(idx)), idx, ip), shape=(len(ip) - 1, n_items))
    ip, idx = g["vad_te_indptr"].astype(np.int64), g["vad_te_indices"].astype(np.int64)
    vte = sparse.csr_matrix((np.ones(len(idx)), idx, ip), shape=(len(ip) - 1, n_items))
    rnd = np.random.RandomState(0).rand(vtr.shape[0], n_items).astype(np.float32)
    pred = rnd.copy()
    pred[vtr.nonzero()] = -np.inf
    nd = ef.NDCG_binary_at_k_batch(pred, vte, k=100)
    r20, _ = ef.Recall_at_k_batch(pred, vte, k=20)
    r50, _ = ef.Recall_at_k_batch(pred, vte, k=50)
    assert len(nd) == int(g["ka_rand_vad"][3])
    assert abs(np.mean(nd) - g["ka_rand_vad"][0]) < 1e-9
    assert abs(np.mean(r20) - g["ka_rand_vad"][1]) < 1e-7
    assert abs(np.mean(r50) - g["ka_rand_vad"][2]) < 1e-7
    assert np.allclose(nd[:64], g["ka_block_ndcg"], atol=1e-12)
    # popularity scores (heavy ties): the tie order is unspecified in the reference, so only closeness is required
    pop = np.bincount(g["train_indices"].astype(np.int64), minlength=n_items).astype(np.float32)
    predp = np.tile(pop, (vtr.shape[0], 1))
    predp[vtr.nonzero()] = -np.inf
    assert abs(np.mean(ef.NDCG_binary_at_k_batch(predp, vte, k=100)) - g["ka_pop_vad"][0]) < 2e-3


def test_train_and_test_cli_roundtrip(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    train = importlib.import_module("long-tail-gan_b200.train")
    test = importlib.import_module("long-tail-gan_b200.test")
    cfg = dict(h0_size=100, h1_size=150, h2_size=250, h3_size=300, NUM_EPOCH=8, NUM_SUB_EPOCHS=1, BATCH_SIZE=100, DISPLAY_ITER=50,
               LEARNING_RATE=1e-3, to_restore=0, model_name="LT_GAN", dataset=GOLD, GANLAMBDA=1.0)
    out = train.train_GAN(max_epochs=2, quiet=True, seed=3, **cfg)
    h = out["history"]
    assert len(h) == 2 and all(np.isfinite(x["ndcg"]) for x in h)
    # an untrained VAE-CF scores ~0.03 NDCG@100 on this split; two short epochs at lr 1e-3 must already rank far better
    assert h[-1]["ndcg"] > 0.12, h
    ck = os.path.join("chkpt", "askubuntu_sample_LT_GAN_1.0", "model_1")
    assert os.path.exists(ck)
    n100, r20, r50 = test.test_GAN(output_path=ck, quiet=True, **cfg)
    assert abs(n100 - h[-1]["ndcg"]) < 0.03 and r50 > r20 > 0


def test_resume_reproduces_uninterrupted_run(tmp_path, monkeypatch):
    """to_restore = 1 (parsed and ignored by the reference, train.py:374): a run stopped after epoch 0 and resumed from its checkpoint
    (weights, Adam moments, shared step counter, RNG step, shuffle sequence) ends where the uninterrupted 2-epoch run ends -- up to
    the summation order of float atomics."""
    train = importlib.import_module("long-tail-gan_b200.train")
    cfg = dict(h0_size=100, h1_size=150, h2_size=250, h3_size=300, NUM_EPOCH=8, NUM_SUB_EPOCHS=1, BATCH_SIZE=100, DISPLAY_ITER=50,
               LEARNING_RATE=1e-3, model_name="LT_GAN", dataset=GOLD, GANLAMBDA=1.0)
    # one initial state for both runs (the reference leaves the discriminator initialiser unseeded, discriminator.py:14-41)
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    vae0 = gen.MultiVAE([200, 600, 1000], lam=0.0, random_seed=98765); vae0.init_weights(98765)
    disc0 = dis.Discriminator(1000, 1000, 100, 150, 250, 300, seed=11)
    init = ([t.detach().clone().cpu() for t in vae0.params], disc0.E.detach().clone().cpu(), [t.cpu() for t in disc0.get_params()])
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    monkeypatch.chdir(tmp_path / "a")
    full = train.train_GAN(max_epochs=2, quiet=True, seed=3, to_restore=0, init=init, **cfg)
    monkeypatch.chdir(tmp_path / "b")
    first = train.train_GAN(max_epochs=1, quiet=True, seed=3, to_restore=0, init=init, **cfg)
    W_start, Wq_start, D_start = first["vae"].WdT.clone(), first["vae"].W_q0.clone(), first["disc"].arena.clone()
    assert os.path.exists(os.path.join("chkpt", "askubuntu_sample_LT_GAN_1.0", "model_0"))
    rest = train.train_GAN(max_epochs=2, quiet=True, seed=3, to_restore=1, init=init, **cfg)   # (the checkpoint replaces `init`)
    assert [h["epoch"] for h in rest["history"]] == [1] and [h["epoch"] for h in full["history"]] == [0, 1]
    assert torch.equal(rest["engine"].words.cpu()[:3], full["engine"].words.cpu()[:3])       # rng step, Adam t, G-update count
    assert abs(rest["history"][0]["ndcg"] - full["history"][1]["ndcg"]) < 1e-2
    moved = (full["vae"].WdT.float() - W_start.float()).norm().item()          # what epoch 1 did to the decoder weights
    assert (full["vae"].WdT.float() - rest["vae"].WdT.float()).norm().item() < 0.1 * moved   # (lost Adam moments or dropout steps would show as tens of percent)
    moved_q = (full["vae"].W_q0.float() - Wq_start.float()).norm().item()
    assert (full["vae"].W_q0.float() - rest["vae"].W_q0.float()).norm().item() < 0.1 * moved_q
    moved_d = (full["disc"].arena - D_start).norm().item()
    assert (full["disc"].arena - rest["disc"].arena).norm().item() < 0.25 * moved_d
    # (and the resumed epoch really continued from the checkpoint: it moved the weights of epoch 0 about as far as the full run's epoch 1)
    assert (rest["vae"].WdT.float() - W_start.float()).norm().item() > 0
