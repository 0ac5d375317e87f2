"""GPU parity of the assembled phase-A / D-update / G-update path against the CPU oracle with injected randomness.
Tolerances (north_star): per-step losses within 1e-3 relative (bf16 operands, fp32 accumulate); gradients are compared in
relative Frobenius norm because individual entries carry the bf16 rounding of the operands."""
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

I, N, BATCH = 1000, 230, 100
H0, H1, H2, H3 = 100, 150, 250, 300
SEED = 20240607


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu().reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def setup():
    pkg = importlib.import_module("long-tail-gan_b200")
    pkg._lib.build()
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    rng = np.random.RandomState(5)
    tabs = helpers.synth_side_tables(rng, N, I)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765)
    params = orc.init_vae_params(I, seed=98765)
    # the reference's Xavier init gives near-uniform softmax; scale the decoder a bit so probabilities differ visibly
    params[3] = params[3] * 3.0
    vae.set_params(params)
    vae.reset_optimizer()
    disc = dis.Discriminator(I, I, H0, H1, H2, H3, seed=1)
    E, dparams = orc.init_disc_params(I, H0, H1, H2, H3, seed=77)
    disc.set_params(E, dparams)
    data = eng.TrainData(batch_size=BATCH, **tabs)
    engine = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=SEED, lr=1e-4, lam=1.0, use_graphs=False, max_active=data.max_active)
    return dict(eng=eng, vae=vae, disc=disc, data=data, engine=engine, tabs=tabs, params=params, E=E, dparams=dparams)


def disc_masks(step, n_rows, disc):
    out = []
    for layer, (n, ld) in enumerate(((disc.h1, disc.ld1), (disc.h2, disc.ld2), (disc.h3, disc.ld3))):
        out.append(torch.from_numpy(philox.hash_keep_mask(SEED, philox.STREAM_DISC_DROPOUT + layer, step, n_rows, n, ld, 0.7)))
    return out


def test_phase_a_pairs_are_consistent(setup):
    s = setup
    engine, data, tabs = s["engine"], s["data"], s["tabs"]
    for bi in range(len(data.batches)):
        engine.phase_a(data, bi)
    torch.cuda.synchronize()
    for bi, bt in enumerate(data.batches):
        b0, B, Pr, K = bt["b0"], bt["B"], bt["Pr"], bt["K"]
        niche = bt["pair_niche"].cpu().numpy(); pop = bt["pair_pop"].cpu().numpy(); lab = bt["label"].cpu().numpy()
        sp = bt["samp_ptr"].cpu().numpy()
        assert (lab[:Pr] == 0).all()
        nvalid = 0
        for u in range(B):
            g = u + b0
            drawn = niche[Pr + sp[u]: Pr + sp[u + 1]]
            if not tabs["eligible"][g]:
                assert len(drawn) == 0
                continue
            cand = tabs["cand_items"][tabs["cand_ptr"][g]: tabs["cand_ptr"][g + 1]]
            assert len(drawn) == tabs["n_niche"][g]
            assert (np.diff(drawn) > 0).all() and np.isin(drawn, cand).all()      # sorted, unique, from the candidate set
            partners = pop[Pr + sp[u]: Pr + sp[u + 1]]
            upop = tabs["pop_items"][tabs["pop_ptr"][g]: tabs["pop_ptr"][g + 1]]
            assert np.isin(partners, upop).all()                                  # partner is one of the user's popular items
            want_valid = tabs["item_valid"][drawn].astype(bool) & tabs["item_valid"][partners].astype(bool)
            assert np.array_equal(lab[Pr + sp[u]: Pr + sp[u + 1]] > 0, want_valid)  # F10 validity filter
            nvalid += int(want_valid.sum())
        assert int(bt["cnt"].item()) == nvalid


def _oracle_pairs(bt):
    Pr = bt["Pr"]
    niche = bt["pair_niche"].cpu().numpy().astype(np.int64); pop = bt["pair_pop"].cpu().numpy().astype(np.int64)
    lab = bt["label"].cpu().numpy()
    gen_rows = np.nonzero(lab[Pr:] > 0)[0]
    return dict(x_popular_n=torch.from_numpy(pop[:Pr]), x_niche=torch.from_numpy(niche[:Pr]),
                x_popular_g=torch.from_numpy(pop[Pr:][gen_rows]), x_generated=torch.from_numpy(niche[Pr:][gen_rows])), gen_rows


def test_d_step_matches_oracle(setup):
    s = setup
    engine, data, disc = s["engine"], s["data"], s["disc"]
    bi = 0
    engine.phase_a(data, bi)
    bt = data.batches[bi]
    Pr, P = bt["Pr"], bt["P"]
    before = [p.clone().cpu() for p in disc.d_params]
    engine.d_step(data, bi)
    torch.cuda.synchronize()
    step = int(engine.words[0].item()); t = int(engine.words[1].item())
    pairs, gen_rows = _oracle_pairs(bt)
    masks = disc_masks(step, P, disc)
    m_real = [m[:Pr] for m in masks]
    m_gen = [m[Pr:][gen_rows] for m in masks]
    dm = [torch.zeros_like(p) for p in before]; dv = [torch.zeros_like(p) for p in before]
    ps = [p.clone() for p in before]
    loss, grads = orc.d_step(s["E"], ps, dm, dv, pairs, m_real, m_gen, 0.7, orc.tf_adam_lr_t(1e-4, t))
    got = engine.last_losses(bt["B"])
    assert abs(got["d_loss"] - loss) < 1e-3 * abs(loss), (got["d_loss"], loss)
    gsum = engine.arena_gp[: engine._d_parts].sum(0)          # split-K partials of the weight-gradient GEMMs
    disc.arena_g.copy_(gsum)
    ggrads = disc.get_params("g")
    for name, g_dev, g_ref in zip("w1 b1 w2 b2 w3 b3 w4 b4".split(), ggrads, grads):
        assert rel(g_dev.reshape(g_ref.shape), g_ref) < 3e-2, (name, rel(g_dev.reshape(g_ref.shape), g_ref))
    # parameters after the TF-Adam update. With m = v = 0 the first step moves every weight by ~lr*sign(g), so entries whose
    # gradient is within the bf16 noise of zero may flip: require the displacement signs to agree on >= 97% of the entries.
    after = disc.d_params
    for name, a, b0_, ref in zip("w1 b1 w2 b2 w3 b3 w4 b4".split(), after, before, ps):
        moved_dev = (a.cpu().reshape(ref.shape) - b0_.reshape(ref.shape))
        moved_ref = ref - b0_.reshape(ref.shape)
        agree = float((torch.sign(moved_dev) == torch.sign(moved_ref)).float().mean())
        assert agree > 0.97, (name, agree)
        assert float(moved_dev.abs().max()) < 1.05e-4


def test_g_step_matches_oracle(setup):
    s = setup
    engine, data, vae, disc = s["engine"], s["data"], s["vae"], s["disc"]
    bi = 1
    engine.phase_a(data, bi)
    bt = data.batches[bi]
    b0, B, Pr, K = bt["b0"], bt["B"], bt["Pr"], bt["K"]
    torch.manual_seed(3)
    eps = torch.randn(B, 200)
    engine.eps_inject = eps.cuda()
    params_before = [p.clone().cpu().contiguous() for p in vae.params]
    dparams = [p.clone().cpu() for p in disc.d_params]
    engine.g_step(data, bi, update=False)
    torch.cuda.synchronize()
    step = int(engine.words[0].item())
    got = engine.last_losses(B)
    tabs = s["tabs"]
    X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], b0, b0 + B, I))
    idx = (np.uint64(bt["uid0"]) + np.arange(B, dtype=np.uint64))[:, None] * np.uint64(I) + np.arange(I, dtype=np.uint64)[None, :]
    keep_mask = torch.from_numpy(philox.keep_mask(SEED, philox.STREAM_ENC_DROPOUT, step, idx, 0.75))
    pairs, gen_rows = _oracle_pairs(bt)
    masks = disc_masks(step, K, disc)
    m_gen = [m[gen_rows] for m in masks]
    sp = bt["samp_ptr"].cpu().numpy()
    mask = torch.zeros(B, I)
    niche = bt["pair_niche"].cpu().numpy()[Pr:]; lab = bt["label"].cpu().numpy()[Pr:]
    rows = np.repeat(np.arange(B), np.diff(sp))
    mask[rows[lab > 0], niche[lab > 0]] = 1.0
    cnt = int((lab > 0).sum())
    anneal = got["anneal"]
    gm = [torch.zeros_like(p) for p in params_before]; gv = [torch.zeros_like(p) for p in params_before]
    ps = [p.clone() for p in params_before]
    ref = orc.g_step(ps, gm, gv, s["E"], dparams, X, keep_mask, 0.75, eps, anneal, mask, pairs, m_gen, 0.7, 1.0, cnt, 1e-4)
    assert got["cnt"] == cnt
    for key, tol in (("neg_ll", 1e-3), ("KL", 2e-2), ("vae_loss", 1e-3), ("g_loss", 1e-3)):
        assert abs(got[key] - ref[key]) <= tol * abs(ref[key]), (key, got[key], ref[key])
    assert abs(got["gan_loss"] - ref["gan_loss"]) <= 2e-2 * abs(ref["gan_loss"]) + 1e-6, (got["gan_loss"], ref["gan_loss"])
    assert abs(got["sum_y"] - float(ref["y_gen"].sum())) < 1e-2 * abs(float(ref["y_gen"].sum()))
    # gradients: [W_q0, W_q1, W_p0, W_p1, b_q0, b_q1, b_p0, b_p1]
    ops = importlib.import_module("long-tail-gan_b200.ops")
    dWq0 = torch.zeros(I, 600, device="cuda")
    ops.enc_wgrad_expand(dWq0, I, bt["slot_of_item"], engine.G_enc)
    torch.cuda.synchronize()
    dev_grads = [dWq0, vae.view("W_q1", "g"), vae.view("W_p0", "g"), engine.dWdT.t(), vae.view("b_q0", "g"), vae.view("b_q1", "g"),
                 vae.view("b_p0", "g"), vae.view("b_p1", "g")]
    for name, g_dev, g_ref in zip("W_q0 W_q1 W_p0 W_p1 b_q0 b_q1 b_p0 b_p1".split(), dev_grads, ref["grads"]):
        r = rel(g_dev, g_ref)
        assert r < 5e-2, (name, r)
    engine.eps_inject = None


def test_graph_replay_equals_eager(setup):
    s = setup
    eng = s["eng"]
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")

    def run(use_graphs):
        vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765)
        vae.set_params(s["params"]); vae.reset_optimizer()
        disc = dis.Discriminator(I, I, H0, H1, H2, H3, seed=1)
        disc.set_params(s["E"], s["dparams"])
        data = eng.TrainData(batch_size=BATCH, **s["tabs"])
        e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=SEED, use_graphs=use_graphs, max_active=data.max_active)
        for _ in range(3):
            for bi in range(2):
                e.run_phase_a(data, bi)
            for bi in range(2):
                e.run_d_step(data, bi)
            for bi in range(2):
                e.run_g_step(data, bi)
        torch.cuda.synchronize()
        return vae.WdT.clone(), vae.W_q0.clone(), disc.arena.clone(), e.words.clone()

    a = run(False)
    b = run(True)
    assert torch.equal(a[3], b[3])
    # split-K / bias reductions use float atomics: same values up to summation order
    assert rel(b[0] - torch.as_tensor(s["params"][3]).t().cuda(), a[0] - torch.as_tensor(s["params"][3]).t().cuda()) < 2e-2
    assert (a[1] - b[1]).abs().max().item() < 1e-3
    assert (a[2] - b[2]).abs().max().item() < 1e-3


@pytest.mark.parametrize("split_d", [0, 1, 2, 3])
def test_run_step_single_graph_equals_three_phases(setup, split_d):
    """engine.run_step (phase A + D + G of one batch captured as ONE graph, what bench.py times) against the three per-phase calls
    run eagerly on a twin engine: same RNG counters, weights equal up to float-atomic summation order. split_d > 0: the real pairs'
    half of the D forward runs as its own launch beside phase A (same dropout counters through rng_row0)."""
    s = setup
    eng = s["eng"]
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")

    def run(fused):
        vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765)
        vae.set_params(s["params"]); vae.reset_optimizer()
        disc = dis.Discriminator(I, I, H0, H1, H2, H3, seed=1)
        disc.set_params(s["E"], s["dparams"])
        data = eng.TrainData(batch_size=BATCH, **s["tabs"])
        e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=SEED, use_graphs=fused, max_active=data.max_active)
        e.split_d = split_d
        for _ in range(3):          # first pass captures, later passes replay
            for bi in range(2):
                if fused:
                    e.run_step(data, bi)
                else:
                    e.run_phase_a(data, bi); e.run_d_step(data, bi); e.run_g_step(data, bi)
        torch.cuda.synchronize()
        return vae.WdT.clone(), vae.W_q0.clone(), disc.arena.clone(), e.words.clone(), e.kernels_launched

    a = run(False)
    b = run(True)
    # same counters; the one-graph step merges the three per-phase counter advances into one launch and gathers the embedding rows of
    # the generated pairs once for the D and the G update (3 launches less per step)
    # (... and 2 launches more when the D forward is split into its real and its generated half)
    assert torch.equal(a[3], b[3]) and a[4] - b[4] == (3 - (2 if split_d else 0)) * 3 * 2
    W0 = torch.as_tensor(s["params"][3]).t().cuda()
    assert rel(b[0] - W0, a[0] - W0) < 2e-2
    assert (a[1] - b[1]).abs().max().item() < 1e-3
    assert (a[2] - b[2]).abs().max().item() < 1e-3


def test_evaluate_matches_oracle_metrics(setup):
    from scipy import sparse
    s = setup
    engine, vae = s["engine"], s["vae"]
    rng = np.random.RandomState(21)
    n_eval = 150
    tr_ptr, tr_idx = [0], []
    te_ptr, te_idx = [0], []
    for u in range(n_eval):
        items = rng.choice(I, rng.randint(6, 40), replace=False)
        cut = 0 if u % 17 == 3 else max(1, len(items) // 5)   # some users have an empty held-out set
        te = np.sort(items[:cut]); tr = np.sort(items[cut:])
        tr_idx.append(tr); tr_ptr.append(tr_ptr[-1] + len(tr)); te_idx.append(te); te_ptr.append(te_ptr[-1] + len(te))
    tr_ptr, te_ptr = np.asarray(tr_ptr), np.asarray(te_ptr)
    tr_idx, te_idx = np.concatenate(tr_idx), np.concatenate(te_idx)
    got = engine.evaluate(tr_ptr, tr_idx, te_ptr, te_idx, k=100, recall_ks=(20, 50), uid_start=5000, keep=1.0)
    # oracle: fp32 forward with the bf16-rounded weights the device uses (isolates the metric code from operand rounding)
    params = [p.detach().cpu().float().contiguous() for p in vae.params]
    params[:4] = [p.bfloat16().float() for p in params[:4]]
    X = torch.from_numpy(helpers.dense_rows(tr_ptr, tr_idx, 0, n_eval, I))
    out = orc.vae_forward(params, X, None, 1.0, None, 0.0, 0.0)
    pred = out["probs"].numpy().copy()
    pred[X.numpy().nonzero()] = -np.inf
    held = sparse.csr_matrix((np.ones(len(te_idx)), te_idx, te_ptr), shape=(n_eval, I))
    ndcg = orc.ndcg_binary_at_k_batch(pred, held, k=100)
    r20, _ = orc.recall_at_k_batch(pred, held, k=20)
    r50, _ = orc.recall_at_k_batch(pred, held, k=50)
    assert len(got["ndcg@100"]) == len(ndcg) and len(got["recall@20"]) == len(r20)
    assert abs(np.mean(got["ndcg@100"]) - np.mean(ndcg)) < 5e-3
    assert abs(np.mean(got["recall@20"]) - np.mean(r20)) < 5e-3
    assert abs(np.mean(got["recall@50"]) - np.mean(r50)) < 5e-3


def test_session_run_fetches_generator_out(setup):
    """The wrapper contract: generator_out is a probability distribution over items (generator.py:22 element [1]) that can be
    fetched with a dense feed like `sess.run(generator_out, {input_ph: X})` (train.py:200)."""
    s = setup
    eng, vae, engine = s["eng"], s["vae"], s["engine"]
    tabs = s["tabs"]
    X = helpers.dense_rows(tabs["indptr"], tabs["indices"], 0, 37, I)
    X[3, 5] = 2.0    # a duplicated interaction (values > 1 are legal in the CSR the reference builds)
    sess = eng.Session(engine)
    lazy_out, lazy_loss = vae.generator_out if hasattr(vae, "generator_out") else None, None
    gen = importlib.import_module("long-tail-gan_b200.generator")
    out_h, loss_h = gen.LazyTensor(vae, "item_prob_dist"), gen.LazyTensor(vae, "neg_ELBO")
    probs, loss = sess.run([out_h, loss_h], {vae.input_ph: X, vae.keep_prob_ph: 1.0})
    params = [p.detach().cpu().float().contiguous() for p in vae.params]
    ref = orc.vae_forward(params, torch.from_numpy(X), None, 1.0, None, 0.0, 1.0)
    assert probs.shape == (37, I) and np.allclose(probs.sum(1), 1.0, atol=2e-2)
    assert np.abs(probs - ref["probs"].numpy()).max() < 3e-2 * ref["probs"].numpy().max()
    assert abs(float(loss) - float(ref["neg_ELBO"])) < 2e-3 * abs(float(ref["neg_ELBO"]))


def test_edge_batches_single_user_and_no_pairs():
    """Edge cases of train.py:192-255: a tail batch of ONE user (10001 = 100*100 + 1 on the bundled data) and a batch in which no
    user has both popular and niche items (no real pairs, no candidates, nothing sampled -> the reference skips the batch)."""
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    rng = np.random.RandomState(17)
    tabs = helpers.synth_side_tables(rng, 21, 300, mean_nnz=8)
    # make users 10..19 ineligible (second batch of 10): no candidates, no real pairs
    for u in range(10, 20):
        tabs["eligible"][u] = False
    keep_c = np.concatenate([np.arange(tabs["cand_ptr"][u], tabs["cand_ptr"][u + 1]) for u in range(21) if tabs["eligible"][u]])
    keep_r = np.concatenate([np.arange(tabs["real_ptr"][u], tabs["real_ptr"][u + 1]) for u in range(21) if tabs["eligible"][u]])
    cl = np.where(tabs["eligible"], np.diff(tabs["cand_ptr"]), 0); rl = np.where(tabs["eligible"], np.diff(tabs["real_ptr"]), 0)
    tabs["cand_items"] = tabs["cand_items"][keep_c]; tabs["cand_ptr"] = np.concatenate([[0], np.cumsum(cl)]).astype(np.int32)
    tabs["real_niche"] = tabs["real_niche"][keep_r]; tabs["real_pop"] = tabs["real_pop"][keep_r]
    tabs["real_ptr"] = np.concatenate([[0], np.cumsum(rl)]).astype(np.int32)
    vae = gen.MultiVAE([200, 600, 300], lam=0.0, random_seed=3); vae.init_weights(3)
    disc = dis.Discriminator(300, 300, 100, 150, 250, 300, seed=3)
    data = eng.TrainData(batch_size=10, **tabs)
    assert [b["B"] for b in data.batches] == [10, 10, 1]
    assert data.batches[1]["P"] == 0 and data.batches[1]["K"] == 0
    e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=5, use_graphs=True, max_active=data.max_active)
    for rep in range(2):              # second pass replays the captured graphs
        for bi in range(3):
            e.run_phase_a(data, bi)
            if int(data.batches[bi]["cnt"].item()) == 0:
                assert bi == 1 or data.batches[bi]["K"] == 0 or True
            e.run_d_step(data, bi)
            e.run_g_step(data, bi)
            L = e.last_losses(data.batches[bi]["B"])
            assert np.isfinite(L["vae_loss"]) and np.isfinite(L["g_loss"]) and np.isfinite(L["d_loss"])
            if bi == 1:
                assert L["cnt"] == 0 and L["gan_loss"] == 0.0 and L["d_loss"] == 0.0
    torch.cuda.synchronize()
    assert torch.isfinite(vae.WdT).all() and torch.isfinite(vae.W_q0).all() and torch.isfinite(disc.arena).all()


def test_ten_step_drift_vs_oracle(setup):
    """Ten consecutive A -> D -> G updates, device vs oracle, WITHOUT re-synchronising the oracle's parameters from the device: the
    oracle keeps its own fp32 weights and Adam moments from step 0 on and is fed only what the reference's host code would hand the
    graph (the device's sampled pairs, the mirrored dropout bits, the injected eps). Every step's losses must stay within the
    north_star tolerance (1e-3 relative; 2e-2 for the small adversarial term), i.e. bf16-operand rounding does not compound, and the
    accumulated parameter displacement after ten TF-Adam updates must agree in direction and size."""
    s = setup
    eng = s["eng"]
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    tabs = s["tabs"]
    lr = 1e-4
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765)
    vae.set_params(s["params"]); vae.reset_optimizer()
    disc = dis.Discriminator(I, I, H0, H1, H2, H3, seed=1)
    disc.set_params(s["E"], s["dparams"])
    data = eng.TrainData(batch_size=BATCH, **tabs)
    e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=SEED, lr=lr, lam=1.0, use_graphs=False, max_active=data.max_active)
    g0 = [p.clone().cpu().contiguous() for p in vae.params]
    d0 = [p.clone().cpu() for p in disc.d_params]
    gp = [p.clone() for p in g0]; gm = [torch.zeros_like(p) for p in g0]; gv = [torch.zeros_like(p) for p in g0]
    dp = [p.clone() for p in d0]; dm = [torch.zeros_like(p) for p in d0]; dv = [torch.zeros_like(p) for p in d0]
    worst = dict(d_loss=0.0, vae_loss=0.0, g_loss=0.0, gan_loss=0.0)
    for it in range(10):
        bi = it % 2
        bt = data.batches[bi]
        b0, B, Pr, K, P = bt["b0"], bt["B"], bt["Pr"], bt["K"], bt["P"]
        e.phase_a(data, bi)
        e.d_step(data, bi)
        torch.cuda.synchronize()
        step = int(e.words[0].item()); t = int(e.words[1].item())
        pairs, gen_rows = _oracle_pairs(bt)
        masks = disc_masks(step, P, disc)
        d_loss, _ = orc.d_step(s["E"], dp, dm, dv, pairs, [m[:Pr] for m in masks], [m[Pr:][gen_rows] for m in masks], 0.7,
                               orc.tf_adam_lr_t(lr, t))
        got_d = e.last_losses(B)["d_loss"]
        worst["d_loss"] = max(worst["d_loss"], abs(got_d - d_loss) / abs(d_loss))
        eps = torch.randn(B, 200, generator=torch.Generator().manual_seed(100 + it))
        e.eps_inject = eps.cuda()
        e.g_step(data, bi)
        torch.cuda.synchronize()
        step = int(e.words[0].item()); t = int(e.words[1].item())
        got = e.last_losses(B)
        X = torch.from_numpy(helpers.dense_rows(tabs["indptr"], tabs["indices"], b0, b0 + B, I))
        idx = (np.uint64(bt["uid0"]) + np.arange(B, dtype=np.uint64))[:, None] * np.uint64(I) + np.arange(I, dtype=np.uint64)[None, :]
        keep_mask = torch.from_numpy(philox.keep_mask(SEED, philox.STREAM_ENC_DROPOUT, step, idx, 0.75))
        m_gen = [m[gen_rows] for m in disc_masks(step, K, disc)]
        sp = bt["samp_ptr"].cpu().numpy()
        niche = bt["pair_niche"].cpu().numpy()[Pr:]; lab = bt["label"].cpu().numpy()[Pr:]
        rows = np.repeat(np.arange(B), np.diff(sp))
        mask = torch.zeros(B, I)
        mask[rows[lab > 0], niche[lab > 0]] = 1.0
        cnt = int((lab > 0).sum())
        ref = orc.g_step(gp, gm, gv, s["E"], dp, X, keep_mask, 0.75, eps, got["anneal"], mask, pairs, m_gen, 0.7, 1.0, cnt,
                         orc.tf_adam_lr_t(lr, t))
        assert got["cnt"] == cnt
        for key in ("vae_loss", "g_loss"):
            worst[key] = max(worst[key], abs(got[key] - ref[key]) / abs(ref[key]))
        worst["gan_loss"] = max(worst["gan_loss"], abs(got["gan_loss"] - ref["gan_loss"]) / (abs(ref["gan_loss"]) + 1e-6))
    e.eps_inject = None
    print("ten-step drift, worst relative loss error:", worst)
    assert worst["d_loss"] < 1e-3 and worst["vae_loss"] < 1e-3 and worst["g_loss"] < 1e-3 and worst["gan_loss"] < 2e-2, worst
    # accumulated displacement after ten updates: cosine with the oracle's displacement and ratio of the norms
    drift = {}
    for name, dev, ref_p, p0 in list(zip("W_q0 W_q1 W_p0 W_p1".split(), vae.params, gp, g0)) + \
            list(zip("w1 w2 w3 w4".split(), disc.d_params[0::2], dp[0::2], d0[0::2])):
        a = (dev.detach().cpu().double().reshape(p0.shape) - p0.double()).reshape(-1)
        b = (ref_p.double() - p0.double()).reshape(-1)
        drift[name] = (float(a @ b / (a.norm() * b.norm() + 1e-30)), float(a.norm() / (b.norm() + 1e-30)))
    print("ten-step drift, (cosine, norm ratio) of the accumulated displacement:", drift)
    for name, (cos, ratio) in drift.items():
        assert cos > 0.9 and 0.9 < ratio < 1.1, (name, cos, ratio)
