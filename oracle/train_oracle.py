"""TEST INFRASTRUCTURE ONLY -- the reference's epoch loop (Codes/train.py:180-348) on the CPU oracle, for end-to-end
NDCG/Recall parity runs against the device engine: phase A over every batch with the epoch-start generator, then
NUM_SUB_EPOCHS passes of D updates and of G updates over one shuffled order of the batches that produced pairs, then the
validation pass with dropout left on (MultiVAE.py:31) and the metric code of eval_functions.py."""
import numpy as np
import torch
from scipy import sparse

from . import ltgan_oracle as orc


def _dense(indptr, indices, b0, b1, I):
    X = np.zeros((b1 - b0, I), dtype=np.float32)
    rows = np.repeat(np.arange(b1 - b0), np.diff(indptr[b0:b1 + 1]))
    X[rows, indices[indptr[b0]:indptr[b1]]] = 1.0
    return torch.from_numpy(X)


def run_epochs(tabs, vad, cfg, init, n_epochs, seed=0, log=None):
    """tabs: dict like data_processing.build_train_tables; vad = (tr_indptr, tr_indices, te_indptr, te_indices);
    cfg: BATCH_SIZE, NUM_SUB_EPOCHS, LEARNING_RATE, GANLAMBDA; init = (vae params list, E, d_params list)."""
    I = int(tabs["n_items"])
    rng = np.random.RandomState(seed)
    params = [p.clone() for p in init[0]]
    E, dparams = init[1].clone(), [p.clone() for p in init[2]]
    gm = [torch.zeros_like(p) for p in params]; gv = [torch.zeros_like(p) for p in params]
    dm = [torch.zeros_like(p) for p in dparams]; dv = [torch.zeros_like(p) for p in dparams]
    hs = (dparams[0].shape[1], dparams[2].shape[1], dparams[4].shape[1])
    N = len(tabs["indptr"]) - 1
    B = cfg["BATCH_SIZE"]
    lr, lam = cfg["LEARNING_RATE"], cfg["GANLAMBDA"]
    valid = set(np.nonzero(tabs["item_valid"])[0].tolist())
    t = 0
    update_count = 0
    history = []
    masks = lambda n: [torch.from_numpy(rng.rand(n, w) < 0.7) for w in hs]  # noqa: E731
    te = sparse.csr_matrix((np.ones(len(vad[3])), vad[3].astype(np.int64), vad[2].astype(np.int64)), shape=(len(vad[2]) - 1, I))
    for ep in range(n_epochs):
        cache = []
        for b0 in range(0, N, B):                                      # train.py:192-269
            b1 = min(N, b0 + B)
            X = _dense(tabs["indptr"], tabs["indices"], b0, b1, I)
            keep = torch.from_numpy(rng.rand(b1 - b0, I) < 0.75)
            with torch.no_grad():
                probs = orc.vae_forward(params, X, keep, 0.75, None, 0.0, 0.0)["probs"].numpy()
            up, un, rn, rp, cand = {}, {}, {}, {}, {}
            for u in range(b0, b1):
                if not tabs["eligible"][u]:
                    continue
                up[u] = tabs["pop_items"][tabs["pop_ptr"][u]:tabs["pop_ptr"][u + 1]].tolist()
                un[u] = [0] * int(tabs["n_niche"][u])
                rn[u] = tabs["real_niche"][tabs["real_ptr"][u]:tabs["real_ptr"][u + 1]].tolist()
                rp[u] = tabs["real_pop"][tabs["real_ptr"][u]:tabs["real_ptr"][u + 1]].tolist()
                cand[u] = tabs["cand_items"][tabs["cand_ptr"][u]:tabs["cand_ptr"][u + 1]]
            pairs = orc.build_pairs_for_batch(list(range(b0, b1)), probs, up, un, rn, rp, cand, valid, I, rng)
            if pairs["cnt"] == 0:                                      # train.py:254-255
                continue
            cache.append((X, pairs))
        order = np.arange(len(cache))
        rng.shuffle(order)                                             # train.py:284-285
        d_loss = float("nan")
        for _ in range(cfg["NUM_SUB_EPOCHS"]):                         # train.py:287-303
            for bi in order:
                X, pairs = cache[bi]
                tp = {k: torch.from_numpy(pairs[k]) for k in ("x_popular_n", "x_niche", "x_popular_g", "x_generated")}
                t += 1
                d_loss, _ = orc.d_step(E, dparams, dm, dv, tp, masks(len(pairs["x_niche"])), masks(pairs["cnt"]), 0.7, orc.tf_adam_lr_t(lr, t))
        g = None
        diag = dict(sp=[], ybar=[], vae=[], gan=[])                    # last G sub-epoch: what the adversarial term acts on
        for j_gen in range(cfg["NUM_SUB_EPOCHS"]):                     # train.py:307-329
            for bi in order:
                X, pairs = cache[bi]
                tp = {k: torch.from_numpy(pairs[k]) for k in ("x_popular_n", "x_niche", "x_popular_g", "x_generated")}
                t += 1
                anneal = orc.anneal_value(update_count)
                update_count += 1
                keep = torch.from_numpy(rng.rand(X.shape[0], I) < 0.75)
                eps = torch.from_numpy(rng.randn(X.shape[0], orc.L).astype(np.float32))
                g = orc.g_step(params, gm, gv, E, dparams, X, keep, 0.75, eps, anneal, torch.from_numpy(pairs["mask"].astype(np.float32)), tp,
                               masks(pairs["cnt"]), 0.7, lam, pairs["cnt"], orc.tf_adam_lr_t(lr, t), literal_outer=False)
                if j_gen == cfg["NUM_SUB_EPOCHS"] - 1:
                    diag["sp"].append(float((g["probs"] * torch.from_numpy(pairs["mask"].astype(np.float32))).sum()) / pairs["cnt"])
                    diag["ybar"].append(float(g["y_gen"].mean())); diag["vae"].append(g["vae_loss"]); diag["gan"].append(g["gan_loss"])
        # validation, dropout still on (train.py:333-348)
        Nv = len(vad[0]) - 1
        Xv = _dense(vad[0], vad[1], 0, Nv, I)
        keep = torch.from_numpy(rng.rand(Nv, I) < 0.75)
        with torch.no_grad():
            pred = orc.vae_forward(params, Xv, keep, 0.75, None, 0.0, 0.0)["probs"].numpy()
        pred[Xv.numpy().nonzero()] = -np.inf
        rec = dict(epoch=ep, ndcg=float(np.mean(orc.ndcg_binary_at_k_batch(pred, te, 100))),
                   r20=float(np.mean(orc.recall_at_k_batch(pred, te, 20)[0])), r50=float(np.mean(orc.recall_at_k_batch(pred, te, 50)[0])),
                   d_loss=float(d_loss), g_loss=None if g is None else g["g_loss"],
                   sp_mean=float(np.mean(diag["sp"])) if diag["sp"] else None, ybar_mean=float(np.mean(diag["ybar"])) if diag["ybar"] else None,
                   vae_loss_mean=float(np.mean(diag["vae"])) if diag["vae"] else None, gan_loss_mean=float(np.mean(diag["gan"])) if diag["gan"] else None)
        history.append(rec)
        if log:
            log(rec)
    return history
