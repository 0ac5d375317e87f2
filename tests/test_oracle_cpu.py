"""CPU checks of the oracle's RNG mirrors (no GPU): Philox4x32-10 known answers and the statistical quality of the dropout
masks produced by the counter hash the GEMM epilogues use (oracle/philox.py mirrors ltg_common.cuh bit for bit; the bit-exact
device-vs-mirror comparison is tests/test_kernels_gpu.py::test_gemm_dropout_epilogue_bits)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox  # noqa: E402


def test_hash_quad_known_values_are_stable():
    # pins the definition (multiplier, round count, key schedule, word order): any change must be made on the device side too
    L, R = philox.hash_quad(0x12345678, np.array([0, 1, 2, (1 << 32) + 5], dtype=np.uint64))
    again = philox.hash_quad(0x12345678, np.array([0, 1, 2, (1 << 32) + 5], dtype=np.uint64))
    assert np.array_equal(L, again[0]) and np.array_equal(R, again[1])
    assert len(set(int(x) for x in L)) == 4 and len(set(int(x) for x in R)) == 4
    # one round by hand: L1 = hi(L0*M) ^ k ^ R0, R1 = lo(L0*M)
    M, k = 0xD256D193, 0x12345678
    l, r = 2, 0
    for _ in range(5):
        p = l * M
        l, r = ((p >> 32) ^ k ^ r) & 0xFFFFFFFF, p & 0xFFFFFFFF
        k = (k + 0x9E3779B9) & 0xFFFFFFFF
    assert int(L[2]) == l and int(R[2]) == r


def test_dropout_mask_statistics():
    """keep-rate and correlations of the [rows, cols] mask at the sampling-noise level (4 sigma), for the discriminator's pitches."""
    for ld, cols, keep in ((152, 150, 0.7), (256, 250, 0.7), (304, 300, 0.7), (64, 64, 0.5)):
        rows = 8192
        for step in (1, 2):
            m = philox.hash_keep_mask(20260101, philox.STREAM_DISC_DROPOUT, step, rows, cols, ld, keep).astype(np.float64)
            n = m.size
            p = int(np.float32(keep) * 65536.0) / 65536.0
            assert abs(m.mean() - p) < 4.0 * np.sqrt(p * (1 - p) / n)
            tol = 4.5 / np.sqrt(n)
            for a, b in ((m[:, :-1], m[:, 1:]), (m[:-1], m[1:]), (m[:, :-4], m[:, 4:]), (m[:-2], m[2:]), (m[:, :-2], m[:, 2:])):
                c = np.corrcoef(a.ravel(), b.ravel())[0, 1]
                assert abs(c) < tol, (ld, step, c, tol)
            # per-column and per-row keep rates scatter like binomials
            assert m.mean(0).std() < 1.5 * np.sqrt(p * (1 - p) / rows)
            assert m.mean(1).std() < 1.5 * np.sqrt(p * (1 - p) / cols)
        m2 = philox.hash_keep_mask(20260101, philox.STREAM_DISC_DROPOUT, 3, rows, cols, ld, keep).astype(np.float64)
        c = np.corrcoef(m.ravel(), m2.ravel())[0, 1]
        assert abs(c) < 4.5 / np.sqrt(m.size)   # masks of consecutive steps are independent


def test_philox4x32_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    out = philox.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0))
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox.philox4x32_10(np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff),
                               np.uint32(0xffffffff), np.uint32(0xffffffff))
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


# --------------------------------------------------------------------------------------------------------------------
# The TF-graph part of the oracle is "parity unpinned" by the reference (no TensorFlow here, no recorded outputs). These
# checks pin it against an independent float64 NumPy evaluation of SURVEY Appendix A (the reference's math written out
# from MultiVAE.py / discriminator.py / train.py line by line) and against the closed-form gradient of the F3 term.
# --------------------------------------------------------------------------------------------------------------------
import torch  # noqa: E402
from oracle import ltgan_oracle as orc  # noqa: E402


def _np(t):
    return t.detach().numpy().astype(np.float64)


def test_vae_forward_matches_float64_appendix_a():
    rng = np.random.RandomState(0)
    I, B = 57, 5
    params = orc.init_vae_params(I, seed=3)
    params[5] = params[5] + 0.05 * torch.randn(params[5].shape, generator=torch.Generator().manual_seed(1))   # non-trivial logvar bias
    X = (rng.rand(B, I) < 0.2).astype(np.float32); X[0, :3] = 1
    keep = rng.rand(B, I) < 0.75
    eps = rng.randn(B, orc.L).astype(np.float32)
    out = orc.vae_forward(params, torch.from_numpy(X), torch.from_numpy(keep), 0.75, torch.from_numpy(eps), 1, 0.13)
    W_q0, W_q1, W_p0, W_p1, b_q0, b_q1, b_p0, b_p1 = [_np(p) for p in params]
    x = X.astype(np.float64)
    h0 = x / np.sqrt(np.maximum((x * x).sum(1, keepdims=True), 1e-12))           # MultiVAE.py:148
    h0 = h0 * keep / 0.75                                                         # :149
    h1 = np.tanh(h0 @ W_q0 + b_q0)                                                # :151-155
    ml = h1 @ W_q1 + b_q1                                                         # :157-158
    mu, lv = ml[:, :orc.L], ml[:, orc.L:]
    KL = np.mean(np.sum(0.5 * (-lv + np.exp(lv) + mu ** 2 - 1), axis=1))          # :161-162
    z = mu + eps * np.exp(0.5 * lv)                                               # :160,178-181
    h2 = np.tanh(z @ W_p0 + b_p0)
    logits = h2 @ W_p1 + b_p1                                                     # :168-172
    lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1, keepdims=True)) + logits.max(1, keepdims=True)
    neg_ll = -np.mean(np.sum((logits - lse) * x, axis=1))                         # :108-112
    assert abs(out["KL"].item() - KL) < 1e-5 * max(1.0, abs(KL))
    assert abs(out["neg_ll"].item() - neg_ll) < 1e-5 * abs(neg_ll)
    assert abs(out["neg_ELBO"].item() - (neg_ll + 0.13 * KL)) < 1e-5 * abs(neg_ll)
    assert np.abs(_np(out["probs"]) - np.exp(logits - lse)).max() < 1e-6
    # inference mode (phase A / evaluation): z = mu, dropout still on (F4)
    out0 = orc.vae_forward(params, torch.from_numpy(X), torch.from_numpy(keep), 0.75, torch.from_numpy(eps), 0, 0.0)
    assert np.abs(_np(out0["z"]) - mu).max() < 1e-5


def test_disc_forward_and_d_loss_float64():
    rng = np.random.RandomState(1)
    I, P = 40, 9
    E, dp = orc.init_disc_params(I, 100, 150, 250, 300, seed=5)
    pop = rng.randint(0, I, P); niche = rng.randint(0, I, P)
    masks = [rng.rand(P, n) < 0.7 for n in (150, 250, 300)]
    y = orc.disc_forward(E, dp, torch.from_numpy(pop), torch.from_numpy(niche), [torch.from_numpy(m) for m in masks], 0.7)
    w1, b1, w2, b2, w3, b3, w4, b4 = [_np(p) for p in dp]
    En = _np(E)
    a1 = np.tanh(En[pop] @ w1 + b1) * masks[0] / 0.7                              # discriminator.py:25-30
    a2 = np.tanh(En[niche] @ w2 + b2) * masks[1] / 0.7
    a3 = np.tanh(np.concatenate([a1, a2], 1) @ w3 + b3) * masks[2] / 0.7          # :36-44
    want = 1.0 / (1.0 + np.exp(-(a3 @ w4 + b4).reshape(-1)))                      # :45
    assert np.abs(_np(y) - want).max() < 1e-6
    d = orc.d_loss_fn(y[:4], y[4:]).item()                                        # train.py:142
    assert abs(d - (-np.log(want[:4]).sum() - np.log(1 - want[4:]).sum())) < 1e-5 * abs(d)


def test_gan_term_broadcast_quirk_and_its_gradient():
    """F3: tf.multiply([K], [P,1]) broadcasts to [P,K]; the sum is (sum p)(sum y). Gradient w.r.t. the logits of user u:
    -(lam/cnt)(sum y) * pi_u o (m_u - sum_i m_ui pi_ui)   (SURVEY Appendix A)."""
    rng = np.random.RandomState(2)
    B, I, P = 3, 11, 5
    logits = torch.from_numpy(rng.randn(B, I)).requires_grad_(True)
    mask = torch.from_numpy((rng.rand(B, I) < 0.3).astype(np.float64))
    y = torch.from_numpy(rng.rand(P))
    probs = torch.softmax(logits, dim=-1)
    lam, cnt = 1.7, float(mask.sum().item())
    a = orc.gan_term(probs, mask, y, lam, cnt, literal_outer=True)
    b = orc.gan_term(probs, mask, y, lam, cnt, literal_outer=False)
    assert abs(a.item() - b.item()) < 1e-12
    assert abs(a.item() - (-(lam / cnt) * (probs * mask).sum().item() * y.sum().item())) < 1e-12
    (g,) = torch.autograd.grad(a, logits)
    pi = probs.detach().numpy(); m = mask.numpy()
    want = -(lam / cnt) * y.sum().item() * pi * (m - (m * pi).sum(1, keepdims=True))
    assert np.abs(g.numpy() - want).max() < 1e-12


def test_tf_adam_formula_and_shared_step_counter():
    """[ext] TF1 AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t), epsilon OUTSIDE the bias correction; one optimizer object serves the
    D and the G minimize() (train.py:160-164), so t advances on both (F6)."""
    lr = 1e-4
    for t in (1, 2, 7, 1000):
        assert abs(orc.tf_adam_lr_t(lr, t) - lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)) < 1e-9
    p = torch.tensor([0.5, -0.25]); m = torch.tensor([0.01, 0.0]); v = torch.tensor([1e-4, 0.0]); g = torch.tensor([0.2, 0.0])
    orc.tf_adam_step(p, m, v, g, orc.tf_adam_lr_t(lr, 3))
    m0 = 0.9 * 0.01 + 0.1 * 0.2
    v0 = 0.999 * 1e-4 + 0.001 * 0.04
    assert abs(m[0].item() - m0) < 1e-7 and abs(v[0].item() - v0) < 1e-9
    assert abs(p[0].item() - (0.5 - orc.tf_adam_lr_t(lr, 3) * m0 / (np.sqrt(v0) + 1e-8))) < 1e-7
    assert p[1].item() == -0.25                       # zero gradient, zero moments: the parameter does not move ...
    m2 = torch.tensor([0.01]); v2 = torch.tensor([1e-4]); p2 = torch.tensor([1.0])
    orc.tf_adam_step(p2, m2, v2, torch.tensor([0.0]), 1e-4)
    assert p2.item() < 1.0                            # ... but with non-zero momentum it does (dense Adam, F7)
    # first step from zero moments: displacement = lr * g / (|g| + eps*sqrt(1-b2)) ~ lr * sign(g)
    p3 = torch.tensor([0.0]); m3 = torch.zeros(1); v3 = torch.zeros(1)
    orc.tf_adam_step(p3, m3, v3, torch.tensor([3.0]), orc.tf_adam_lr_t(lr, 1))
    assert abs(p3.item() + lr) < 1e-8


def test_anneal_schedule():
    assert orc.anneal_value(0) == 0.0                                             # train.py:319-324: pre-increment count
    assert abs(orc.anneal_value(1000) - 0.05) < 1e-12
    assert orc.anneal_value(4000) == 0.2 and orc.anneal_value(10 ** 6) == 0.2
    assert orc.anneal_value(5, total_anneal_steps=0) == 0.2


def test_dae_forward_matches_float64_numpy():
    """oracle.dae_forward (MultiVAE.py:33-69: tanh on every layer but the last, NLL + lam * sum ||W||^2) against a float64 NumPy
    re-derivation; with lam = 0 the loss equals the NLL."""
    p = orc.init_dae_params(50, seed=3)
    assert [tuple(t.shape) for t in p] == [(50, 600), (600, 200), (200, 600), (600, 50), (600,), (200,), (600,), (50,)]
    X = (torch.rand(7, 50, generator=torch.Generator().manual_seed(1)) < 0.2).float()
    X[:, 0] = 1
    out = orc.dae_forward(p, X, None, 1.0, lam=0.01)
    W = [t.double().numpy() for t in p]
    x = X.double().numpy()
    h = x / np.sqrt((x * x).sum(1, keepdims=True))
    h = np.tanh(h @ W[0] + W[4]); h = np.tanh(h @ W[1] + W[5]); h = np.tanh(h @ W[2] + W[6])
    lg = h @ W[3] + W[7]
    mx = lg.max(1, keepdims=True)
    ls = lg - mx - np.log(np.exp(lg - mx).sum(1, keepdims=True))
    nll = -(ls * x).sum(1).mean()
    reg = 0.01 * sum((w * w).sum() for w in W[:4])
    assert abs(float(out["neg_ll"]) - nll) < 1e-4 * abs(nll) and abs(float(out["neg_ELBO"]) - (nll + reg)) < 1e-4 * (nll + reg)
    out0 = orc.dae_forward(p, X, None, 1.0)
    assert float(out0["neg_ELBO"]) == float(out0["neg_ll"]) and float(out0["KL"]) == 0.0
    assert np.allclose(out0["probs"].sum(1).numpy(), 1.0, atol=1e-5)
