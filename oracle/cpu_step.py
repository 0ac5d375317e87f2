"""TEST/BENCH INFRASTRUCTURE ONLY -- the reference's GAN step on the host CPU, in the reference's own data flow.

TensorFlow cannot be installed in this image (no network), so the `cpu_baseline` and `--impl reference`
legs of bench.py time this restatement: dense [B, I] fp32 feeds built from the CSR rows (train.py:194-198), a dense
softmax forward for phase A (train.py:200), the verbatim host sampling / pairing loop (train.py:212-251 via
ltgan_oracle.build_pairs_for_batch -> sample.py:40-67), one discriminator update (train.py:300) and one generator update
(train.py:326) with dense per-variable TF-Adam, all in PyTorch-CPU fp32 with every host thread torch is given.
"""
import time

import numpy as np
import torch

from . import ltgan_oracle as orc


class CpuGanStep(object):
    def __init__(self, n_items, h0=100, h1=150, h2=250, h3=300, lr=1e-4, lam=1.0, seed=98765):
        self.I = n_items
        self.params = orc.init_vae_params(n_items, seed)
        self.gm = [torch.zeros_like(p) for p in self.params]
        self.gv = [torch.zeros_like(p) for p in self.params]
        self.E, self.dparams = orc.init_disc_params(n_items, h0, h1, h2, h3, seed + 1)
        self.dm = [torch.zeros_like(p) for p in self.dparams]
        self.dv = [torch.zeros_like(p) for p in self.dparams]
        self.h = (h1, h2, h3)
        self.lr, self.lam = lr, lam
        self.t = 0
        self.update_count = 0
        self.rng = np.random.RandomState(seed)

    def _dense(self, tabs, b0, b1):
        ip, idx = tabs["indptr"], tabs["indices"]
        X = np.zeros((b1 - b0, self.I), dtype=np.float32)
        rows = np.repeat(np.arange(b1 - b0), np.diff(ip[b0:b1 + 1]))
        X[rows, idx[ip[b0]:ip[b1]]] = 1.0
        return torch.from_numpy(X)

    def _dicts(self, tabs, b0, b1):
        up, un, rn, rp, cand = {}, {}, {}, {}, {}
        idx, ip = tabs["indices"], tabs["indptr"]
        for u in range(b0, b1):
            if not tabs["eligible"][u]:
                continue
            pops = tabs["pop_items"][tabs["pop_ptr"][u]:tabs["pop_ptr"][u + 1]]
            up[u] = pops.tolist()
            un[u] = [0] * int(tabs["n_niche"][u])
            rn[u] = tabs["real_niche"][tabs["real_ptr"][u]:tabs["real_ptr"][u + 1]].tolist()
            rp[u] = tabs["real_pop"][tabs["real_ptr"][u]:tabs["real_ptr"][u + 1]].tolist()
            cand[u] = tabs["cand_items"][tabs["cand_ptr"][u]:tabs["cand_ptr"][u + 1]]
        return up, un, rn, rp, cand

    def _masks(self, n):
        return [torch.from_numpy(self.rng.rand(n, w) < 0.7) for w in self.h]

    def step(self, tabs, b0, b1, valid_set=None):
        """One A + D + G pass over users [b0, b1). Returns dict of wall-clock seconds per phase and the losses."""
        B = b1 - b0
        t0 = time.perf_counter()
        X = self._dense(tabs, b0, b1)
        keep = torch.from_numpy(self.rng.rand(B, self.I) < 0.75)
        with torch.no_grad():
            probs = orc.vae_forward(self.params, X, keep, 0.75, None, 0.0, 0.0)["probs"].numpy()  # train.py:200
        up, un, rn, rp, cand = self._dicts(tabs, b0, b1)
        valid = valid_set if valid_set is not None else set(np.nonzero(tabs["item_valid"])[0].tolist())
        pairs = orc.build_pairs_for_batch(list(range(b0, b1)), probs, up, un, rn, rp, cand, valid, self.I, self.rng)
        t1 = time.perf_counter()
        tp = {k: torch.from_numpy(pairs[k]) for k in ("x_popular_n", "x_niche", "x_popular_g", "x_generated")}
        out = dict(cnt=pairs["cnt"])
        if pairs["cnt"] > 0:  # train.py:254-255
            self.t += 1
            d_loss, _ = orc.d_step(self.E, self.dparams, self.dm, self.dv, tp, self._masks(len(pairs["x_niche"])),
                                   self._masks(pairs["cnt"]), 0.7, orc.tf_adam_lr_t(self.lr, self.t))
            t2 = time.perf_counter()
            self.t += 1
            anneal = orc.anneal_value(self.update_count)
            self.update_count += 1
            keep = torch.from_numpy(self.rng.rand(B, self.I) < 0.75)
            eps = torch.from_numpy(self.rng.randn(B, orc.L).astype(np.float32))
            mask = torch.from_numpy(pairs["mask"].astype(np.float32))
            g = orc.g_step(self.params, self.gm, self.gv, self.E, self.dparams, X, keep, 0.75, eps, anneal, mask, tp, self._masks(pairs["cnt"]),
                           0.7, self.lam, pairs["cnt"], orc.tf_adam_lr_t(self.lr, self.t), literal_outer=False)
            out.update(d_loss=d_loss, g_loss=g["g_loss"], vae_loss=g["vae_loss"], gan_loss=g["gan_loss"])
        else:
            t2 = t1
        t3 = time.perf_counter()
        out.update(t_a=t1 - t0, t_d=t2 - t1, t_g=t3 - t2, t=t3 - t0)
        return out
