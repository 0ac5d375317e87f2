"""Discriminator plugin: `discriminator(n_items, FEATURE_LEN, h0, h1, h2, h3)` with the reference's 9-tuple contract
(Codes/discriminator.py:3-58) over device-resident parameters.

HBM layout: one fp32 arena (master + Adam m, v + gradient + bf16 shadow, identical layouts) so a D update is a single
fused Adam launch. Matrices are stored with pitches that are multiples of 8 elements (16 B in bf16) so TMA can read
them, padding is zero and receives zero gradient:
  W1 [h0+1, ld1]  W2 [h0+1, ld2]  W3 [k3, ld3]  w4 [ld3]  b4 [4]
Biases are the LAST ROW of their weight matrix ("ones-column" form): the gathered embedding rows carry a constant 1 in
column h0 and the hidden activation [h1 | pad | h2 | 1 | pad] carries a constant 1 in column `one3`, so x*W + b is a
single GEMM with K+1, and the bias gradient is simply one more row of the weight-gradient GEMM -- no bias vectors, no
column-sum reductions. k3 is the pitch of the hidden activation (branch 2 starts at a 16-byte aligned column), W3 has
matching zero rows. The item-embedding table E [n_items, h0] is a frozen random constant (it is not in d_params,
discriminator.py:14,47 / SURVEY F5): stored once as bf16 [n_items, 128] with column h0 = 1.
"""
import numpy as np
import torch

from .generator import LazyTensor, Placeholder


def _pad(n, q):
    return (n + q - 1) // q * q


class Discriminator(object):
    def __init__(self, n_items, FEATURE_LEN, h0_size, h1_size, h2_size, h3_size, device=None, seed=None):
        assert FEATURE_LEN == n_items or FEATURE_LEN is None
        assert h0_size < 128, "embedding rows are staged as 128-column bf16 rows (h0 values + the ones column)"
        self.n_items, self.h0, self.h1, self.h2, self.h3 = int(n_items), int(h0_size), int(h1_size), int(h2_size), int(h3_size)
        self.device = torch.device("cuda" if device is None else device)
        self.ld1 = _pad(self.h1, 8)
        self.ld2 = _pad(self.h2, 8)
        self.ld3 = _pad(self.h3, 8)
        self.off2 = _pad(self.h1, 8)                 # column where branch 2 starts inside the hidden activation
        self.one3 = self.off2 + self.h2              # column of the constant 1 inside the hidden activation
        self.k3 = _pad(self.one3 + 1, 8)             # pitch / K of the fc1 input
        segs = [("W1", (self.h0 + 1) * self.ld1), ("W2", (self.h0 + 1) * self.ld2), ("W3", self.k3 * self.ld3), ("w4", self.ld3), ("b4", 4)]
        self._off = {}
        off = 0
        for name, n in segs:
            self._off[name] = (off, n)
            off += _pad(n, 4)
        self.arena_n = off
        f32 = dict(dtype=torch.float32, device=self.device)
        self.arena = torch.zeros(off, **f32); self.arena_m = torch.zeros(off, **f32); self.arena_v = torch.zeros(off, **f32)
        self.arena_g = torch.zeros(off, **f32)
        self.arena_b = torch.zeros(off, dtype=torch.bfloat16, device=self.device)
        self.E = torch.zeros(self.n_items, self.h0, **f32)
        self.E_b = torch.zeros(self.n_items, 128, dtype=torch.bfloat16, device=self.device)
        # placeholders of the reference graph (discriminator.py:5-12)
        self.x_generated_id = Placeholder("x_generated")
        self.x_popular_n_id = Placeholder("x_popular_n")
        self.x_popular_g_id = Placeholder("x_popular_g")
        self.x_niche_id = Placeholder("x_niche")
        self.item_feature_arr = Placeholder("item_feature_arr")  # fed at train.py:300, consumed nowhere
        self.keep_prob = Placeholder("keep_prob")
        self.y_data = LazyTensor(self, "y_data")
        self.y_generated = LazyTensor(self, "y_generated")
        self.init_weights(seed)

    def view(self, name, which="p"):
        arena = {"p": self.arena, "m": self.arena_m, "v": self.arena_v, "g": self.arena_g, "b": self.arena_b}[which]
        off, n = self._off[name]
        t = arena[off:off + n]
        shapes = {"W1": (self.h0 + 1, self.ld1), "W2": (self.h0 + 1, self.ld2), "W3": (self.k3, self.ld3)}
        return t.view(*shapes[name]) if name in shapes else t

    # ---- parameters in the reference order d_params = [w1,b1,w2,b2,w3,b3,w4,b4] (discriminator.py:47) ----------------
    def _rows3(self):
        return torch.cat([torch.arange(0, self.h1), torch.arange(self.off2, self.off2 + self.h2)]).to(self.device)

    def get_params(self, which="p"):
        W1, W2, W3 = self.view("W1", which), self.view("W2", which), self.view("W3", which)
        h0 = self.h0
        return [W1[:h0, : self.h1].clone(), W1[h0, : self.h1].clone(), W2[:h0, : self.h2].clone(), W2[h0, : self.h2].clone(),
                W3[self._rows3()][:, : self.h3].clone(), W3[self.one3, : self.h3].clone(),
                self.view("w4", which)[: self.h3].clone().reshape(self.h3, 1), self.view("b4", which)[:1].clone()]

    @property
    def d_params(self):
        return self.get_params("p")

    def set_params(self, E, d_params):
        w1, b1, w2, b2, w3, b3, w4, b4 = [torch.as_tensor(p, dtype=torch.float32).to(self.device) for p in d_params]
        self.arena.zero_()
        h0 = self.h0
        W1, W2, W3 = self.view("W1"), self.view("W2"), self.view("W3")
        W1[:h0, : self.h1] = w1; W1[h0, : self.h1] = b1
        W2[:h0, : self.h2] = w2; W2[h0, : self.h2] = b2
        W3[: self.h1, : self.h3] = w3[: self.h1]
        W3[self.off2: self.off2 + self.h2, : self.h3] = w3[self.h1:]
        W3[self.one3, : self.h3] = b3
        self.view("w4")[: self.h3] = w4.reshape(-1)
        self.view("b4")[:1] = b4.reshape(-1)
        self.arena_b.copy_(self.arena)
        if E is not None:
            self.E.copy_(torch.as_tensor(E, dtype=torch.float32).to(self.device))
            self._refresh_E()
        self.arena_m.zero_(); self.arena_v.zero_()

    def _refresh_E(self):
        self.E_b.zero_()
        self.E_b[:, : self.h0] = self.E
        self.E_b[:, self.h0] = 1.0   # ones column: the bias rows of W1 / W2 ride the same GEMM

    def init_weights(self, seed=None):
        """discriminator.py:14-41: truncated_normal(stddev=0.1) matrices (unseeded in the reference), zero biases."""
        g = torch.Generator(device="cpu")
        g.manual_seed(int(np.random.SeedSequence().entropy % (2 ** 31)) if seed is None else int(seed))

        def tn(*shape):
            n = int(np.prod(shape))
            x = torch.randn(2 * n + 64, generator=g)
            return (x[x.abs() <= 2.0][:n] * 0.1).reshape(*shape)

        E = tn(self.n_items, self.h0)
        self.set_params(E, [tn(self.h0, self.h1), torch.zeros(self.h1), tn(self.h0, self.h2), torch.zeros(self.h2),
                            tn(self.h1 + self.h2, self.h3), torch.zeros(self.h3), tn(self.h3, 1), torch.zeros(1)])

    def state_dict(self):
        return {k: getattr(self, k).detach().cpu() for k in ("arena", "arena_m", "arena_v", "E")}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            getattr(self, k).copy_(v.to(self.device))
        self.arena_b.copy_(self.arena)
        self._refresh_E()

    def __repr__(self):
        return "Discriminator(h0=%d,h1=%d,h2=%d,h3=%d)" % (self.h0, self.h1, self.h2, self.h3)


def discriminator(n_items, FEATURE_LEN, h0_size, h1_size, h2_size, h3_size):
    """Drop-in for Codes/discriminator.py:3-58. Returns
    (y_data, y_generated, d_params, x_generated_id, x_popular_n_id, x_popular_g_id, x_niche_id, item_feature_arr, keep_prob);
    the network object itself is reachable as `y_data.owner`."""
    net = Discriminator(n_items, FEATURE_LEN, h0_size, h1_size, h2_size, h3_size)
    return (net.y_data, net.y_generated, net.d_params, net.x_generated_id, net.x_popular_n_id, net.x_popular_g_id, net.x_niche_id,
            net.item_feature_arr, net.keep_prob)
