"""Generator plugin: `generator_VAECF(pro_dir)` with the reference's wrapper contract (Codes/generator.py:4-22,
README.md:74-79) and the device-resident MultiVAE it wraps (Codes/Base_Recommender/MultiVAE.py:95-230).

The returned 7-tuple is (model, item probability distribution, loss, params, p_dims, total_anneal_steps, anneal_cap).
There is no TensorFlow graph here: the "tensors" in slots [1] and [2] are lazy handles that are evaluated by the
engine (engine.GanEngine), which runs the CUDA kernels; `model` carries the reference's placeholder names
(input_ph, keep_prob_ph, is_training_ph, anneal_ph) as feed keys.

HBM layout (all row-major):
  W_q0   fp32 [I, 600]  + Adam m, v  + bf16 shadow (gathered by the encoder kernel, 1200 B per item row)
  W_p1^T fp32 [I, 600]  + Adam m, v  + bf16 shadow (K-major B operand of the decoder GEMM, MN-major operand of dgrad)
  small arena fp32: W_q1 [600,400] | W_p0 [200,600] | b_q0 | b_q1 | b_p0 | b_p1, one Adam launch, one bf16 shadow arena
The decoder weight is stored transposed so that forward, dgrad and wgrad all read the same tensor (see gemm_sm100.cuh).
"""
import os

import numpy as np
import torch

H = 600
L = 200


class Placeholder(object):
    """Stand-in for a tf.placeholder: only a named feed key."""

    def __init__(self, name, default=None):
        self.name = name
        self.default = default

    def __repr__(self):
        return "<placeholder %s>" % self.name


class LazyTensor(object):
    """Handle for a graph output (generator_out / neg_ELBO); evaluated by the engine."""

    def __init__(self, owner, name):
        self.owner = owner
        self.name = name

    def __repr__(self):
        return "<lazy %s of %r>" % (self.name, self.owner)


def _pad4(n):
    return (n + 3) // 4 * 4


class MultiVAE(object):
    is_dae = False      # MultiDAE below: deterministic encoder (no mu/logvar split, no KL)
    q1_out = 2 * L      # width of the second encoder layer: mu | logvar (MultiVAE.py:150,157-158)

    def __init__(self, p_dims, q_dims=None, lam=0.0, lr=1e-3, random_seed=None, device=None):
        # MultiVAE.py:12-31 (MultiDAE.__init__ / construct_placeholders) and 97-102
        assert len(p_dims) == 3 and p_dims[0] == L and p_dims[1] == H, "this build implements the VAE-CF 200-600-I generator"
        self.p_dims = list(p_dims)
        self.q_dims = self.p_dims[::-1] if q_dims is None else list(q_dims)
        assert self.q_dims[0] == self.p_dims[-1] and self.q_dims[-1] == self.p_dims[0]
        self.dims = self.q_dims + self.p_dims[1:]
        self.lam = lam
        assert lam == 0.0, "generator.py:18 builds MultiVAE with lam=0.0; the L2 term is dead code in the reference"
        self.lr = lr
        self.random_seed = random_seed
        self.n_items = int(p_dims[-1])
        self.device = torch.device("cuda" if device is None else device)
        self.input_ph = Placeholder("input_ph")
        self.keep_prob_ph = Placeholder("keep_prob_ph", 0.75)  # MultiVAE.py:31: default 0.75 even when scoring (F4)
        self.is_training_ph = Placeholder("is_training_ph", 0.0)
        self.anneal_ph = Placeholder("anneal_ph", 1.0)
        self._alloc()

    # ---- storage ----------------------------------------------------------------------------------------------
    def _alloc(self):
        I, dev = self.n_items, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.W_q0 = torch.zeros(I, H, **f32); self.W_q0_m = torch.zeros(I, H, **f32); self.W_q0_v = torch.zeros(I, H, **f32)
        self.WdT = torch.zeros(I, H, **f32); self.WdT_m = torch.zeros(I, H, **f32); self.WdT_v = torch.zeros(I, H, **f32)
        self.W_q0_b = torch.zeros(I, H, dtype=torch.bfloat16, device=dev)
        self.WdT_b = torch.zeros(I, H, dtype=torch.bfloat16, device=dev)
        sizes = [("W_q1", H * self.q1_out), ("W_p0", L * H), ("b_q0", H), ("b_q1", self.q1_out), ("b_p0", H), ("b_p1", _pad4(I))]
        self._small_off = {}
        off = 0
        for name, n in sizes:
            self._small_off[name] = (off, n)
            off += _pad4(n)
        self.small_n = off
        self.small = torch.zeros(off, **f32); self.small_m = torch.zeros(off, **f32); self.small_v = torch.zeros(off, **f32)
        self.small_g = torch.zeros(off, **f32)
        self.small_b = torch.zeros(off, dtype=torch.bfloat16, device=dev)

    def _view(self, arena, name, shape=None):
        off, n = self._small_off[name]
        t = arena[off:off + n]
        return t.view(*shape) if shape is not None else t

    # natural-shape views of the small parameters (fp32 master, gradient, bf16 shadow)
    def view(self, name, which="p"):
        arena = {"p": self.small, "m": self.small_m, "v": self.small_v, "g": self.small_g, "b": self.small_b}[which]
        shapes = {"W_q1": (H, self.q1_out), "W_p0": (L, H)}
        t = self._view(arena, name, shapes.get(name))
        if name == "b_p1":
            t = t[: self.n_items]
        return t

    # ---- parameters in the reference order (MultiVAE.py:129-141) ---------------------------------------------------
    @property
    def params(self):
        return [self.W_q0, self.view("W_q1"), self.view("W_p0"), self.WdT.t(), self.view("b_q0"), self.view("b_q1"), self.view("b_p0"),
                self.view("b_p1")]

    def set_params(self, params):
        """params = [W_q0 [I,600], W_q1 [600,400], W_p0 [200,600], W_p1 [600,I], b_q0, b_q1, b_p0, b_p1] (CPU or CUDA fp32)."""
        W_q0, W_q1, W_p0, W_p1, b_q0, b_q1, b_p0, b_p1 = [torch.as_tensor(p, dtype=torch.float32).to(self.device) for p in params]
        self.W_q0.copy_(W_q0)
        self.WdT.copy_(W_p1.t())
        self.view("W_q1").copy_(W_q1); self.view("W_p0").copy_(W_p0)
        self.view("b_q0").copy_(b_q0); self.view("b_q1").copy_(b_q1); self.view("b_p0").copy_(b_p0); self.view("b_p1").copy_(b_p1)
        self.refresh_shadows()

    def reset_optimizer(self):
        for t in (self.W_q0_m, self.W_q0_v, self.WdT_m, self.WdT_v, self.small_m, self.small_v):
            t.zero_()

    def refresh_shadows(self):
        self.W_q0_b.copy_(self.W_q0)
        self.WdT_b.copy_(self.WdT)
        self.small_b.copy_(self.small)

    def init_weights(self, seed=None):
        """MultiVAE._construct_weights (MultiVAE.py:188-230): Xavier-uniform weights, truncated-normal(0.001) biases.
        TensorFlow's seeded stream (seed 98765, generator.py:18) is not reproducible outside TF; this uses torch's."""
        seed = self.random_seed if seed is None else seed
        g = torch.Generator(device="cpu")
        g.manual_seed(0 if seed is None else int(seed))
        I = self.n_items

        def xavier(fi, fo):
            lim = float(np.sqrt(6.0 / (fi + fo)))
            return (torch.rand(fi, fo, generator=g) * 2 - 1) * lim

        def tn(n, std):
            x = torch.randn(2 * n + 16, generator=g)
            return x[x.abs() <= 2.0][:n] * std

        self.set_params([xavier(I, H), xavier(H, self.q1_out), xavier(L, H), xavier(H, I), tn(H, 0.001), tn(self.q1_out, 0.001),
                         tn(H, 0.001), tn(I, 0.001)])
        self.reset_optimizer()

    def build_graph(self):
        """MultiVAE.build_graph (MultiVAE.py:104-143): returns (softmax(logits), neg_ELBO, params)."""
        self.init_weights()
        self.generator_out = LazyTensor(self, "item_prob_dist")
        self.neg_ELBO = LazyTensor(self, "neg_ELBO")
        return self.generator_out, self.neg_ELBO, self.params

    def state_dict(self):
        return {k: getattr(self, k).detach().cpu() for k in
                ("W_q0", "W_q0_m", "W_q0_v", "WdT", "WdT_m", "WdT_v", "small", "small_m", "small_v")}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            getattr(self, k).copy_(v.to(self.device))
        self.refresh_shadows()

    def __repr__(self):
        return "%s(p_dims=%s)" % (type(self).__name__, self.p_dims)


class MultiDAE(MultiVAE):
    """The reference's other base recommender (Codes/Base_Recommender/MultiVAE.py:11-92): the same 4 dense layers without the
    variational middle -- h = l2_normalize(x) -> dropout -> tanh(h W0 + b0) -> tanh(. W1 + b1) -> tanh(. W2 + b2) -> . W3 + b3
    (MultiVAE.py:58-69: tanh on every layer but the last), loss = multinomial NLL (+ 2 * l2_regularizer(lam), MultiVAE.py:41-48).
    Storage, parameter order [W0, W1, W2, W3, b0, b1, b2, b3] and the engine's kernels are the MultiVAE's; the second encoder layer is
    200 wide instead of 400 and the latent head (mu/logvar, KL, reparameterisation) is replaced by a tanh epilogue. The L2 term is
    accepted only as lam = 0.0 -- what a generator.py-style wrapper passes (generator.py:18 builds its model with lam=0.0)."""
    is_dae = True
    q1_out = L


def count_items(pro_dir):
    """generator.py:6-11: n_items = number of lines of unique_item_id.txt."""
    n = 0
    with open(os.path.join(pro_dir, "unique_item_id.txt"), "r") as f:
        for _ in f:
            n += 1
    return n


def generator_DAECF(pro_dir):
    """A generator.py-style wrapper (README.md:74-79 contract) around MultiDAE: (model, item probability distribution, loss, params,
    p_dims, total_anneal_steps, anneal_cap); the last two are 0 (no KL term to anneal)."""
    n_items = count_items(pro_dir)
    p_dims = [200, 600, n_items]
    dae = MultiDAE(p_dims, lam=0.0, random_seed=98765)
    logits_var, loss_var, params = dae.build_graph()
    return dae, logits_var, loss_var, params, p_dims, 0, 0.0


def generator_VAECF(pro_dir):
    """Drop-in for Codes/generator.py:4-22. Returns
    (vae, item probability distribution handle, loss handle, params, p_dims, total_anneal_steps, anneal_cap)."""
    n_items = count_items(pro_dir)
    p_dims = [200, 600, n_items]  # VAECF recommended values (generator.py:13)
    total_anneal_steps = 20000    # generator.py:15
    anneal_cap = 0.2              # generator.py:16
    vae = MultiVAE(p_dims, lam=0.0, random_seed=98765)
    logits_var, loss_var, params = vae.build_graph()
    return vae, logits_var, loss_var, params, p_dims, total_anneal_steps, anneal_cap
