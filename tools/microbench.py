"""Kernel-level timings at the ML-20M-shaped configuration (B=500, I=20108). Development aid, not the bench."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("long-tail-gan_b200.ops")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ops.init()
    B, I, H = 500, 20108, 600
    ld = (I + 7) // 8 * 8
    dev = "cuda"
    h2 = (torch.randn(B, 608, device=dev) * 0.3).bfloat16()
    WdT = (torch.randn(I, H, device=dev) * 0.05).bfloat16()
    bd = torch.zeros(I, device=dev)
    logits = torch.zeros(B, ld, device=dev, dtype=torch.bfloat16)
    nblk = ops.dec_logits_nblk(B, I)
    partial = torch.zeros(nblk, B, 2, device=dev)
    t = timeit(lambda: ops.dec_logits_fwd(h2, WdT, bd, B, I, logits, partial))
    fl = 2.0 * B * H * I
    print("dec_logits_fwd        %8.1f us  %6.1f TFLOP/s" % (t, fl / t / 1e6))
    dl = (torch.randn(B, ld, device=dev) * 0.01).bfloat16()
    dh2 = torch.zeros(B, H, device=dev)
    for sp in (8, 16, 37):
        t = timeit(lambda: ops.gemm(dl, WdT, B, H, I, b_mn=True, splits=sp, bn=128, out_f32=dh2, atomic=True))
        print("dgrad splits=%-3d      %8.1f us  %6.1f TFLOP/s" % (sp, t, fl / t / 1e6))
    dW = torch.zeros(I, H, device=dev)
    db = torch.zeros(I, device=dev)
    for bn in (128, 256):
        t = timeit(lambda: ops.gemm(dl, h2, I, 601, B, a_mn=True, b_mn=True, lda=ld, ldb=608, bn=bn, out_f32=dW, ld_f32=H, aux_col=600, aux_out=db))
        print("wgrad bn=%-3d          %8.1f us  %6.1f TFLOP/s" % (bn, t, fl / t / 1e6))
    n = I * H
    p = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); g = torch.randn(n, device=dev)
    sh = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.adam(p, m, v, g, sh, lr_t=1e-4))
    print("adam [I,600]          %8.1f us  %6.1f GB/s (30 B/param)" % (t, 30.0 * n / t / 1e3))
    # encoder
    rng = np.random.RandomState(0)
    nnz_per = 73
    indptr = torch.arange(0, (B + 1) * nnz_per, nnz_per, dtype=torch.int32, device=dev)
    idx_np = np.concatenate([np.sort(rng.choice(I, nnz_per, replace=False)) for _ in range(B)]).astype(np.int32)
    indices = torch.from_numpy(idx_np).to(dev)
    Wenc = (torch.randn(I, H, device=dev) * 0.05).bfloat16()
    bq = torch.zeros(H, device=dev)
    h1 = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
    coef = torch.zeros(B * nnz_per, device=dev)
    t = timeit(lambda: ops.enc_gather_fwd(indptr, indices, None, B, I, 0, Wenc, bq, 0.75, 1, 0, None, h1, coef, nnz_per))
    print("enc_gather_fwd        %8.1f us  %6.1f GB/s" % (t, B * nnz_per * 0.75 * 1204 / t / 1e3))
    dlo = torch.zeros(B, ld, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, device=dev); xw = torch.ones(B, device=dev) * 73
    t = timeit(lambda: ops.dec_dlogits(logits, lse, xw, None, B, I, B, 0.0, None, indptr, indices, None, None, None, None, dlo))
    print("dec_dlogits           %8.1f us  %6.1f GB/s" % (t, 4.0 * B * ld / t / 1e3))
    # small GEMMs
    h1b = torch.randn(B, H, device=dev).bfloat16(); Wq1 = torch.randn(H, 400, device=dev).bfloat16(); mulv = torch.zeros(B, 400, device=dev)
    t = timeit(lambda: ops.gemm(h1b, Wq1, B, 400, H, b_mn=True, bn=64, out_f32=mulv))
    print("latent gemm bn=64     %8.1f us" % t)
    t = timeit(lambda: ops.gemm(h1b, Wq1, B, 400, H, b_mn=True, bn=128, out_f32=mulv))
    print("latent gemm bn=128    %8.1f us" % t)
    # top-k
    n_eval = 2000
    sc = torch.randn(n_eval, ld, device=dev).bfloat16()
    hp = torch.arange(0, (n_eval + 1) * 10, 10, dtype=torch.int32, device=dev)
    hi = torch.from_numpy(np.concatenate([np.sort(rng.choice(I, 10, replace=False)) for _ in range(n_eval)]).astype(np.int32)).to(dev)
    topk = torch.zeros(n_eval, 100, dtype=torch.int32, device=dev); dcg = torch.zeros(n_eval, dtype=torch.float64, device=dev)
    hits = torch.zeros(n_eval, 2, dtype=torch.int32, device=dev)
    t = timeit(lambda: ops.topk_metrics(sc, n_eval, I, hp, hi, hp, hi, 100, [20, 50], topk, dcg, hits), iters=5)
    print("topk_metrics 2000 rows%8.1f us  %6.1f GB/s (one pass = 2*I B/row)" % (t, n_eval * I * 2.0 / t / 1e3))


if __name__ == "__main__":
    main()
