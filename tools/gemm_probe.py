"""Fixed cost vs per-k-block cost of the tcgen05 GEMM on small problems (graph-replayed, so host launch cost is excluded)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("long-tail-gan_b200.ops")


def graph_time(fn, reps=20, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters) * 1e3


def main():
    ops.init()
    dev = "cuda"
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 64, 640, 64), (128, 64, 2560, 64), (500, 400, 64, 64), (500, 400, 600, 64), (500, 400, 600, 128),
                          (500, 600, 200, 64), (36000, 150, 101, 192), (36000, 250, 101, 256), (36000, 300, 408, 192), (36000, 408, 300, 256),
                          (500, 20108, 600, 256), (500, 20108, 600, 128)]:
        A = torch.randn(M, (K + 7) // 8 * 8, device=dev).bfloat16()
        B = torch.randn(N, (K + 7) // 8 * 8, device=dev).bfloat16()
        out = torch.zeros(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
        t = graph_time(lambda: ops.gemm(A, B, M, N, K, bn=bn, out_bf16=out))
        t2 = graph_time(lambda: ops.gemm(A, B, M, N, K, bn=bn, out_bf16=out, act=1, keep=0.7, seed=1, rng_stream=3, rng_ld=out.stride(0)))
        print("M=%6d N=%6d K=%5d bn=%3d : plain %7.2f us   tanh+dropout %7.2f us   (%.0f GFLOP -> %.0f TFLOP/s)" %
              (M, N, K, bn, t, t2, 2e-9 * M * N * K, 2e-6 * M * N * K / t))
    # decoder-forward shape: time against the catalog size at fixed batch 500 (fixed cost vs cost per column tile)
    for bn in (128, 192, 256):
        row = []
        for N in (bn, 148 * bn // 4, 148 * bn // 2, 148 * bn, 2 * 148 * bn):   # 1 tile, quarter / half / one / two tiles per SM column-wise
            M, K = 500, 600
            A = torch.randn(M, 608, device=dev).bfloat16()
            B = torch.randn(N, 600, device=dev).bfloat16()
            out = torch.zeros(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
            t = graph_time(lambda: ops.gemm(A, B, M, N, K, bn=bn, out_bf16=out))
            row.append("N=%6d %6.2f us" % (N, t))
        print("M=500 K=600 bn=%3d : " % bn + "   ".join(row))
    # empty kernel launch floor inside a graph
    x = torch.zeros(1024, device=dev)
    t = graph_time(lambda: x.add_(1.0))
    print("torch add_ (1 CTA) in graph: %.2f us" % t)


if __name__ == "__main__":
    main()
