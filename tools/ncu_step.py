"""One eager (no CUDA graphs, no stream overlap) GAN step at the bench shape, bracketed by cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_step python tools/ncu_step.py
Also runs one evaluation batch (fold-in forward + top-k metrics) inside the profiled region when LTG_NCU_EVAL=1."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import bench
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    cfg = os.environ.get("LTG_TL_CONFIG", "ml20m")
    N, I, deg = syn.CONFIGS[cfg]
    B = int(os.environ.get("LTG_TL_BATCH", "500"))
    nb = 2
    tabs = syn.make_config(cfg, n_users=B * nb)
    data = eng.TrainData(batch_size=B, max_batches=nb, **tabs)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765); vae.init_weights(98765)
    disc = dis.Discriminator(I, I, bench.H0, bench.H1, bench.H2, bench.H3, seed=4242)
    e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=2026, lr=bench.LR, lam=bench.LAM, max_active=data.max_active, use_graphs=False)
    e.overlap = False
    for bi in range(nb):
        e.run_step(data, bi)
    ev = None
    if os.environ.get("LTG_NCU_EVAL", "0") == "1":
        ip = np.asarray(tabs["indptr"], dtype=np.int64)[: B + 1]
        idx = np.asarray(tabs["indices"], dtype=np.int32)[: ip[-1]]
        held = np.zeros(len(idx), dtype=bool); held[4::5] = True
        row = np.repeat(np.arange(B), np.diff(ip))
        trp = np.concatenate([[0], np.cumsum(np.bincount(row[~held], minlength=B))]); tep = np.concatenate([[0], np.cumsum(np.bincount(row[held], minlength=B))])
        ev = (trp, idx[~held], tep, idx[held])
        e.evaluate(*ev)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e.run_step(data, 0)
    if ev is not None:
        e.evaluate(*ev)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
