"""Phase timestamps inside disc_fused_kernel (debug): runs the D-step launch on the bench shape with the trace buffer attached and
prints, for a few CTAs, the time of every stamp relative to the CTA's first stamp (us).  python tools/disc_trace.py [fwd]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

NAMES = {0: "P tile start", 1: "P regions free", 2: "P mma2 done->W3 ring", 3: "P W3 ring issued", 4: "P mma3 done->W3T ring",
         8: "M tile start", 9: "M ld0 landed", 10: "M ld1 landed", 11: "M hd ready", 12: "M mma3 issued", 13: "M dz3 ready", 14: "M mma4 issued",
         16: "E start", 17: "E mma1 done", 18: "E epi1 done", 19: "E mma2 done", 20: "E epi2 done/hd", 21: "E mma3 done", 22: "E fc1 done",
         23: "E head done", 24: "E dz3 done", 25: "E mma4 done", 26: "E epi4 done"}


def main():
    fwd_only = len(sys.argv) > 1 and sys.argv[1] == "fwd"
    import bench
    pkg = importlib.import_module("long-tail-gan_b200")
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    ops = importlib.import_module("long-tail-gan_b200.ops")
    N, I, deg = syn.CONFIGS["ml20m"]
    tabs = syn.make_config("ml20m", n_users=1000)
    data = eng.TrainData(batch_size=500, max_batches=2, **tabs)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765); vae.init_weights(98765)
    disc = dis.Discriminator(I, I, bench.H0, bench.H1, bench.H2, bench.H3, seed=4242)
    e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=2026, lr=bench.LR, lam=bench.LAM, max_active=data.max_active, use_graphs=False)
    for _ in range(2):
        e.run_step(data, 0)
    torch.cuda.synchronize()
    tr = torch.zeros(148 * 2 * 32, dtype=torch.int64, device="cuda")
    pkg._lib.load().ltg_disc_fused_set_trace(tr.data_ptr())
    bt = data.batches[0]
    d = e.disc
    gw4 = e.arena_gp[0][d._off["w4"][0]: d._off["w4"][0] + d._off["w4"][1]]; gb4 = e.arena_gp[0][d._off["b4"][0]: d._off["b4"][0] + d._off["b4"][1]]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if fwd_only:
        ops.disc_fwd_fused(e.Xp, e.Xn, bt["K"], d, bt["label"][bt["Pr"]:], e.keep_d, e.seed, 3, e.words, e.Hd, e.y, e.scal)
    else:
        ops.disc_fwd_fused(e.Xp, e.Xn, bt["P"], d, bt["label"], e.keep_d, e.seed, 3, e.words, e.Hd, e.y, e.scal, e.dz3, gw4, gb4, e.dz12)
    b.record()
    torch.cuda.synchronize()
    pkg._lib.load().ltg_disc_fused_set_trace(None)
    print("pairs", bt["K"] if fwd_only else bt["P"], "kernel %.1f us" % (a.elapsed_time(b) * 1e3))
    t = tr.cpu().numpy().reshape(148, 2, 32)
    t0 = t[:, 0, :][t[:, 0, :] > 0].min()
    for cta in (0, 1, 73, 147):
        for it in (0, 1):
            row = t[cta, it]
            if row.max() == 0:
                continue
            ev = sorted((int(v), k) for k, v in enumerate(row) if v > 0)
            print("CTA %d tile-iter %d" % (cta, it))
            for v, k in ev:
                print("   %8.2f us  %s" % ((v - t0) / 1e3, NAMES.get(k, str(k))))


if __name__ == "__main__":
    main()
