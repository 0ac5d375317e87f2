"""Device side of the reference-schedule parity run. For every oracle history committed by tools/parity_oracle.py
(profiles/r2_parity_oracle_s<seed>_lam<lam>.json: Codes/config.ini schedule -- NUM_SUB_EPOCHS = 10, lr 1e-4, batch 100 -- on the
bundled dataset fixture) trains the device engine (train.train_GAN) from the SAME initial weights for the same number of epochs and
writes profiles/r2_parity_train.json: per-epoch validation NDCG@100 / Recall@20 / Recall@50 of both sides, final means +- std over the
seeds, the deltas (bar: 0.005 absolute, north_star), and the GANLAMBDA = 0 control that shows the comparison sees the adversarial
term (mean sampled probability of the last G sub-epoch, which the GAN gradient pushes up directly).
    python tools/parity_device.py [max_epochs]"""
import glob
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "askubuntu_sample.npz")


def main():
    cap = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 9
    from oracle import ltgan_oracle as orc
    train = importlib.import_module("long-tail-gan_b200.train")
    runs = []
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_parity_oracle_s*_lam*.json"))):
        o = json.load(open(path))
        c = o["config"]
        n_ep = min(len(o["history"]), cap)
        if n_ep == 0:
            continue
        s = c["seed_index"]
        init = (orc.init_vae_params(1000, seed=c["init_seeds"][0]),) + orc.init_disc_params(1000, 100, 150, 250, 300, seed=c["init_seeds"][1])
        cfg = dict(h0_size=100, h1_size=150, h2_size=250, h3_size=300, NUM_EPOCH=8 * c["num_sub_epochs"], NUM_SUB_EPOCHS=c["num_sub_epochs"],
                   BATCH_SIZE=100, DISPLAY_ITER=50, LEARNING_RATE=c["lr"], to_restore=0, model_name="LT_GAN", dataset=GOLD, GANLAMBDA=c["lam"])
        cfg["NUM_EPOCH"] = max(cfg["NUM_EPOCH"], n_ep)
        t0 = time.time()
        dev = train.train_GAN(max_epochs=n_ep, quiet=True, save=False, seed=100 + s, init=init, diag=True, **cfg)["history"]
        runs.append(dict(seed_index=s, lam=c["lam"], epochs=n_ep, device=dev, oracle=o["history"][:n_ep], device_seconds=time.time() - t0,
                         oracle_seconds=o["seconds"][n_ep - 1], oracle_file=os.path.basename(path)))
        print("seed %d lam %g: %d epochs, device %.0fs (oracle %.0fs)  final ndcg dev %.4f / ora %.4f   sp dev %.4f / ora %.4f" %
              (s, c["lam"], n_ep, time.time() - t0, o["seconds"][n_ep - 1], dev[-1]["ndcg"], o["history"][n_ep - 1]["ndcg"],
               dev[-1].get("sp_mean") or 0, o["history"][n_ep - 1].get("sp_mean") or 0), flush=True)
    out = dict(schedule="Codes/config.ini: NUM_SUB_EPOCHS 10, lr 1e-4, batch 100, bundled Askubuntu sample", runs=runs)

    def stat(lam, side, key, last=3):
        # mean over the last `last` epochs of each run (epoch-to-epoch noise of the dropout-on validation), then over seeds
        v = [float(np.mean([h[key] for h in r[side][-last:]])) for r in runs if r["lam"] == lam and r[side][-1].get(key) is not None]
        return dict(mean=float(np.mean(v)), std=float(np.std(v)), n=len(v)) if v else None
    summ = {}
    for lam in sorted({r["lam"] for r in runs}):
        d = {}
        for key in ("ndcg", "r20", "r50", "sp_mean", "ybar_mean", "vae_loss_mean", "gan_loss_mean"):
            a, b = stat(lam, "device", key), stat(lam, "oracle", key)
            if a and b:
                d[key] = dict(device=a, oracle=b, delta=a["mean"] - b["mean"])
        summ["lam_%g" % lam] = d
    out["summary_last3_epochs"] = summ
    if "lam_1" in summ:
        out["bar"] = dict(abs_delta_ndcg100=abs(summ["lam_1"]["ndcg"]["delta"]), abs_delta_recall50=abs(summ["lam_1"]["r50"]["delta"]), limit=0.005,
                          ok=bool(abs(summ["lam_1"]["ndcg"]["delta"]) <= 0.005 and abs(summ["lam_1"]["r50"]["delta"]) <= 0.005))
    if "lam_1" in summ and "lam_0" in summ and "sp_mean" in summ["lam_1"] and "sp_mean" in summ["lam_0"]:
        out["adversarial_term_visible"] = dict(
            sp_mean_ratio_lam1_over_lam0=dict(device=summ["lam_1"]["sp_mean"]["device"]["mean"] / summ["lam_0"]["sp_mean"]["device"]["mean"],
                                              oracle=summ["lam_1"]["sp_mean"]["oracle"]["mean"] / summ["lam_0"]["sp_mean"]["oracle"]["mean"]),
            note="mean softmax probability of the sampled niche items over the last G sub-epoch: the quantity the F3 term pushes up")
    print(json.dumps({k: out[k] for k in out if k != "runs"}, indent=1))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_parity_train.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
