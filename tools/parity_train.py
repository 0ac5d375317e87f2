"""End-to-end parity run on the bundled-dataset fixture: trains the device engine (train.train_GAN) and the CPU oracle
(oracle/train_oracle.py) from the SAME initial weights with the reference schedule, and prints validation NDCG@100 /
Recall@20 / Recall@50 per epoch for both. Randomness (dropout, eps, sampling, batch order) is independent on the two
sides, so the comparison is statistical: north_star asks for |delta NDCG@100|, |delta Recall@50| <= 0.005.

    python tools/parity_train.py [epochs] [num_sub_epochs] [lr] [seeds]
"""
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
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
    seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    from oracle import ltgan_oracle as orc
    from oracle import train_oracle
    train = importlib.import_module("long-tail-gan_b200.train")
    dp = importlib.import_module("long-tail-gan_b200.data_processing")
    g = np.load(GOLD)
    tabs = dp.tables_from_golden(g)
    vad = (g["vad_tr_indptr"], g["vad_tr_indices"].astype(np.int32), g["vad_te_indptr"], g["vad_te_indices"].astype(np.int32))
    cfg = dict(h0_size=100, h1_size=150, h2_size=250, h3_size=300, NUM_EPOCH=8 * nsub, NUM_SUB_EPOCHS=nsub, BATCH_SIZE=100,
               DISPLAY_ITER=50, LEARNING_RATE=lr, to_restore=0, model_name="LT_GAN", dataset=GOLD, GANLAMBDA=1.0)
    out = dict(config=dict(epochs=epochs, num_sub_epochs=nsub, lr=lr), device=[], oracle=[])
    # The oracle side depends only on (init, seed, config): PARITY_ORACLE_JSON=<earlier output of this tool with the same arguments>
    # reuses its oracle histories, so a re-check of the device side after a kernel change does not spend minutes of CPU training.
    cached = None
    if os.environ.get("PARITY_ORACLE_JSON"):
        cached = json.load(open(os.environ["PARITY_ORACLE_JSON"]))
        assert cached["config"] == out["config"] and len(cached["oracle"]) >= seeds, "cached oracle run has different arguments"
        out["oracle_source"] = os.environ["PARITY_ORACLE_JSON"]
    for s in range(seeds):
        init = (orc.init_vae_params(1000, seed=98765 + s),) + orc.init_disc_params(1000, 100, 150, 250, 300, seed=77 + s)
        t0 = time.time()
        dev = train.train_GAN(max_epochs=epochs, quiet=True, save=False, seed=100 + s, init=init, **cfg)["history"]
        t1 = time.time()
        ora = cached["oracle"][s] if cached is not None else train_oracle.run_epochs(tabs, vad, cfg, init, epochs, seed=200 + s)
        t2 = time.time()
        out["device"].append(dev); out["oracle"].append(ora)
        print("seed %d: device %.1fs, oracle %.1fs" % (s, t1 - t0, t2 - t1))
        for e in range(epochs):
            print("  epoch %d  device ndcg %.4f r20 %.4f r50 %.4f | oracle ndcg %.4f r20 %.4f r50 %.4f" %
                  (e, dev[e]["ndcg"], dev[e]["r20"], dev[e]["r50"], ora[e]["ndcg"], ora[e]["r20"], ora[e]["r50"]))
    fin = lambda side, k: float(np.mean([h[-1][k] for h in out[side]]))  # noqa: E731
    out["final_mean"] = {k: dict(device=fin("device", k), oracle=fin("oracle", k), delta=fin("device", k) - fin("oracle", k)) for k in ("ndcg", "r20", "r50")}
    print(json.dumps(out["final_mean"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_train.json"), "w"))


if __name__ == "__main__":
    main()
