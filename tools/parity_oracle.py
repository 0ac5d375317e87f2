"""CPU side of the reference-schedule parity run: trains the oracle (oracle/train_oracle.py) on the bundled
dataset fixture with the reference schedule (Codes/config.ini: 80 epochs, NUM_SUB_EPOCHS = 10, lr 1e-4, batch 100) for ONE
(seed, GANLAMBDA) and writes the per-epoch validation history to profiles/r2_parity_oracle_s<seed>_lam<lam>.json after every
epoch. tools/parity_device.py runs the same configuration on the B200 and compares.

    python tools/parity_oracle.py <seed index> <GANLAMBDA> [epochs=80] [num_sub_epochs=10] [lr=1e-4] [threads=1]
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "askubuntu_sample.npz")


def main():
    s = int(sys.argv[1]); lam = float(sys.argv[2])
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    nsub = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    lr = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-4
    torch.set_num_threads(int(sys.argv[6]) if len(sys.argv) > 6 else 1)
    from oracle import ltgan_oracle as orc
    from oracle import train_oracle
    dp = importlib.import_module("long-tail-gan_b200.data_processing")
    g = np.load(GOLD)
    tabs = dp.tables_from_golden(g)
    vad = (g["vad_tr_indptr"], g["vad_tr_indices"].astype(np.int32), g["vad_te_indptr"], g["vad_te_indices"].astype(np.int32))
    cfg = dict(BATCH_SIZE=100, NUM_SUB_EPOCHS=nsub, LEARNING_RATE=lr, GANLAMBDA=lam)
    init = (orc.init_vae_params(1000, seed=98765 + s),) + orc.init_disc_params(1000, 100, 150, 250, 300, seed=77 + s)
    path = os.path.join(ROOT, "profiles", "r2_parity_oracle_s%d_lam%g.json" % (s, lam))
    out = dict(config=dict(epochs=epochs, num_sub_epochs=nsub, lr=lr, lam=lam, seed_index=s, init_seeds=[98765 + s, 77 + s], rng_seed=200 + s),
               history=[], seconds=[])
    t0 = time.time()

    def log(rec):
        out["history"].append(rec); out["seconds"].append(time.time() - t0)
        json.dump(out, open(path + ".tmp", "w")); os.replace(path + ".tmp", path)
        print(rec, flush=True)

    train_oracle.run_epochs(tabs, vad, cfg, init, epochs, seed=200 + s, log=log)


if __name__ == "__main__":
    main()
