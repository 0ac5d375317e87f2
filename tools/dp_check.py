"""2-GPU correctness check of the data-parallel G/D updates: ranks 0 and 1 each take half of a 100-user global batch;
rank 0 also runs the single-GPU engine on the whole batch. Encoder dropout, eps and sampling are keyed by the GLOBAL user
id, so with the discriminator dropout switched off (it is keyed by the local pair row) the two runs must produce the same
update up to bf16/atomic summation noise.   torchrun --nproc-per-node 2 tools/dp_check.py"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import helpers
    from oracle import ltgan_oracle as orc
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    I, B = 1000, 100
    tabs = helpers.synth_side_tables(np.random.RandomState(5), B, I)
    params = orc.init_vae_params(I, seed=98765); params[3] = params[3] * 3.0
    E, dparams = orc.init_disc_params(I, 100, 150, 250, 300, seed=77)

    def build(world_size, batch, first):
        vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=1); vae.set_params(params); vae.reset_optimizer()
        disc = dis.Discriminator(I, I, 100, 150, 250, 300, seed=1); disc.set_params(E, dparams)
        data = eng.TrainData(batch_size=batch, first_batch=first, max_batches=1, **tabs)
        e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=11, lr=1e-3, lam=1.0, keep_d=1.0, use_graphs=False, world_size=world_size,
                          B_global=B, max_active=data.max_active, rank=(rank if world_size > 1 else 0))
        if world_size > 1:
            e.attach_dp_tables(eng.build_dp_shard_tables(data, tabs["indptr"], tabs["indices"], world_size, rank, 1, e.R))
        return vae, disc, data, e

    vae, disc, data, e = build(world, B // world, rank)
    for _ in range(2):
        e.run_phase_a(data, 0); e.run_d_step(data, 0); e.run_g_step(data, 0)
    torch.cuda.synchronize()
    e.gather_master()
    # every rank's bf16 shadows (what its next forward reads) must be the rounding of the gathered fp32 masters
    ok_b = bool(torch.equal(vae.WdT_b, vae.WdT.bfloat16())) and bool(torch.equal(vae.W_q0_b, vae.W_q0.bfloat16()))
    print("rank %d: exchange path = %s, bf16 shadows consistent = %s" % (rank, ("peer memory, multicast %s" % ("on" if e.peer["dWdT_mc"] else "off")) if e.peer is not None else "NCCL collectives", ok_b),
          flush=True)
    assert ok_b
    dist.barrier()
    if rank == 0:
        vae1, disc1, data1, e1 = build(1, B, 0)
        for _ in range(2):
            e1.run_phase_a(data1, 0); e1.run_d_step(data1, 0); e1.run_g_step(data1, 0)
        torch.cuda.synchronize()

        def rel(a, b, a0):
            return float(((a - a0) - (b - a0)).norm() / ((b - a0).norm() + 1e-30))
        W0 = torch.as_tensor(params[3]).t().cuda(); Q0 = torch.as_tensor(params[0]).cuda()
        r1, r2 = rel(vae.WdT, vae1.WdT, W0), rel(vae.W_q0, vae1.W_q0, Q0)
        d0 = dis.Discriminator(I, I, 100, 150, 250, 300, seed=1); d0.set_params(E, dparams)
        r3 = rel(disc.arena, disc1.arena, d0.arena)
        print("DP(2) vs single-GPU displacement mismatch: W_dec %.4f  W_enc %.4f  disc %.4f" % (r1, r2, r3))
        assert r1 < 0.05 and r2 < 0.05 and r3 < 0.05
        print("dp_check ok")
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
