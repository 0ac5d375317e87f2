"""Kernel timeline of one captured GAN step (ML-20M shape): replays engine.run_step under torch.profiler (CUPTI activity records
carry device timestamps also for kernels launched from a CUDA graph) and prints, for the last replay, every kernel with its stream,
start offset and duration -- the tool that shows which chain is the critical path and what overlaps what.
    python tools/timeline.py [phase: step|a|d|g|eval] [out.json]
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    phase = sys.argv[1] if len(sys.argv) > 1 else "step"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "timeline_%s.json" % phase)
    import bench
    syn = importlib.import_module("long-tail-gan_b200.synthetic")
    gen = importlib.import_module("long-tail-gan_b200.generator")
    dis = importlib.import_module("long-tail-gan_b200.discriminator")
    eng = importlib.import_module("long-tail-gan_b200.engine")
    cfg = os.environ.get("LTG_TL_CONFIG", "ml20m")
    N, I, deg = syn.CONFIGS[cfg]
    B = int(os.environ.get("LTG_TL_BATCH", "500"))
    nb = 4
    tabs = syn.make_config(cfg, n_users=B * nb)
    data = eng.TrainData(batch_size=B, max_batches=nb, **tabs)
    vae = gen.MultiVAE([200, 600, I], lam=0.0, random_seed=98765); vae.init_weights(98765)
    disc = dis.Discriminator(I, I, bench.H0, bench.H1, bench.H2, bench.H3, seed=4242)
    e = eng.GanEngine(vae, disc, data.max_B, data.max_P, seed=2026, lr=bench.LR, lam=bench.LAM, max_active=data.max_active)
    if phase == "eval":
        # one engine.evaluate call over SURVEY 8d's 10 k held-out users: every kernel (and memcpy) of the call with its duration
        ev_args = syn.make_eval_split(int(os.environ.get("LTG_TL_EVAL_USERS", "10000")), I, deg)
        e.evaluate(*ev_args)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter(); e.evaluate(*ev_args); wall = time.perf_counter() - t0
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            e.evaluate(*ev_args)
            torch.cuda.synchronize()
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda ev: ev.time_range.start)
        t0 = evs[0].time_range.start
        print("evaluate: %d device activities, %.1f us from first to last, %.1f us host wall without the profiler" %
              (len(evs), max(ev.time_range.end for ev in evs) - t0, wall * 1e6))
        agg = {}
        for ev in evs:
            print("%8.1f %7.1f  %s" % (ev.time_range.start - t0, ev.time_range.end - ev.time_range.start, ev.name[:90]))
            a = agg.setdefault(ev.name[:60], [0, 0.0]); a[0] += 1; a[1] += ev.time_range.end - ev.time_range.start
        print("---- totals")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("%9.1f us  x%-3d %s" % (t, n, k))
        return
    fn = dict(step=e.run_step, a=e.run_phase_a, d=e.run_d_step, g=e.run_g_step)[phase]
    for r in range(3):
        for bi in range(nb):
            e.run_step(data, bi) if phase == "step" else (e.run_phase_a(data, bi), e.run_d_step(data, bi), e.run_g_step(data, bi))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for bi in range(nb):
            fn(data, bi)
            torch.cuda.synchronize()
    evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda ev: ev.time_range.start)
    # split into replays: gaps > 50 us
    groups, cur, last_end = [], [], None
    for ev in evs:
        st, en = ev.time_range.start, ev.time_range.end
        if last_end is not None and st - last_end > 50:
            groups.append(cur); cur = []
        cur.append(ev); last_end = max(last_end or en, en)
    groups.append(cur)
    g = groups[-1]
    t0 = g[0].time_range.start
    rows = []
    for ev in g:
        rows.append(dict(name=ev.name[:70], start_us=round(ev.time_range.start - t0, 1), dur_us=round(ev.time_range.end - ev.time_range.start, 1)))
    total = max(r["start_us"] + r["dur_us"] for r in rows)
    print("phase %s: %d kernels, %.1f us wall" % (phase, len(rows), total))
    for r in rows:
        print("%8.1f %7.1f  %s" % (r["start_us"], r["dur_us"], r["name"]))
    json.dump(dict(phase=phase, total_us=total, kernels=rows), open(out, "w"), indent=0)


if __name__ == "__main__":
    main()
