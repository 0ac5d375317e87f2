"""Per-kernel table from an ncu report (`ncu -i rep --page raw --csv`): duration, DRAM bytes and GB/s, L2 throughput, tensor-pipe
activity, achieved occupancy, registers, IPC. Usage: python tools/ncu_table.py rep.ncu-rep [out.txt] [out.json]"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active": "umma_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "sm__inst_executed.avg.per_cycle_active": "ipc",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if not l.startswith("==")]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    hdr = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if "tensor" in h and "pct" in h]
    rows = []
    for r in rd:
        if len(r) < len(hdr):
            continue
        rec = dict(kernel=re.sub(r"\(.*$", "", r[col["Kernel Name"]])[:80])
        for m, k in WANT.items():
            if m in col:
                v = num(r[col[m]])
                if v is not None:
                    v *= UNIT.get(units[col[m]], 1.0)
                rec[k] = v
        tp = [num(r[col[h]]) for h in tensor_cols]
        rec["tensor_any_pct"] = max([x for x in tp if x is not None], default=None)
        rows.append(rec)
    txt = ["%-64s %8s %8s %8s %7s %7s %7s %6s %6s %5s %5s" % ("kernel", "us", "dramMB", "GB/s", "dram%", "l2%", "tens%", "occ%", "sm%", "ipc", "regs")]
    for r in rows:
        d = r.get("dur") or 0.0
        mb = ((r.get("dram_rd") or 0) + (r.get("dram_wr") or 0)) / 1e6
        gbs = mb / 1e3 / (d * 1e-6) if d > 0 else 0.0
        r["dram_MB"] = mb; r["dram_GBs"] = gbs
        f = lambda k: ("%.1f" % r[k]) if r.get(k) is not None else "-"   # noqa: E731
        txt.append("%-64s %8.1f %8.1f %8.0f %7s %7s %7s %6s %6s %5s %5s" % (r["kernel"][:64], d, mb, gbs, f("dram_pct"), f("l2_pct"), f("tensor_any_pct"),
                                                                         f("occ_pct"), f("sm_pct"), f("ipc"), f("regs")))
    text = "\n".join(txt)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    if len(sys.argv) > 3:
        json.dump(rows, open(sys.argv[3], "w"), indent=0)


if __name__ == "__main__":
    main()
