"""Top SASS lines by warp-stall samples from `ncu -i rep --page source --csv --kernel-name regex:X [--launch-skip n --launch-count 1]`."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    # names are matched against the demangled name WITH template arguments (bool arguments print as 0 / 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "--kernel-name", "regex:" + kern,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    print(lines[0][:200])
    rd = [r for r in csv.DictReader(lines[1:]) if (r.get("# Samples") or "0").replace(".", "").isdigit()]
    tot = sum(float(r["# Samples"] or 0) for r in rd)
    stall_keys = [k for k in rd[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(float(r[k] or 0) for r in rd) for k in stall_keys}
    print("total samples", tot, "instructions", len(rd))
    print("stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = sorted(range(len(rd)), key=lambda i: -float(rd[i]["# Samples"] or 0))[:top]
    for i in sorted(idx):
        r = rd[i]
        st = sorted(((k[6:], float(r[k] or 0)) for k in stall_keys), key=lambda kv: -kv[1])[:2]
        print("%5d %6.1f%%  %-70s %s" % (i, 100 * float(r["# Samples"] or 0) / max(tot, 1), r["Source"].strip()[:70],
                                        " ".join("%s:%d" % (a, b) for a, b in st if b > 0)))


if __name__ == "__main__":
    main()
