"""Run bench.py with the given extra args and print a one-line summary (development aid)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
for l in out.stdout.splitlines():
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.0f users/s  %.4f ms/step  phases %s  e2e %.0f  launches %s" % (d["value"], d["ms_per_step"], {k: round(v, 4) for k, v in d["phases_ms"].items() if k != "epoch_weighted_users_per_sec"}, d["e2e"]["value"], d["gpu_launches"]))
        break
else:
    print("NO JSON", out.stderr[-2000:])
