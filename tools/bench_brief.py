"""Print the handful of numbers of a bench.py JSON line that a tuning session looks at."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path) if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:                      # noqa: BLE001
        print(path, "no JSON line:", e)
        continue
    r = d.get("roofline") or {}
    print("%s: value %.2f M  e2e %.2f M  reward %.2f M  ms/step %.3f  kernel_ms %.3f  frac %.3f  launches %s  clocks %s" % (
        d.get("config", {}).get("workload"), d["value"] / 1e6, (d.get("e2e") or {}).get("value", 0) / 1e6,
        (d.get("reward_only") or {}).get("value", 0) / 1e6, d["ms_per_step"], r.get("kernel_ms", 0), r.get("frac", 0),
        d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
    for k in ("parity_check", "other_workloads", "comm_us_per_step", "per_rank"):
        if k in d:
            print("   ", k, json.dumps(d[k]))
