"""stdin: bench.py JSON line -> one short line with ms per step and per kernel (experiment sweeps of tools/gpu_round.sh)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.get("roofline", {}).get("kernels", {})
    print("%s | step %.3f ms | e2e %.2f ms | %s" % (
        d["config"]["workload"][:28], d["ms_per_step"], d.get("e2e", {}).get("ms_per_step", 0.0),
        " ".join("%s=%.3f" % (n.replace("k_decode_", "").replace("k_", ""), v["ms_per_step"]) for n, v in k.items())))
