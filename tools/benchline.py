#!/usr/bin/env python
"""Run bench.py with the given arguments and print a one-line digest of its JSON line (experiment helper)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "bench.py", *sys.argv[1:]], capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not lines:
    print("ERR", (out.stderr or out.stdout)[-800:])
    sys.exit(1)
d = json.loads(lines[-1])
cfg = d.get("config", {})
print(" ".join(sys.argv[1:]), "| value %.4e" % d["value"], "ms/step %.3f" % d["ms_per_step"],
      "e2e %.4e" % (d.get("e2e", {}).get("value") or 0), "launches", d.get("gpu_launches"),
      "inter/step", cfg.get("interactions_per_step_per_gpu"), "clk", d.get("clocks", {}).get("sm_mhz"),
      "cpu %.3e" % d.get("cpu_baseline", {}).get("value", 0))
