"""Run bench.py with the given extra args and print a one-line summary (tool for quick A/B runs on the GPU box)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "2", "--no-cpu-baseline"] + sys.argv[2:], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(sys.argv[1], "solves/s", round(d["value"]), "ms", round(d["ms_per_step"], 2), "slots", d["solver"]["warp_slots"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, out.stdout[-300:], out.stderr[-600:])
