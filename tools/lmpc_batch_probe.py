"""Why does throughput fall with the batch size?  Samples SM clocks / power while a large batch runs and prints the phase counters."""
import json, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W
PH, MAXIT = 20, 250
samples = []
stop = False
def sampler():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append((time.perf_counter(), o))
        time.sleep(0.05)
for B in [int(v) for v in sys.argv[1:]] or [1776, 3552, 4096, 32768]:
    c = W.build_quadrotor_controller(L, PH, B, MAXIT)
    import os
    if os.environ.get('GENERIC'): c.set_launch(-int(os.environ['GENERIC']), 0)
    if os.environ.get('NOHIST'): c.set_history_order(False)
    if os.environ.get('WPC'): c.set_launch(int(os.environ['WPC']), int(os.environ.get('CPS', 0)))
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
    u0 = np.zeros((B, 4))
    c.optimize(x0, u0)
    c.profile()
    samples.clear(); stop = False
    th = threading.Thread(target=sampler); th.start()
    time.sleep(0.2)
    t0 = time.perf_counter()
    ts = []
    for _ in range(3):
        t = time.perf_counter(); res = c.optimize(x0, u0); ts.append(time.perf_counter() - t)
    t1 = time.perf_counter()
    stop = True; th.join()
    p = c.profile(fetch=True)[:, :8].astype(float)
    inside = [s for (t, s) in samples if t0 <= t <= t1]
    print(json.dumps(dict(batch=B, solves_per_s=[B / t for t in ts], cycles_per_solve_mean=float(p.sum(axis=1).mean()),
                          sweeps_cycles=float(p[:, 2].mean()), clocks=inside[:12])), flush=True)
    del c
