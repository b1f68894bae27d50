"""CPU baseline of the NLMPC workloads of tools/bench_nlmpc.py: the reference's solve path as far as it can be rebuilt without
NLopt -- SciPy's compiled SLSQP core driving the restated formulation evaluated in C (oracle/nlmpc_c_oracle.py), one solve
stream per core.  Same synthetic inputs as the GPU bench (seed 0), bounded samples.  usage: python tools/bench_nlmpc_cpu.py [cores]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import nlmpc_c_oracle as CO
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation

cores = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)


def run(name, system, f, ph, ch, sample, hard, lo, hi):
    lb, ub = S.default_bounds(f, hard)
    if not hard:
        lb[-1] = 0.0
    x0 = np.random.default_rng(0).uniform(lo, hi, (sample, f.nx))          # the first `sample` instances of the GPU batch
    z0 = np.concatenate([np.tile(x0, (1, ph)), np.zeros((sample, ch * f.nu + 1))], axis=1)
    r = CO.time_batch(system, ph, ch, x0, z0, f.params, lb, ub, cores=cores)
    print(json.dumps(dict(workload=name, nz=int(f.nz), sample=sample, cores=cores, kind="port (SciPy SLSQP core + C callbacks)",
                          solves_per_s=r["solves_per_s"], solves_per_s_per_core=r["solves_per_s"] / cores, converged=r["converged"])), flush=True)


if __name__ == "__main__":
    f = vanderpol_formulation(); f.params = np.array([0.1])
    run("vanderpol_ex nx2 nu1 ph10 ch5", 0, f, 10, 5, 64 * cores, True, -1.5, 1.5)
    f = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    run("ugv_ex nx4 nu2 ph10 ch10", 3, f, 10, 10, 8 * cores, False, -0.3, 0.6)
    f = ugv_formulation(30, 30, v_pref=(0.6, 0.8))
    run("ugv_ex nx4 nu2 ph30 ch30", 3, f, 30, 30, 2 * cores, False, -0.3, 0.6)
    f = oscnet_formulation(4, 15, 8); f.params = np.array([0.1, 1.0, 0.1])
    run("networked_oscillators_ex nx8 nu4 ph15 ch8", 1, f, 15, 8, 4 * cores, True, -1.0, 1.0)
