#!/usr/bin/env python
"""bench.py -- batched MPC solves/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # our CUDA engine
  python bench.py --impl reference ...                     # the CPU oracle (restated reference path) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # weak scaling: every rank owns `--batch` controllers
  ... --global-batch 65536 --ph 50                         # strong scaling: a fixed global batch split over the ranks (configs[4])

A "step" is one batched IOptimizer::run over all controllers of BASELINE.json configs[1]: quadrotor LMPC nx=12 nu=4 ny=12
ph=ch=20, batch 4096 per GPU, maximum_iteration=250, synthetic x0 / yRef (seed 20, controller b drawn from PCG64(seed+b);
SURVEY.md 8d).  The timed steps form a CLOSED LOOP: after every solve the plant advances, x+ = A x + B u, u0 <- cmd
(examples/quadrotor_ex.cpp), so every step solves a new batch and the scheduling history is real, not replayed.
Prints ONE JSON line (rank 0).  The same line carries `nlmpc`: BASELINE configs[2] (unicycle nx3 nu2 Tph30, batch 1024) and
configs[3] (networked oscillators nx8 nu4 Tph15, batch 8192 globally) with their own CPU baselines.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from libmpc_b200 import workloads as W      # noqa: E402  (data only: no oracle, no CUDA)

NX, NU, NDU, NY = W.QUAD_NX, W.QUAD_NU, W.QUAD_NDU, W.QUAD_NY


def algorithmic_bytes(ph, shared_model):
    """per-controller model + weights + bounds + (x0,u0) + references over the horizon + outputs; the
    shared-model variant drops the model/weights/bounds term (3 296 B for the quadrotor)."""
    nx, nu, ny = NX, NU, NY
    model = 8 * (nx * nx + nx * nu + ny * nx + (ny + 2 * nu) + 2 * (nx + nu + ny))
    return (0 if shared_model else model) + 8 * ((nx + nu) + ph * (ny + 2 * nu)) + (8 * nu + 16)


def algorithmic_flops(ph, iters, rho_updates, polished):
    """F_iter=(ph+1)6b^2+4nnz(A)+10(n+m); F_factor=(ph+1)(7/3)b^3; polish = factor + 4 iters."""
    b = NX + 2 * NU
    ne = NX + NU
    n = (ph + 1) * ne + ph * NU
    m = 2 * (ph + 1) * ne + (ph + 1) * NY + ph * NU + (ph + 1)
    nnzA = 2 * (ph + 1) * ne + ph * 98 + (ph + 1) * NY + ph * NU
    f_iter = (ph + 1) * 6 * b * b + 4 * nnzA + 10 * (n + m)
    f_fac = (ph + 1) * (7.0 / 3.0) * b ** 3
    return iters * f_iter + (1 + rho_updates) * f_fac + polished * (f_fac + 4 * f_iter)


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = False
        self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ---- CPU baselines (the restated reference path, oracle/): ONLY here and in --impl reference -----------------------------------
def cpu_baseline(ph, max_iter, n_solves, cores):
    """Times the CPU oracle (the restated reference path) on a bounded sample of the same workload."""
    from oracle import c_oracle
    x0, r = W.quadrotor_inputs(0, n_solves)
    c_oracle.lib()
    return c_oracle.time_batch(ph, x0, r, max_iter, cores)


def cpu_latency_p50(ph, max_iter, n=24):
    """p50 of single solves (one controller, one core) through the C port: what one mpc::LMPC<>::optimize costs on the host."""
    from oracle import c_oracle
    x0, r = W.quadrotor_inputs(0, n)
    ts = []
    for k in range(n):
        t = time.perf_counter()
        c_oracle.time_batch(ph, x0[k:k + 1], r[k:k + 1], max_iter, 1)
        ts.append(time.perf_counter() - t)
    return float(np.median(ts[2:])) * 1e3


def _nlmpc_cpu_one(k):
    """One cold-start solve of BASELINE configs[2] (unicycle) through the numpy restatement + SciPy SLSQP."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import nlmpc_slsqp as S
    from user_systems import unicycle_formulation
    x0, params = W.unicycle_inputs(k, 1)
    f = unicycle_formulation(params=params[0])
    lb, ub = W.soft_bounds(f.nz)
    z0 = W.cold_start(x0, np.zeros(2), 30, 30)[0]
    t = time.perf_counter()
    r = S.solve(f, x0[0], z0, lb, ub, maxiter=300)
    return time.perf_counter() - t, bool(r["success"])


def nlmpc_cpu_baseline(name, cores):
    """SLSQP (SciPy's compiled Kraft core: the algorithm NLopt's LD_SLSQP implements) on the restated formulation, finite differences
    included, one solve stream per core; a bounded sample of the same cold-start workload (a single solve takes 10-25 s)."""
    n = max(1, min(cores, 16))
    t = time.perf_counter()
    if name == "unicycle":                       # user-defined system: numpy restatement (tests/user_systems.py)
        import multiprocessing as mp
        with mp.get_context("fork").Pool(n) as pool:
            res = pool.map(_nlmpc_cpu_one, list(range(n)))
        dt = time.perf_counter() - t
        conv, how = int(sum(1 for _, ok in res if ok)), "oracle/nlmpc_formulation.py (numpy) callbacks"
    else:                                        # built-in system: the C restatement of the callbacks (oracle/nlmpc_oracle.c)
        from oracle import nlmpc_c_oracle as CO
        x0, params = W.oscnet4_inputs(0, n)
        lb, ub = W.hard_bounds(15 * 8 + 8 * 4 + 1)
        z0 = W.cold_start(x0, np.zeros(4), 15, 8)
        r = CO.time_batch(1, 15, 8, x0, z0, params, lb, ub, cores=n)
        dt = time.perf_counter() - t
        conv, how = int(r["converged"]), "oracle/nlmpc_oracle.c (C) callbacks"
    return {"value": n / dt, "unit": "solves/s", "cores": n, "kind": "port",
            "sample": f"{n} cold-start solves of the same workload, one per core: SciPy SLSQP (Kraft's algorithm, as NLopt LD_SLSQP) with {how} "
                      "-- the reference's finite-difference objective / constraints restated", "converged": conv}


def nlmpc_flops(nx, nu, ph, ch, K, sqp_it, qp_it):
    """Algorithmic FP64 flops of one solve by the STAGE-STRUCTURED kernel (libmpc_b200/csrc/nlmpc_structured.cuh), from its own
    iteration counters: b = nx+nu (stage block), nb = nu+1 (border), compact Jacobians of 2nx+nu / nx+nu+1 entries per row.
      per SQP iteration: finite-difference evaluation (one model / cost / constraint pass per perturbed variable, ~10 flops
        per touched entry) + ~2 factorisations (QP set-up and polish; adaptive-rho refactorisations are not counted) of
        ph [(14/3) b^3 + 2 nb b^2] multiply-adds (Cholesky, explicit inverse, sub-diagonal, Schur, the two solve-time products, border)
        + the reduced-KKT assembly J' R J on the compact rows;
      per ADMM iteration: the bordered block-tridiagonal solve ph (3 b^2 + 2 b nb) multiply-adds, A x and A' y on the compact rows,
        ~10 flops per variable / row for the relaxation, projection and dual update.
    The dense kernel's O(nz^3) count is NOT used: it would overstate what this kernel executes."""
    b, nb = nx + nu, nu + 1
    me, mi = ph * nx, (ph + 1) * K
    we, wi = 2 * nx + nu, nx + nu + 1
    n = ph * nx + ch * nu + 1
    m = me + mi + n
    fd = (ph * (nx + nu) + 3) * 2 * ph * (nx + nu) * 10
    fac = 2 * (ph * ((14.0 / 3.0) * b ** 3 + 2 * nb * b * b) + me * we * we + mi * wi * wi)     # multiply-adds
    per_sqp = fd + 2.0 * fac
    per_qp = 2.0 * (ph * (3 * b * b + 2 * b * nb) + 2 * (me * we + mi * wi + n)) + 10.0 * (n + m)
    return sqp_it * per_sqp + qp_it * per_qp


def bench_nlmpc(L, torch, rank, world, steps, fp64_peak, with_cpu):
    """BASELINE configs[2] and configs[3] through the C ABI with host buffers (cold start, as the first optimize() of a controller)."""
    out = {}
    cfgs = [("unicycle", "configs[2]: unicycle NLMPC nx=3 nu=2 Tph=Tch=30, 2 obstacle inequalities per stage (Tineq=62), soft constraints, "
             "cold start; batch 1024 per GPU (user-defined system, NVRTC)", 1024, True),
            ("oscnet4", "configs[3]: networked oscillators N=4 NLMPC nx=8 nu=4 Tph=15 Tch=8, u<=0.5 (Tineq=64), cold start; global batch "
             "8192 sharded over the GPUs", 8192 // world, False)]
    for name, desc, B, per_gpu in cfgs:
        first = rank * B
        if name == "unicycle":
            sid = L.register_system(W.UNICYCLE_SRC, W.UNICYCLE_TYPE)
            ph, ch, nx, nu = 30, 30, 3, 2
            x0, params = W.unicycle_inputs(first, B)
            lb, ub = W.soft_bounds(ph * nx + ch * nu + 1)
            kw = dict(max_sqp=300)
            K = 2
        else:
            sid = L.SYS_OSCNET4
            ph, ch, nx, nu = 15, 8, 8, 4
            x0, params = W.oscnet4_inputs(first, B)
            lb, ub = W.hard_bounds(ph * nx + ch * nu + 1)
            kw = {}
            K = 4
        nz = ph * nx + ch * nu + 1
        z0 = W.cold_start(x0, np.zeros(nu), ph, ch)
        L.nlmpc_solve(sid, ph, ch, z0[:8], x0[:8], params if params.ndim == 1 else params[:8], lb, ub, **kw)      # compile / warm up
        ts = []
        for _ in range(steps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = L.nlmpc_solve(sid, ph, ch, z0, x0, params, lb, ub, **kw)
            ts.append(time.perf_counter() - t)
        ms = float(np.mean(ts)) * 1e3
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        flops = float(sum(nlmpc_flops(nx, nu, ph, ch, K, int(a), int(b)) for a, b in zip(r["iters"], r["qp_iters"])))
        tfl = flops * world / (ms * 1e-3) / 1e12
        out[name] = {"workload": desc, "value": world * B / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "batch_per_gpu": B,
                     "e2e": "host buffers in, results out, inside the timed call",
                     "converged": int((r["status"] == 0).sum()), "sqp_iterations_mean": float(r["iters"].mean()),
                     "qp_iterations_mean": float(r["qp_iters"].mean()), "viol_max": float(r["viol"].max()),
                     "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak,
                                  "flops_per_solve_mean": flops / B,
                                  "note": "algorithmic flops of the stage-structured kernel (bench.py nlmpc_flops) from its own SQP / ADMM counters"}}
        if with_cpu and rank == 0:
            try:
                out[name]["cpu_baseline"] = nlmpc_cpu_baseline(name, os.cpu_count() or 1)
            except Exception as e:                                        # noqa: BLE001
                out[name]["cpu_baseline"] = {"error": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="controllers per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: this many controllers in total, split over the ranks")
    ap.add_argument("--ph", type=int, default=20)
    ap.add_argument("--max-iter", type=int, default=250)
    ap.add_argument("--cpu-sample", type=int, default=0, help="solves in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nlmpc", action="store_true", help="skip the NLMPC configurations (configs[2], configs[3])")
    ap.add_argument("--per-instance-model", action="store_true", help="give every controller its own copy of A,B,C")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 warp per controller, 2 CTA per controller (include/b200mpc.h)")
    ap.add_argument("--no-history-order", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    ph = a.ph
    strong = a.global_batch > 0
    B = (a.global_batch + world - 1) // world if strong else a.batch
    config = {"workload": f"quadrotor LMPC nx=12 nu=4 ny=12 ph=ch={ph}, batch={B} per GPU, maximum_iteration={a.max_iter}, "
                          f"{'per-instance' if a.per_instance_model else 'shared'} model, per-controller x0/yRef synthetic seed 20 (BASELINE.json "
                          f"configs[{4 if strong else 1}]); closed loop: every timed step solves from the state the previous command produced",
              "batch_per_gpu": B, "global_batch": B * world, "ph": ph, "parallelism": f"dp{world}",
              "l2": "L2 flushed (512 MiB write) between timed steps; each step timed by its own CUDA-event pair on the solve stream"}

    if a.impl == "reference":
        # the reference's own CPU implementation of the path, restated (oracle/): rank 0 only
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        per_step = a.cpu_sample or 64 * cores
        vals = []
        for s in range(a.warmup + a.steps):
            cb = cpu_baseline(ph, a.max_iter, per_step, cores)
            if s >= a.warmup:
                vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        line = {"impl": "reference", "metric": "LMPC solves/sec (batched)", "value": v, "unit": "solves/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * per_step / v, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": cb, "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import libmpc_b200 as L
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c = W.build_quadrotor_controller(L, ph, B, a.max_iter, a.per_instance_model, device=local_rank)
    if a.engine:
        c.set_engine(a.engine)
    c.set_history_order(not a.no_history_order)
    stream = torch.cuda.current_stream()
    c.set_stream(stream.cuda_stream)
    x0_h, r = W.quadrotor_inputs(rank * B, B)
    yref = np.zeros((B, NY, ph))
    yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((NU, ph)), np.zeros((NU, ph)))
    x_d = torch.from_numpy(x0_h).cuda()
    x_init = x_d.clone()
    u_d = torch.zeros((B, NU), dtype=torch.float64, device="cuda")
    cmd_d = torch.empty((B, NU), dtype=torch.float64, device="cuda")
    cmd_all = torch.empty((world * B, NU), dtype=torch.float64, device="cuda") if world > 1 else None
    comm = None
    if world > 1:
        # the engine's own communicator (include/b200mpc.h: b200mpc_comm_*): torch.distributed only carries the 128-byte id
        uid = [L.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = L.Comm(world, rank, uid[0], local_rank)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    launches = {"n": 0}

    def step(advance=True):
        c.solve_async(x_d.data_ptr(), u_d.data_ptr(), dev=True)          # history ordering (1 CTA) + ONE solve kernel
        launches["n"] += 2 if (not a.no_history_order) else 1
        if world > 1:
            c.allgather_cmd(comm, cmd_all.data_ptr())                      # the one exchange step, through the C ABI
            launches["n"] += 1
        if advance:
            c.advance(x_d.data_ptr(), u_d.data_ptr())                      # x+ = A x + B u, u0 <- cmd (our plant-step kernel)
            launches["n"] += 1

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    launches["n"] = 0
    iters_acc, rho_acc, pol_acc = [], [], []
    for s in range(a.steps):
        flush.zero_()
        evs[s][0].record(stream)
        step()
        evs[s][1].record(stream)
    torch.cuda.synchronize()
    n_launch = launches["n"]
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    times = np.array([e0.elapsed_time(e1) for e0, e1 in evs])
    ms = float(times.mean())
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = c.fetch_result()                       # counters of the LAST timed step (the flop estimate below uses them)
    value = world * B / (ms * 1e-3)

    # the exchange step on its own (N > 1): all-gather of the command block, timed with events
    coll_ms = None
    if world > 1:
        cevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for e0, e1 in cevs:
            e0.record(stream); c.allgather_cmd(comm, cmd_all.data_ptr()); e1.record(stream)
        torch.cuda.synchronize()
        coll_ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in cevs[5:]]))

    # the first solve of a handle on a fresh batch: no scheduling history, nothing warm (what a batch of unrelated problems gets)
    cold_ms = None
    c.set_history_order(False)
    x_d.copy_(x_init); u_d.zero_()
    step(advance=False); torch.cuda.synchronize()
    cevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(a.steps, 5))]
    for e0, e1 in cevs:
        flush.zero_()
        e0.record(stream); step(advance=False); e1.record(stream)
    torch.cuda.synchronize()
    cold_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in cevs]))
    if world > 1:
        t = torch.tensor([cold_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_ms = float(t.item())
    c.set_history_order(not a.no_history_order)

    # end-to-end through the public API with host buffers (pinned), H2D + D2H inside the timed region; closed loop on the host
    x_pin = torch.from_numpy(x0_h.copy()).pin_memory()
    u_pin = torch.zeros((B, NU), dtype=torch.float64).pin_memory()
    Ad, Bd = W.quadrotor_model()
    e2e_t = []
    for s in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = c.optimize(x_pin.numpy(), u_pin.numpy())      # H2D x0,u0 -> solve -> D2H cmd,cost,status...
        if world > 1:
            c.allgather_cmd(comm, cmd_all.data_ptr())
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            e2e_t.append(dt)
        xn = x_pin.numpy() @ Ad.T + out.cmd @ Bd.T            # the plant (host side of the loop, outside the timed region)
        x_pin.numpy()[:] = xn; u_pin.numpy()[:] = out.cmd
    e2e_ms = float(np.mean(e2e_t)) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = B * (NX + NU) * 8
    d2h = B * (NU * 8 + 8 + 6 * 4)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp64_peak, fp64_src = 37.0, "nominal B200 FP64 vector (no FP64 entry in MEASURED_PEAKS.json)"
    try:
        fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["fp64_tflops"]
        fp64_src = ("DFMA throughput measured on this pool by tools/ubench.cu (profiles/fp64_peak.json); MEASURED_PEAKS.json has no FP64 "
                    "entry; nominal 37")
    except Exception:
        pass

    nl = None
    if not a.no_nlmpc:
        nl = bench_nlmpc(L, torch, rank, world, max(1, min(a.steps, 2)), fp64_peak, with_cpu=(world == 1 and not a.no_cpu_baseline))

    if rank == 0:
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        abytes = algorithmic_bytes(ph, not a.per_instance_model) * B
        achieved = abytes / (ms * 1e-3) / 1e9
        flops = sum(algorithmic_flops(ph, int(i), int(u), int(p == 1)) for i, u, p in zip(res.iterations, res.rho_updates, res.status_polish))
        tfl = flops / (ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("batch") == B and tj.get("ph", 20) == ph and tj.get("engine") == c.get_engine()["engine"]:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # BASELINE's second metric: p50 per-solve latency of ONE controller (batch 1 = one mpc::LMPC<> object) through the public
        # API with host buffers, over 200 timed solves on successive closed-loop states, next to the CPU port's p50
        lat = None
        if world == 1:
            c1 = W.build_quadrotor_controller(L, ph, 1, a.max_iter, False)
            yr1 = np.zeros((1, NY, ph)); yr1[:, 2, :] = r[0]
            c1.setReferences(yr1, np.zeros((NU, ph)), np.zeros((NU, ph)))
            x1 = x0_h[:1].copy(); u1 = np.zeros((1, NU))
            ts1 = []
            for s in range(210):
                t0 = time.perf_counter()
                o1 = c1.optimize(x1, u1)
                ts1.append(time.perf_counter() - t0)
                if s % 30 == 29:
                    x1 = x0_h[(s // 30) % B: (s // 30) % B + 1].copy(); u1 = np.zeros((1, NU))   # a new transient every 30 steps
                else:
                    x1 = x1 @ Ad.T + o1.cmd @ Bd.T; u1 = o1.cmd.copy()
            ts1 = np.array(ts1[10:]) * 1e3
            lat = {"single_controller_p50_ms": float(np.percentile(ts1, 50)), "single_controller_p95_ms": float(np.percentile(ts1, 95)),
                   "solves_timed": int(ts1.size), "engine": c1.get_engine(),
                   "note": "batch 1 through the public API with host buffers (one mpc::LMPC<> object), closed loop"}
            del c1
        line = {
            "metric": "LMPC solves/sec (batched)", "value": value, "unit": "solves/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "value_cold_order": (world * B / (cold_ms * 1e-3)) if cold_ms else None,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            # per step: the one-CTA ordering kernel, the solve kernel, the plant-step kernel (profiles/*launches.csv)
            "gpu_launches": n_launch,
            "collective_ms": coll_ms,
            # the governing bound of an on-chip solver is the FP64 pipe; HBM (algorithmic bytes and measured traffic) is secondary
            "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak,
                         "traffic": traffic, "peak_source": fp64_src, "flops_per_solve_mean": flops / B,
                         "note": "algorithmic flops (SURVEY 8d formulae with the kernel's own per-controller iteration / rho-update / polish "
                                 "counters of the last timed step) / step time"},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "note": "algorithmic bytes/solve (SURVEY 8d) x batch / step time; traffic = dram bytes per launch from the ncu "
                                     "capture of this launch (profiles/traffic.json)"},
            "solver": {"iterations_mean": float(res.iterations.mean()), "iterations_max": int(res.iterations.max()),
                       "rho_updates_mean": float(res.rho_updates.mean()), "solved": int((res.solver_status == 1).sum()),
                       "polished": int((res.status_polish == 1).sum()), "engine": c.get_engine(), **c.info()},
            "latency": lat,
            "clocks": sampler.summary(),
        }
        if nl is not None:
            line["nlmpc"] = nl
        if world == 1 and not a.no_cpu_baseline:
            n = a.cpu_sample or 2048
            line["cpu_baseline"] = cpu_baseline(ph, a.max_iter, n, 1)
            if lat is not None:
                lat["cpu_port_p50_ms"] = cpu_latency_p50(ph, a.max_iter)
        print(json.dumps(line))
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
