#!/usr/bin/env python
"""bench.py -- batched LMPC solves/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # our CUDA engine
  python bench.py --impl reference ...                     # the CPU oracle (restated reference path) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # weak scaling: every rank owns `--batch` instances

A "step" is one batched IOptimizer::run over all instances of the workload configs[1] of BASELINE.json:
quadrotor LMPC nx=12 nu=4 ny=12 ph=ch=20, batch 4096 per GPU, maximum_iteration=250, synthetic x0 / yRef (seed 20,
instance b drawn from PCG64(seed+b)).  Prints ONE JSON line (rank 0).
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

NX, NU, NDU, NY = 12, 4, 4, 12
X0_SCALE = np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)


def synth_inputs(first, count, seed=20):
    """x0 ~ U(-1,1)*scale clipped into the state box, u0 = 0, yRef = [0,0,r,0..], r ~ U(0.5,1.5)."""
    x0 = np.empty((count, NX))
    r = np.empty(count)
    for k in range(count):
        g = np.random.Generator(np.random.PCG64(seed + first + k))
        x0[k] = g.uniform(-1, 1, NX) * X0_SCALE
        r[k] = g.uniform(0.5, 1.5)
    x0[:, 0:2] = np.clip(x0[:, 0:2], -np.pi / 6, np.pi / 6)
    x0[:, 5] = np.maximum(x0[:, 5], -1.0)
    return x0, r


def algorithmic_bytes(ph, shared_model):
    """per-instance model + weights + bounds + (x0,u0) + references over the horizon + outputs; the
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


def build_controller(L, ph, batch, max_iter, per_instance_model=False):
    from oracle.lmpc_formulation import quadrotor_formulation, quadrotor_model
    f = quadrotor_formulation(ph)
    c = L.LMPC(NX, NU, NDU, NY, ph, ph, batch=batch, device=int(os.environ.get("LOCAL_RANK", 0)))
    Ad, Bd = quadrotor_model()
    if per_instance_model:   # batch copies of A,B,C (the per-instance-model variant of SURVEY 8d)
        c.setStateSpaceModel(np.broadcast_to(Ad, (batch, NX, NX)), np.broadcast_to(Bd, (batch, NX, NU)),
                             np.broadcast_to(np.eye(NX), (batch, NY, NX)))
    else:                    # the quadrotor_ex model/weights/bounds, per-instance x0 / yRef
        c.setStateSpaceModel(Ad, Bd, np.eye(NX))
    c.setObjectiveWeights(f.wOutput[:, 1], f.wU[:, 1], f.wDeltaU[:, 0], (0, ph))
    c.setStateBounds(f.minX[:, 1], f.maxX[:, 1], (0, ph))
    c.setInputBounds(f.minU[:, 0], f.maxU[:, 0], (0, ph))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=max_iter))
    return f, c


def cpu_oracle_solve(args):
    ph, x0, r, max_iter = args
    from oracle.lmpc_formulation import quadrotor_formulation
    from oracle.osqp_restated import Settings, lmpc_optimize
    f = quadrotor_formulation(ph)
    yr = np.zeros(NY)
    yr[2] = r
    f.set_references(yr, np.zeros(NU), np.zeros(NU))
    t = time.perf_counter()
    res = lmpc_optimize(f, x0, np.zeros(NU), Settings(max_iter=max_iter))
    return time.perf_counter() - t, res["cmd"]


def cpu_baseline(ph, max_iter, n_solves, cores):
    """Times the CPU oracle (the restated reference path) on a bounded sample of the same workload."""
    x0, r = synth_inputs(0, n_solves)
    try:
        from oracle import c_oracle   # C port (oracle/Makefile), preferred when built
        c_oracle.lib()
        return c_oracle.time_batch(ph, x0, r, max_iter, cores)
    except RuntimeError:
        pass
    jobs = [(ph, x0[k], r[k], max_iter) for k in range(n_solves)]
    t = time.perf_counter()
    if cores > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(cpu_oracle_solve, jobs)
    else:
        for j in jobs:
            cpu_oracle_solve(j)
    dt = time.perf_counter() - t
    return {"value": n_solves / dt, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{n_solves} solves of the same workload through oracle/osqp_restated.py (numpy, dense KKT LU)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU (weak scaling)")
    ap.add_argument("--ph", type=int, default=20)
    ap.add_argument("--max-iter", type=int, default=250)
    ap.add_argument("--cpu-sample", type=int, default=0, help="solves in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-instance-model", action="store_true", help="give every instance its own copy of A,B,C")
    ap.add_argument("--schedule", default="gang", choices=["gang", "free"], help="persistent-warp scheduling (include/b200mpc.h)")
    ap.add_argument("--no-history-order", action="store_true",
                    help="do not draw instances in the order of their previous solve's iteration counts (include/b200mpc.h)")
    ap.add_argument("--warps-per-cta", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    ph, B = a.ph, a.batch
    config = {"workload": f"quadrotor LMPC nx=12 nu=4 ny=12 ph=ch={ph}, batch={B} per GPU, maximum_iteration={a.max_iter}, "
                          f"{'per-instance' if a.per_instance_model else 'shared'} model, per-instance x0/yRef synthetic seed 20 (BASELINE.json configs[1])",
              "batch_per_gpu": B, "global_batch": B * world, "ph": ph, "parallelism": f"dp{world}",
              "schedule": a.schedule + ("" if a.no_history_order else ", instances drawn longest-first by the iteration counts of the handle's "
                                         "previous solve (every timed step repeats the same batch, so that history is exact here; "
                                         "value_cold_order is the same measurement without it)"),
              "l2": "L2 flushed (512 MiB write) between timed steps; each step timed by its own CUDA-event pair"}

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
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
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
    f, c = build_controller(L, ph, B, a.max_iter, a.per_instance_model)
    if a.warps_per_cta or a.ctas_per_sm:
        c.set_launch(a.warps_per_cta, a.ctas_per_sm)
    c.set_schedule(a.schedule == "gang")
    c.set_history_order(not a.no_history_order)
    stream = torch.cuda.current_stream()
    c.set_stream(stream.cuda_stream)
    x0_h, r = synth_inputs(rank * B, B)
    yref = np.zeros((B, NY, ph))
    yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((NU, ph)), np.zeros((NU, ph)))
    x0_d = torch.from_numpy(x0_h).cuda()
    u0_d = torch.zeros((B, NU), dtype=torch.float64, device="cuda")
    cmd_d = torch.empty((B, NU), dtype=torch.float64, device="cuda")
    cmd_all = torch.empty((world * B, NU), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def step():
        c.solve_async(x0_d.data_ptr(), u0_d.data_ptr(), dev=True)      # ONE kernel launch
        if world > 1:
            c.get_result_into(cmd_ptr=cmd_d.data_ptr())
            dist.all_gather_into_tensor(cmd_all, cmd_d)                  # the one exchange step

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    for s in range(a.steps):
        flush.zero_()
        evs[s][0].record(stream)
        step()
        evs[s][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    times = np.array([e0.elapsed_time(e1) for e0, e1 in evs])
    ms = float(times.mean())
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = c.fetch_result()
    value = world * B / (ms * 1e-3)
    # the same measurement with history ordering off (what the FIRST solve of a handle, or a batch of unrelated problems, gets)
    cold_ms = None
    if not a.no_history_order:
        c.set_history_order(False)
        step(); torch.cuda.synchronize()
        cevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(a.steps, 5))]
        for e0, e1 in cevs:
            flush.zero_()
            e0.record(stream); step(); e1.record(stream)
        torch.cuda.synchronize()
        cold_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in cevs]))
        if world > 1:
            t = torch.tensor([cold_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cold_ms = float(t.item())
        c.set_history_order(True)

    # end-to-end through the public API with host buffers (pinned), H2D + D2H inside the timed region
    x0_pin = torch.from_numpy(x0_h).pin_memory()
    u0_pin = torch.zeros((B, NU), dtype=torch.float64).pin_memory()
    e2e_t = []
    for s in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = c.optimize(x0_pin.numpy(), u0_pin.numpy())      # H2D x0,u0 -> solve -> D2H cmd,cost,status...
        if world > 1:
            c.get_result_into(cmd_ptr=cmd_d.data_ptr())
            dist.all_gather_into_tensor(cmd_all, cmd_d)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            e2e_t.append(dt)
    e2e_ms = float(np.mean(e2e_t)) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = B * (NX + NU) * 8
    d2h = B * (NU * 8 + 8 + 6 * 4)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        kernel_ms = ms  # one kernel per step; at N>1 the all-gather is inside the step time too
        abytes = algorithmic_bytes(ph, not a.per_instance_model) * B
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        flops = sum(algorithmic_flops(ph, int(i), int(u), int(p == 1)) for i, u, p in zip(res.iterations, res.rho_updates, res.status_polish))
        tfl = flops / (kernel_ms * 1e-3) / 1e12
        fp64_peak, fp64_src = 37.0, "nominal B200 FP64 vector"
        try:
            fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["fp64_tflops"]
            fp64_src = "measured DFMA throughput on this pool (tools/ubench.cu -> profiles/fp64_peak.json)"
        except Exception:
            pass
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("batch") == B and tj.get("ph", 20) == ph:      # the capture is of the default launch only
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # per-solve latency of ONE controller (batch = 1: what a single mpc::LMPC<> object sees), host buffers, p50 of 15
        lat1 = None
        if world == 1:
            f1, c1 = build_controller(L, ph, 1, a.max_iter, False)
            yr1 = np.zeros((1, NY, ph)); yr1[:, 2, :] = r[0]
            c1.setReferences(yr1, np.zeros((NU, ph)), np.zeros((NU, ph)))
            ts1 = []
            for s in range(18):
                t0 = time.perf_counter()
                c1.optimize(x0_h[:1], np.zeros((1, NU)))
                ts1.append(time.perf_counter() - t0)
            lat1 = float(np.median(ts1[3:])) * 1e3
            del c1
        line = {
            "metric": "LMPC solves/sec (batched)", "value": value, "unit": "solves/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "value_cold_order": (world * B / (cold_ms * 1e-3)) if cold_ms else None,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            # per step: the solve kernel, plus the one-CTA ordering kernel when history ordering is on (profiles/r01_launches.csv)
            "gpu_launches": a.steps * (1 if a.no_history_order else 2),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic bytes/solve (SURVEY 8d) x batch / kernel time; the solve is bound by instruction issue / "
                                 "dependent latency of the stage recurrence, not by HBM: see roofline_fp64, DESIGN.md 5, profiles/r01_icache.md"},
            "roofline_fp64": {"bound": "fp64-pipe", "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak,
                              "peak_source": fp64_src,
                              "flops_per_solve_mean": flops / B},
            "solver": {"iterations_mean": float(res.iterations.mean()), "iterations_max": int(res.iterations.max()),
                       "rho_updates_mean": float(res.rho_updates.mean()), "solved": int((res.solver_status == 1).sum()),
                       "polished": int((res.status_polish == 1).sum()), **c.info()},
            "p50_latency_us_per_solve": 1e3 * ms / B,
            "latency": {"amortised_us_per_solve": 1e3 * ms / B, "single_controller_p50_ms": lat1,
                        "note": "single_controller = batch 1 through the public API with host buffers (one mpc::LMPC<> object)"},
            "clocks": sampler.summary(),
        }
        if world == 1 and not a.no_cpu_baseline:
            cores = 1
            n = a.cpu_sample or 2048
            line["cpu_baseline"] = cpu_baseline(ph, a.max_iter, n, cores)
            line["latency"]["cpu_port_ms_per_solve"] = 1e3 / line["cpu_baseline"]["value"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
