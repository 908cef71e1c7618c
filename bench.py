#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: ensemble filter steps/s (traj x steps).

Workload (BASELINE.json configs[1]): FitzHugh-Nagumo parameter sweep,
1e6 trajectories per GPU, EK1(order=3), fixed dt = 0.01 on (0, 20) => 2000 filter steps per
trajectory, filtering only, final state saved.  One bench "step" = one pass of the hot path over the
whole ensemble (2e9 trajectory-filter-steps per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the ensemble shards with no collective on the data
path (weak scaling: 1e6 trajectories per GPU).  `--impl reference` times the CPU restatement of the
reference's dense algorithm (oracle/pnde_ref.c; Julia is not installed in this image) on all host
threads for a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ensemble filter steps/s (traj x steps)"
UNIT = "filter-steps/s"
N_TRAJ_PER_GPU = 1_000_000
T1, DT, ORDER = 20.0, 0.01, 3
STEPS_PER_TRAJ = 2000
FLOP_PER_STEP = 1.9e3      # F_filt(d=2, q=3), the algorithmic figure
# flops the kernel EXECUTES per step: 610 DFMA + 181 DMUL + 59 DADD (profiles/r1_filter_kernel_loop_instruction_mix.csv)
FLOP_EXECUTED_PER_STEP = 2 * 610 + 181 + 59
BYTES_IN_PER_TRAJ = 5 * 8  # u0 (2) + p (3)
BYTES_OUT_PER_TRAJ = (8 + 36 + 2) * 8  # final mean, packed covariance, t, log-likelihood
SEED = 20260118


def make_inputs(n, offset=0):
    """a, b ~ U(0.1, 0.3), c ~ U(2, 4), u0 = (-1, 1); trajectory i uses draw i."""
    import numpy as np

    rng = np.random.default_rng(SEED + offset)
    p = np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2.0, 4.0, n)], axis=0)
    u0 = np.stack([np.full(n, -1.0), np.full(n, 1.0)], axis=0)
    return np.ascontiguousarray(u0), np.ascontiguousarray(p)  # SoA [k][n]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(n, threads=0, inputs=None):
    """One pass of the reference-faithful CPU restatement over n trajectories (the first n of `inputs` when given);
    returns steps/s, steps, seconds, threads, and the final means [n, D]."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pnde_ref as R

    u0, p = inputs if inputs is not None else make_inputs(n)
    u0, p = u0[:, :n], p[:, :n]
    t0 = time.perf_counter()
    out = R.solve_ensemble("fhn_readme", "EK1", ORDER, u0.T, p.T, (0.0, T1), adaptive=False, dt=DT,
                           nthreads=threads, want_cov=False)
    dt = time.perf_counter() - t0
    steps = int(np.sum(out["naccept"] + out["nreject"]))
    return steps / dt, steps, dt, R.max_threads(), out["mean"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pnde_ref as R

    cores = R.max_threads()
    n = max(256, 400 * cores)  # bounded sample: a few seconds per step on the host cores
    for _ in range(args.warmup):
        cpu_reference_run(max(64, n // 8))
    tot_steps, tot_t = 0, 0.0
    for _ in range(args.steps):
        _, steps, dt, _, _ = cpu_reference_run(n)
        tot_steps += steps
        tot_t += dt
    value = tot_steps / tot_t
    sample = f"{n} of {N_TRAJ_PER_GPU} trajectories x {STEPS_PER_TRAJ} steps per bench step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference's dense algorithm (oracle/pnde_ref.c), one trajectory per task over "
                "all host threads; Julia is not installed, so this stands in for EnsembleThreads and is faster than "
                "the real package (no allocation / dispatch overhead)",
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "FitzHugh-Nagumo parameter-sweep ensemble, EK1(order=3), fixed dt=0.01, tspan (0,20), "
                        "filtering only (BASELINE configs[1])",
            "trajectories_per_gpu": N_TRAJ_PER_GPU, "trajectories_total": N_TRAJ_PER_GPU * n_gpus,
            "steps_per_trajectory": STEPS_PER_TRAJ, "d": 2, "order": ORDER, "state_dim": 8,
            "diffusion": "dynamic", "save": "final state", "parallelism": f"ensemble-shard x{n_gpus}",
            "l2": "flushed between timed steps (256 MiB memset); outputs (368 MB) exceed L2"}


def run_config5(args):
    """BASELINE configs[4]: Lotka-Volterra ensemble, 1e6 trajectories IN TOTAL sharded over the GPUs (strong scaling),
    EK1(order=3), fixed dt = 0.05 on (0, 10) => 200 steps, filter with every step saved + RTS smoother over the full
    grid, processed in waves of 250 k trajectories per GPU (history: 201 x 288 B + smoothed 201 x 352 B per
    trajectory).  One bench step = filter + smoother over the rank's whole shard; value = filter steps / time."""
    import numpy as np
    import torch

    import odefilters_b200 as B

    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    total, wave, nsteps = 1_000_000, 250_000, 200
    lo, hi = B.shard_range(total, rank, world)
    rng = np.random.default_rng(SEED)
    p_all = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (total, 4)))
    p_np = np.ascontiguousarray(p_all[lo:hi].T)
    u0_np = np.ones((2, hi - lo))
    prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
    solver = B.FilterSolver(prob, B.EK1(order=3, smooth=True), adaptive=False, dt=0.05, save_everystep=True, device=local)
    waves = [(a, min(a + wave, hi - lo)) for a in range(0, hi - lo, wave)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_pass():
        f = s = 0.0
        for a, b in waves:
            solver.upload(np.ascontiguousarray(u0_np[:, a:b]), np.ascontiguousarray(p_np[:, a:b]), soa=True)
            solver.run(sync=False)
            solver.smooth(sync=True)
            fm, sm = solver.last_run_ms()
            f += fm
            s += sm
        return f, s

    for _ in range(args.warmup):
        one_pass()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    w0 = time.perf_counter()
    fms = sms = 0.0
    for _ in range(args.steps):
        f, s = one_pass()
        fms += f
        sms += s
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    c = solver.counts()
    assert (c["retcode"] == 0).all() and (c["naccept"] == nsteps).all()
    t_dev = (fms + sms) * 1e-3
    if world > 1:
        t = torch.tensor([t_dev, fms, sms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, fms, sms = (float(x) for x in t.tolist())
    if rank == 0:
        steps_total = total * nsteps * args.steps
        flop = 1.9e3 + 8.6e3  # F_filt(2,3) + F_smooth(2,3)
        line = {
            "metric": METRIC, "value": steps_total / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Lotka-Volterra ensemble, EK1(order=3), dt=0.05 on (0,10), filter (every step saved) + "
                                   "RTS smoother over the full grid (BASELINE configs[4])",
                       "trajectories_total": total, "trajectories_per_gpu": hi - lo, "wave": wave,
                       "steps_per_trajectory": nsteps, "parallelism": f"ensemble-shard x{world}",
                       "l2": "history per 250 k wave (14.5 GB filtered + 17.7 GB smoothed) far exceeds L2"},
            "filter_ms_per_step": fms / args.steps, "smoother_ms_per_step": sms / args.steps,
            "wall_ms_per_step": 1e3 * wall / args.steps, "clocks": clocks,
            "gpu_launches": args.steps * 2 * len(waves) * world,
            "roofline": {"bound": "fp64", "kernel": "filter_kernel + smoother_kernel (DenseEK1<VfLotkaVolterra,3>)",
                         "achieved": flop * steps_total / t_dev / world / 1e12, "peak": 33.8, "unit": "TFLOP/s",
                         "frac": flop * steps_total / t_dev / world / 1e12 / 33.8, "flop_per_unit": flop,
                         "peak_source": "DFMA micro-benchmark of round 1 (33.8 TFLOP/s)", "traffic": None},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--trajectories", type=int, default=N_TRAJ_PER_GPU, help="per GPU (default: the BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, help="2: the headline (BASELINE configs[1]); 5: Lotka-Volterra "
                    "filter + RTS smoother, 1e6 trajectories sharded over the GPUs (BASELINE configs[4])")
    args = ap.parse_args()
    if args.config == 5:
        return run_config5(args)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import ctypes as C

    import numpy as np
    import torch

    import odefilters_b200 as B
    from odefilters_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = args.trajectories

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # inputs in pinned host memory (SoA, trajectory fastest); rank r gets its own draws
    u0_np, p_np = make_inputs(n, offset=rank)
    u0_pin = torch.from_numpy(u0_np).pin_memory()
    p_pin = torch.from_numpy(p_np).pin_memory()
    mean_pin = torch.empty((8, n), dtype=torch.float64).pin_memory()
    cov_pin = torch.empty((36, n), dtype=torch.float64).pin_memory()
    t_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    ll_pin = torch.empty(n, dtype=torch.float64).pin_memory()

    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, T1), (0.2, 0.2, 3.0))
    solver = B.FilterSolver(prob, B.EK1(order=ORDER, smooth=False), adaptive=False, dt=DT, save_everystep=False,
                            device=local)
    lib, h = solver.lib, solver._h

    def upload():
        solver._check(lib.pnde_upload(h, n, u0_pin.data_ptr(), p_pin.data_ptr()), "pnde_upload")
        solver.n = n

    def fetch():
        solver._check(lib.pnde_get_final(h, mean_pin.data_ptr(), cov_pin.data_ptr(), t_pin.data_ptr(),
                                         ll_pin.data_ptr()), "pnde_get_final")

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    upload()
    solver.synchronize()
    for _ in range(args.warmup):
        solver.run()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident timing: K steps, CUDA events on the launching stream -------------
    barrier()
    ev_ms = []
    w0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.zero_()
        torch.cuda.synchronize()
        solver.run()  # synchronises the handle's stream
        ev_ms.append(solver.last_run_ms()[0])
    barrier()
    wall = time.perf_counter() - w0
    launches = args.steps * solver.launch_count() * world  # every rank launches the same kernels
    fetch()
    steps_done = int(np.sum(solver.counts()["naccept"]))
    assert steps_done == n * STEPS_PER_TRAJ, steps_done
    assert float(t_pin.min()) == T1 and bool(torch.isfinite(mean_pin).all())
    t_dev = max_over_ranks(sum(ev_ms) * 1e-3)
    total_steps = steps_done * world * args.steps
    value = total_steps / t_dev
    # ---- end to end through the C ABI with host buffers (H2D + solve + D2H inside) -------
    def e2e_call():
        solver._check(lib.pnde_solve_ensemble_to_host(h, n, u0_pin.data_ptr(), p_pin.data_ptr(), mean_pin.data_ptr(),
                                                      cov_pin.data_ptr(), t_pin.data_ptr(), ll_pin.data_ptr()),
                      "pnde_solve_ensemble_to_host")

    e2e_call()  # untimed warm-up of this entry point (creates its copy stream and events)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        # one C-ABI call with host buffers in and out: H2D, the sliced solve, D2H of every finished slice
        e2e_call()
    torch.cuda.synchronize()
    assert float(t_pin.min()) == T1 and bool(torch.isfinite(mean_pin).all())
    t_e2e = max_over_ranks(time.perf_counter() - e0)
    barrier()
    clocks = sampler.stop()
    e2e_value = total_steps / t_e2e
    # ---- strong scaling: the BASELINE ensemble (1e6 trajectories in total) split over the ranks ----------------
    strong = None
    if world > 1:
        n_s = N_TRAJ_PER_GPU // world + (1 if rank < N_TRAJ_PER_GPU % world else 0)
        solver.n = n_s
        u0_s, p_s = u0_pin[:, :n_s].contiguous(), p_pin[:, :n_s].contiguous()  # keep alive across the call
        solver._check(lib.pnde_upload(h, n_s, u0_s.data_ptr(), p_s.data_ptr()), "pnde_upload")
        for _ in range(2):
            solver.run()
        barrier()
        s_ms = []
        for _ in range(args.steps):
            flush_buf.zero_()
            torch.cuda.synchronize()
            solver.run()
            s_ms.append(solver.last_run_ms()[0])
        barrier()
        t_s = max_over_ranks(sum(s_ms) * 1e-3)
        strong = {"trajectories_total": N_TRAJ_PER_GPU, "trajectories_per_gpu": N_TRAJ_PER_GPU // world,
                  "value": N_TRAJ_PER_GPU * STEPS_PER_TRAJ * args.steps / t_s, "unit": UNIT,
                  "ms_per_step": 1e3 * t_s / args.steps,
                  "note": "same kernel, 1e6 trajectories in total: device time, max over ranks"}

    if rank == 0:
        peak = C.c_double()
        rc = lib.pnde_measure_fp64_peak(local, C.byref(peak))
        fp64_peak = peak.value if rc == 0 else None
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json"
        k_ms = sum(ev_ms) / len(ev_ms)
        steps_per_launch = n * STEPS_PER_TRAJ
        achieved_tf = FLOP_PER_STEP * steps_per_launch / (k_ms * 1e-3) / 1e12
        nominal = 37.2
        roofline = {
            "bound": "fp64", "kernel": "filter_kernel<DenseEK1<VfFhnReadme,3>,false>",
            "achieved": achieved_tf, "peak": fp64_peak if fp64_peak else nominal, "unit": "TFLOP/s",
            "frac": achieved_tf / (fp64_peak if fp64_peak else nominal),
            # the same with the flops the kernel actually executes (fewer than the survey's algorithmic count)
            "flop_executed_per_unit": FLOP_EXECUTED_PER_STEP,
            "achieved_executed": achieved_tf * FLOP_EXECUTED_PER_STEP / FLOP_PER_STEP,
            "frac_executed": achieved_tf * FLOP_EXECUTED_PER_STEP / FLOP_PER_STEP / (fp64_peak if fp64_peak else nominal),
            "frac_of_nominal": achieved_tf / nominal,
            "peak_source": "DFMA micro-benchmark measured in this run (pnde_measure_fp64_peak); MEASURED_PEAKS.json "
                           "has no FP64 entry" if fp64_peak else "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz",
            "nominal_peak": nominal, "flop_per_unit": FLOP_PER_STEP, "units_per_launch": steps_per_launch,
            "launch_ms": k_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel for this workload, one ncu --set full
            # capture of the final round-2 build (profiles/r2_filter_kernel_ncu_summary.csv): 50.5 MB + 345.4 MB
            # (round 1: 51.6 + 349.5); algorithmic bytes 408 MB
            "traffic": 395.8e6 if n == N_TRAJ_PER_GPU else None, "traffic_unit": "bytes per launch",
            "algorithmic_bytes": (BYTES_IN_PER_TRAJ + BYTES_OUT_PER_TRAJ) * n,
            "hbm": {"achieved": (BYTES_IN_PER_TRAJ + BYTES_OUT_PER_TRAJ) * n / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": hbm_src,
                    "note": "state lives in registers across the time loop: HBM is touched once per trajectory"},
        }
        cpu = None
        parity = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pnde_ref as R

            cores = R.max_threads()
            n_cpu = min(n, max(256, 1500 * cores))
            v, steps, dt_cpu, _, ref_mean = cpu_reference_run(n_cpu, inputs=(u0_np, p_np))
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"the first {n_cpu} of the {n} trajectories x {STEPS_PER_TRAJ} steps ({dt_cpu:.1f} s), "
                             "oracle/pnde_ref.c (reference-faithful dense algorithm, pthreads)"}
            # the CPU sample is the same inputs the GPU just solved: check u(t1) of every one of them.  The yardstick
            # is the reference arithmetic's own sensitivity: the same CPU code on inputs moved by one ulp (a few
            # draws of this sweep have condition numbers of 1e9 over 2000 steps; tests/test_gpu_parity.py)
            got = mean_pin[:2, :n_cpu].numpy().T
            sc = np.abs(ref_mean[:, :2]).max(axis=1)
            per = np.abs(got - ref_mean[:, :2]).max(axis=1) / sc
            n_own = n_cpu  # the same sample: the maximum of a heavy-tailed quantity grows with the sample size
            rng = np.random.default_rng(1)
            p_ulp = p_np[:, :n_own] * (1 + 2.2e-16 * rng.choice([-1.0, 0.0, 1.0], (3, n_own)))
            own_mean = cpu_reference_run(n_own, inputs=(u0_np[:, :n_own], p_ulp))[4]
            own = np.abs(own_mean[:, :2] - ref_mean[:n_own, :2]).max(axis=1) / sc[:n_own]
            qs = lambda x: {"median": float(np.median(x)), "p99": float(np.quantile(x, 0.99)), "max": float(x.max())}  # noqa: E731
            parity = {"n": int(n_cpu), "max_rel_u": float(per.max()), "median_rel_u": float(np.median(per)),
                      "p99_rel_u": float(np.quantile(per, 0.99)),
                      "against": "oracle/pnde_ref.c on the same inputs, full 2000 steps",
                      "reference_vs_itself_inputs_moved_1ulp": dict(n=int(n_own), **qs(own)),
                      "criterion": "median < 1e-11 and max < 5 x the reference arithmetic's own 1-ulp sensitivity",
                      "ok": bool(np.median(per) < 1e-11 and per.max() < 5 * max(own.max(), 1e-9))}
            assert parity["ok"], parity
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BYTES_IN_PER_TRAJ * n,
                    "d2h_bytes_per_step": BYTES_OUT_PER_TRAJ * n, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "strong": strong,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
