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


def cpu_reference_run(n, threads=0):
    """One pass of the reference-faithful CPU restatement over n trajectories; returns steps/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pnde_ref as R

    u0, p = make_inputs(n)
    t0 = time.perf_counter()
    out = R.solve_ensemble("fhn_readme", "EK1", ORDER, u0.T, p.T, (0.0, T1), adaptive=False, dt=DT,
                           nthreads=threads, want_cov=False)
    dt = time.perf_counter() - t0
    steps = int(np.sum(out["naccept"] + out["nreject"]))
    return steps / dt, steps, dt, R.max_threads()


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
        _, steps, dt, _ = cpu_reference_run(n)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--trajectories", type=int, default=N_TRAJ_PER_GPU, help="per GPU (default: the BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
            "peak_source": "DFMA micro-benchmark measured in this run (pnde_measure_fp64_peak); MEASURED_PEAKS.json "
                           "has no FP64 entry" if fp64_peak else "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz",
            "nominal_peak": nominal, "flop_per_unit": FLOP_PER_STEP, "units_per_launch": steps_per_launch,
            "launch_ms": k_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel for this workload, one ncu --set full
            # capture (profiles/r1_filter_kernel_ncu_summary.csv): 51.6 MB + 349.5 MB; algorithmic bytes 408 MB
            "traffic": 401.0e6 if n == N_TRAJ_PER_GPU else None, "traffic_unit": "bytes per launch",
            "algorithmic_bytes": (BYTES_IN_PER_TRAJ + BYTES_OUT_PER_TRAJ) * n,
            "hbm": {"achieved": (BYTES_IN_PER_TRAJ + BYTES_OUT_PER_TRAJ) * n / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": hbm_src,
                    "note": "state lives in registers across the time loop: HBM is touched once per trajectory"},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pnde_ref as R

            cores = R.max_threads()
            n_cpu = max(256, 1500 * cores)
            v, steps, dt_cpu, _ = cpu_reference_run(n_cpu)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_cpu} of {n} trajectories x {STEPS_PER_TRAJ} steps ({dt_cpu:.1f} s), "
                             "oracle/pnde_ref.c (reference-faithful dense algorithm, pthreads)"}
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
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
