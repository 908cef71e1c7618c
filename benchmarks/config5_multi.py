#!/usr/bin/env python
"""BASELINE configs[4]: Lotka-Volterra ensemble, 1e6 trajectories, EK1(order=3), dt = 0.05 on (0, 10), filter + RTS
smoother over the full grid, sharded over the ranks of one node (strong scaling: the total is fixed).

    python benchmarks/config5_multi.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/config5_multi.py

Each rank processes its contiguous shard in waves of <= 250k trajectories (history + smoothed history of a wave:
250k x 201 x (288 + 352) B = 32 GB) and copies the smoothed marginals (u, Sigma_u) of the final wave's first
trajectories back as a spot check.  No collective on the data path; NCCL only for the barrier / max reduction.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import odefilters_b200 as B  # noqa: E402

N_TOTAL, WAVE, STEPS = 1_000_000, 250_000, 200


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lo, hi = B.shard_range(N_TOTAL, rank, world)
    rng = np.random.default_rng(20260118 + rank)
    n = hi - lo
    p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
    prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
    solver = B.FilterSolver(prob, B.EK1(order=3, smooth=True), adaptive=False, dt=0.05, save_everystep=True, device=local)
    waves = [(a, min(n, a + WAVE)) for a in range(0, n, WAVE)]

    to_host = "--marginals" in sys.argv  # also bring the smoothed marginals (t, u, Sigma_u) of every step to the host
    d2h = [0.0, 0]
    nrec = min(WAVE, n) * (STEPS + 1)
    pinned = (B.pinned_empty(nrec), B.pinned_empty((nrec, 2)), B.pinned_empty((nrec, 3))) if to_host else None

    def one_pass():
        f_ms = s_ms = 0.0
        for a, b in waves:
            solver.upload(np.ones((b - a, 2)), p[a:b])
            solver.run()
            solver.smooth()
            fm, sm = solver.last_run_ms()
            f_ms += fm
            s_ms += sm
            if to_host:
                t1 = time.perf_counter()
                _, t, u, cu, _ = solver.history(1, 0, b - a, marginals=True, out=pinned)
                d2h[0] += time.perf_counter() - t1
                d2h[1] += t.nbytes + u.nbytes + cu.nbytes
        return f_ms, s_ms

    one_pass()  # warm-up
    d2h[:] = [0.0, 0]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    f_ms, s_ms = one_pass()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    _, t, u, cu, _ = solver.history(1, 0, 2, marginals=True)
    ok = bool(np.isfinite(u).all() and (solver.counts()["naccept"] == STEPS).all())
    dev = torch.tensor([f_ms, s_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev, op=dist.ReduceOp.MAX)
    if rank == 0:
        f, s, w = (float(x) for x in dev)
        print(json.dumps({"config": 5, "n_gpus": world, "trajectories": N_TOTAL, "steps_per_trajectory": STEPS,
                          "filter_ms_max": f, "smoother_ms_max": s, "wall_ms_incl_h2d": w,
                          "filter_plus_smooth_steps_per_s": N_TOTAL * STEPS / ((f + s) * 1e-3), "ok": ok,
                          "marginals_to_host": ({"seconds_rank0": d2h[0], "bytes_rank0": d2h[1],
                                                 "note": "pnde_get_marginals into reused pinned host arrays (pnde_host_alloc): convert kernel + D2H copy"}
                                                if to_host else None)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
