#!/usr/bin/env python
"""Ensemble-scale parity statistic of the adaptive path (SURVEY 8c protocol (iii)): the fraction of trajectories whose
(naccept, nreject) is identical to the reference arithmetic's (C restatement, oracle/pnde_ref.c), and the error of u(t1).

    python benchmarks/parity_ensemble.py [--n 10000]
    PNDE_LIB=odefilters.jl_b200/libpnde_pow.so python benchmarks/parity_ensemble.py     # pow-controller build

One JSON line per ensemble.  With the arbiter fixture (tests/golden/arbiter_config3.npz: the same draws through the
60-digit recursion) it also prints, for config 3, how far the kernel and the reference arithmetic are from the EXACT
recursion of the reference's algorithm.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ensembles import ENSEMBLES, count_parity  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    a = ap.parse_args()
    lib = os.environ.get("PNDE_LIB", "libpnde.so")
    for name in ENSEMBLES:
        stats, raw = count_parity(name, a.n)
        stats["lib"] = os.path.basename(lib)
        if name == "config3_vdp_ek1q5":
            g = np.load(os.path.join(ROOT, "tests", "golden", "arbiter_config3.npz"))
            k = len(g["index"])
            cg, ref = raw["gpu_counts"], raw["ref"]
            sc = np.abs(g["u1"]).max(axis=1)
            stats["arbiter"] = {
                "n": int(k),
                "exact_counts": [[int(x), int(y)] for x, y in zip(g["naccept"], g["nreject"])][:8],
                "gpu_counts": [[int(x), int(y)] for x, y in zip(cg["naccept"][:k], cg["nreject"][:k])][:8],
                "ref_counts": [[int(x), int(y)] for x, y in zip(ref["naccept"][:k], ref["nreject"][:k])][:8],
                "mean_abs_dnaccept_gpu_vs_exact": float(np.abs(cg["naccept"][:k] - g["naccept"]).mean()),
                "mean_abs_dnaccept_ref_vs_exact": float(np.abs(ref["naccept"][:k] - g["naccept"]).mean()),
                "mean_abs_dnreject_gpu_vs_exact": float(np.abs(cg["nreject"][:k] - g["nreject"]).mean()),
                "mean_abs_dnreject_ref_vs_exact": float(np.abs(ref["nreject"][:k] - g["nreject"]).mean()),
                "frac_identical_gpu_vs_exact": float(((cg["naccept"][:k] == g["naccept"]) & (cg["nreject"][:k] == g["nreject"])).mean()),
                "frac_identical_ref_vs_exact": float(((ref["naccept"][:k] == g["naccept"]) & (ref["nreject"][:k] == g["nreject"])).mean()),
                "max_rel_u1_gpu_vs_exact": float((np.abs(raw["gpu_mean"][:k, :2] - g["u1"]).max(axis=1) / sc).max()),
                "max_rel_u1_ref_vs_exact": float((np.abs(ref["mean"][:k, :2] - g["u1"]).max(axis=1) / sc).max()),
            }
        print(json.dumps(stats), flush=True)
