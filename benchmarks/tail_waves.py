"""Wave quantisation of small shards (strong scaling): the headline kernel at 125 k / 250 k / 500 k trajectories with
128-, 64- and 32-thread CTAs (PNDE_FILTER_BLOCK_RT), against the ideal n / 1e6 x the 1e6 time."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import odefilters_b200 as B
from ensembles import config2_inputs
n = int(sys.argv[1])
u0, p = config2_inputs(n)
prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
s = B.FilterSolver(prob, B.EK1(order=3, smooth=False), adaptive=False, dt=0.01, save_everystep=False)
s.upload(u0, p)
ms = []
for _ in range(4):
    s.run(); ms.append(s.last_run_ms()[0])
print(min(ms[1:]))
''' % (ROOT, ROOT)
base = None
for n in (1000000, 500000, 250000, 125000):
    row = {"n": n}
    for blk in ("128", "64", "32"):
        env = dict(os.environ, PNDE_FILTER_BLOCK_RT=blk)
        row[blk] = float(subprocess.run([sys.executable, "-c", code, str(n)], env=env, capture_output=True, text=True).stdout.strip() or "nan")
    if base is None:
        base = row["128"]
    row["ideal_ms"] = base * n / 1e6
    row["eff_128"], row["eff_64"], row["eff_32"] = (row["ideal_ms"] / row[b] for b in ("128", "64", "32"))
    print(json.dumps(row), flush=True)
