"""Development aid: a few launches of the Schur phase on a synthetic workload (run under ncu): python scripts/prof_schur.py [workload] [reps]"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
pkg = load_package(); capi = pkg.capi
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
which = sys.argv[3] if len(sys.argv) > 3 else "SCHUR"
p = bench.make_problem(pkg, wl)
ctx = capi.Context(0)
ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos(), capi.ROBUST_HUBER, (bench.HUBER_WIDTH,))
ctx.lm_begin(pkg.NLLSOptions(maxiters=100, maxtime=1e5).c())
ctx.time_kernels(getattr(capi, "TIME_" + which), reps=1, flush_l2=False)   # first launch: lazy module load
print(wl, which, ctx.time_kernels(getattr(capi, "TIME_" + which), reps=reps, flush_l2=True))
