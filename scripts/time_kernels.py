"""Per-kernel-group timings on a synthetic workload (development aid): python scripts/time_kernels.py [workload]"""
import sys, os, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
pkg = load_package(); capi = pkg.capi
wl = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = bench.make_problem(pkg, wl)
ctx = capi.Context(0)
ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos(), capi.ROBUST_HUBER, (bench.HUBER_WIDTH,))
ctx.lm_begin(pkg.NLLSOptions(maxiters=100, maxtime=1e5).c())
for _ in range(2):
    info = ctx.lm_iterate(); ctx.lm_advance(info.cost, 0)
out = {}
for name in ["LINEARIZE", "LIN_POINT", "LIN_CAM", "COST", "SCHUR", "SOLVE_REDUCED", "BACKSUB", "TRY", "MEMSET_H", "LIN_POINT"]:
    out[name] = round(ctx.time_kernels(getattr(capi, "TIME_" + name), reps=5, flush_l2=True), 4)
print(wl, json.dumps(out))
