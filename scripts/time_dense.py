"""North-star item (c), the DENSE corner: a reduced camera system in which every camera pair is coupled (random visibility), so every
72 x 72 tile of S exists and the tile LDL' is a dense blocked factorisation on the FP64 tensor cores.  Prints the time of the reduced
solve, its algorithmic flops (n^3 / 3 for the factorisation + sweeps) and the fraction of the measured DMMA peak.
python scripts/time_dense.py [ncam] [npt]"""
import sys, os, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); capi = pkg.capi
ncam = int(sys.argv[1]) if len(sys.argv) > 1 else 480
npt = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
rng = np.random.default_rng(0)
p = pkg.synthetic.create_scattered(ncam, npt, 4, 8, rng, noise=0.01)
pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
ctx = capi.Context(0)
ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos(), capi.ROBUST_HUBER, (0.03,))
ctx.lm_begin(pkg.NLLSOptions(maxiters=100, maxtime=1e5).c())
tr = []
for _ in range(3):
    info, conv = ctx.lm_step(); tr.append(info.cost)
out = {"cameras": ncam, "points": npt, "observations": int(p.nobs), "n_reduced": 6 * ncam, "cost_trace": tr}
for name in ["SCHUR", "SOLVE_REDUCED", "TRY"]:
    out[name.lower() + "_ms"] = round(ctx.time_kernels(getattr(capi, "TIME_" + name), reps=5, flush_l2=True), 4)
fl = ctx.algorithmic_flops(capi.TIME_SOLVE_REDUCED)
out["reduced_solve_flops"] = fl
out["reduced_solve_tflops"] = fl / (out["solve_reduced_ms"] * 1e-3) / 1e12
out["dense_cholesky_flops_n3_over_3"] = (6 * ncam) ** 3 / 3
out["fraction_of_dmma_peak_35.5_tflops"] = out["reduced_solve_tflops"] / 35.48
print(json.dumps(out))
