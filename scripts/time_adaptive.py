"""C3 (examples/adaptivekernel.jl scaled to 1M residual blocks): kernel timings and a full LM solve on cuda:0 (development aid)."""
import sys, os, json, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); capi = pkg.capi
n_in, n_out = 333_334, 666_666
rng = np.random.default_rng(0)
data = 1.0 + np.concatenate([rng.standard_normal(n_in), rng.standard_normal(n_out) * 10.0])
ctx = capi.Context(0)
ctx.set_variables(capi.VAR_CONTAMGAUSS, pkg.ContaminatedGaussian(0.5, 5.0, 0.6).stored().reshape(1, 3), first_index=1)
ctx.set_variables(capi.VAR_SCALAR, np.zeros((1, 1)), first_index=2)
aos = np.zeros(len(data), dtype=pkg.ADAPTIVE_DTYPE)
aos["data"], aos["varind"] = data, 2
ctx.set_costs(capi.RES_ADAPTIVE_OFFSET, aos, capi.ROBUST_NONE, (), kernel_var=1)
ctx.linearize()
out = {name: round(ctx.time_kernels(getattr(capi, "TIME_" + name), reps=10, flush_l2=True), 4) for name in ["LINEARIZE", "COST"]}
t0 = time.perf_counter(); res = ctx.optimize(pkg.NLLSOptions().c()); dt = time.perf_counter() - t0
out.update(residuals=len(data), lm_iterations=int(res.niterations), optimize_ms=round(1e3 * dt, 3), bestcost=res.bestcost,
           residual_blocks_per_s_linearize=len(data) / (out["LINEARIZE"] * 1e-3))
print("adaptive", json.dumps(out))
