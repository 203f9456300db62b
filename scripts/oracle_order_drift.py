"""R18 evidence: how far do two EXACT solvers of the same damped system drift apart over an LM run?

Runs the oracle (CPU restatement of the reference) twice on the Ladybug-shaped problem of tests/test_gpu_parity.py with two
different elimination orders of its sparse LDL' (oracle.Problem.set_elimination_order) and prints, per LM iteration, the two
costs, the inner-try counts, lambda, and the relative difference; then the relative difference of the final costs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import oracle as orc  # noqa: E402

KERNELS = {"none": (0, 0.0, False, 1.0), "huber": (1, 0.02, False, 1.0), "huber2o": (2, 0.02, False, 1.0)}


def run(kname, order, maxiters=30, seed=0):
    pkg = load_package()
    rng = np.random.default_rng(seed)
    p = pkg.synthetic.create_bal_shaped(*pkg.synthetic.SHAPES["ladybug"], rng, noise=0.01, outlier_frac=0.05 if kname != "none" else 0.0)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    P = orc.Problem()
    P.set_elimination_order(order)
    P.add_variables(orc.VT_EUCLID, p.cameras)
    P.add_variables(orc.VT_EUCLID, p.points)
    P.add_costs(orc.RT_AFFINE_BA, np.stack([p.cam_idx, p.pt_idx], 1), p.z, kernel=KERNELS[kname])
    res, tr = P.optimize(orc.Options(maxiters=maxiters))
    return res, tr


if __name__ == "__main__":
    out = {}
    for kname in KERNELS:
        r0, t0 = run(kname, 0)
        r1, t1 = run(kname, 1)
        rows = []
        for a, b in zip(t0, t1):
            rows.append({"cost0": a.cost, "cost1": b.cost, "rel": abs(a.cost - b.cost) / abs(a.cost), "tries0": a.ntries, "tries1": b.ntries, "lambda0": a.lambda_})
        out[kname] = {"final_rel": abs(r0.bestcost - r1.bestcost) / abs(r0.bestcost), "niter": (r0.niterations, r1.niterations), "rows": rows}
        print(f"== {kname}: final cost {r0.bestcost:.15e} vs {r1.bestcost:.15e}  rel {out[kname]['final_rel']:.3e}  iterations {r0.niterations}/{r1.niterations}")
        for i, r in enumerate(rows):
            print(f"  it {i + 1:2d}  rel {r['rel']:.2e}  tries {r['tries0']}/{r['tries1']}  lambda {r['lambda0']:.3e}")
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
