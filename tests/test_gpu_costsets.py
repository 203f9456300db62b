"""Several residual types in one problem: the reference sums cost, gradient and Hessian over all cost types (src/cost.jl:54,
src/VectorRepo.jl:64-69; test/functional.jl:14-24 registers two).  On the CUDA path the types must share one registered kernel
family (same residual struct) and may differ in their robustkernel()."""
import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu

TOL_H = 1e-12
TOL_COST = 1e-10


def _split_problem(pkg, seed=2):
    rng = np.random.default_rng(seed)
    p = pkg.synthetic.create_bal_shaped(60, 5000, 24000, rng, noise=0.01, outlier_frac=0.05)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    which = rng.integers(0, 3, p.nobs)                       # cost i belongs to residual type which[i]
    return p, which


def test_three_cost_sets_match_oracle(pkg, orc):
    capi = pkg.capi
    p, which = _split_problem(pkg)
    kernels_o = [(orc.RK_HUBER, 0.03, False, 1.0), (orc.RK_NONE, 0.0, False, 1.0), (orc.RK_GEMANMCCLURE, 0.05, True, 2.0)]
    kernels_c = [(capi.ROBUST_HUBER, (0.03,)), (capi.ROBUST_NONE, ()), (capi.ROBUST_GEMANMCCLURE | capi.ROBUST_SCALED, (0.05, 2.0))]
    P = orc.Problem()
    P.add_variables(orc.VT_EUCLID, p.cameras)
    P.add_variables(orc.VT_EUCLID, p.points)
    vi = np.stack([p.cam_idx, p.pt_idx], 1)
    for s in range(3):
        P.add_costs(orc.RT_AFFINE_BA, vi[which == s], p.z[which == s], kernel=kernels_o[s])
    c_ref = P.linearize()
    lam = 1e-3
    x_ref = P.solve(lam)
    ctx = capi.Context(0)
    ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
    aos = p.costs_aos()
    ctx.set_costs(capi.RES_AFFINE_BA, aos[which == 0], *kernels_c[0])
    for s in (1, 2):
        ctx.add_costs(capi.RES_AFFINE_BA, aos[which == s], *kernels_c[s])
    c = ctx.linearize()
    assert abs(c - c_ref) <= TOL_COST * abs(c_ref)
    assert relerr(ctx.gradient(), P.grad()) <= TOL_H
    assert relerr(ctx.hessian_blocks(), P.hess_data()) <= TOL_H
    assert abs(ctx.cost(0) - c_ref) <= TOL_COST * abs(c_ref)          # the camera-major cost pass reads the set ids in its own order
    ctx.solve(lam)
    assert relerr(ctx.step(), x_ref) <= 1e-9
    res_o, _ = P.optimize(orc.Options(maxiters=5, maxtime=1e5))
    res = ctx.optimize(pkg.NLLSOptions(maxiters=5, maxtime=1e5).c())
    assert int(res.niterations) == int(res_o.niterations) and int(res.costcomputations) == int(res_o.costcomputations)
    assert abs(res.bestcost - res_o.bestcost) <= TOL_COST * abs(res_o.bestcost)
    # a different kernel family in the same problem has no kernel
    with pytest.raises(capi.NLLSError):
        ctx.add_costs(capi.RES_PINHOLE_BA, aos[:4], capi.ROBUST_NONE, ())
    ctx.close()


def test_two_residual_types_through_the_problem_api(pkg, orc):
    """NLLSProblem / addcost! with two residual types (one robustified, one not), optimize!: same result as the oracle."""
    class RobustReprojection(pkg.AffineReprojection):
        robustkernel = pkg.HuberKernel(0.03)

    rng = np.random.default_rng(4)
    p = pkg.synthetic.create_ba_problem(10, 50, 0.3, rng)
    p.z = p.z + rng.standard_normal(p.z.shape) * 0.01
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    prob = pkg.NLLSProblem()
    for cam in p.cameras:
        prob.addvariable(pkg.EuclideanVector(cam))
    for X in p.points:
        prob.addvariable(pkg.EuclideanVector(X))
    robust = np.arange(p.nobs) % 2 == 0
    for i in range(p.nobs):
        T = RobustReprojection if robust[i] else pkg.AffineReprojection
        prob.addcost(T(p.z[i], p.cam_idx[i], p.pt_idx[i]))
    P = orc.Problem()
    P.add_variables(orc.VT_EUCLID, p.cameras)
    P.add_variables(orc.VT_EUCLID, p.points)
    vi = np.stack([p.cam_idx, p.pt_idx], 1)
    P.add_costs(orc.RT_AFFINE_BA, vi[robust], p.z[robust], kernel=(orc.RK_HUBER, 0.03, False, 1.0))
    P.add_costs(orc.RT_AFFINE_BA, vi[~robust], p.z[~robust])
    c_ref = P.cost()
    assert abs(pkg.cost(prob) - c_ref) <= TOL_COST * abs(c_ref)
    res_o, _ = P.optimize(orc.Options(maxiters=6, maxtime=1e5))
    res = pkg.optimize(prob, pkg.NLLSOptions(maxiters=6, maxtime=1e5))
    assert res.niterations == res_o.niterations
    assert abs(res.bestcost - res_o.bestcost) <= 1e-9 * abs(res_o.bestcost) + 1e-18
