"""the Newton, Dogleg and gradient-descent iterators (src/iterators.jl:11-115,177-208) on the CUDA path against the
oracle (which is pinned to test/functional.jl:57-96 on the CPU).  The affine BA problem has a 12-DoF gauge freedom, so the undamped
Newton system the first two need is singular; four landmarks are fixed (an affine frame) through the `unfixed` mask to remove it."""
import numpy as np
import pytest

from helpers import cuda_context, oracle_problem

pytestmark = pytest.mark.gpu


def _problem(pkg, seed=3):
    rng = np.random.default_rng(seed)
    p = pkg.synthetic.create_bal_shaped(12, 300, 1500, rng, noise=0.005)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    mask = np.ones(p.ncam + p.npt, dtype=np.uint8)
    mask[p.ncam + np.array([0, 99, 199, 299])] = 0
    return p, mask


@pytest.mark.parametrize("name,it,iters,tol", [("newton", 0, 6, 1e-9), ("dogleg", 2, 12, 1e-9), ("gradientdescent", 3, 12, 1e-9), ("levenbergmarquardt", 1, 8, 1e-9)])
def test_iterator_trajectory(pkg, orc, name, it, iters, tol):
    p, mask = _problem(pkg)
    P = oracle_problem(orc, p)
    P.set_unfixed(mask)
    res_ref, tr_ref = P.optimize(orc.Options(maxiters=iters, iterator=it))
    ctx = cuda_context(pkg, p)
    ctx.set_unfixed(mask)
    ctx.lm_begin(pkg.NLLSOptions(maxiters=iters, iterator=it).c())
    conv, tr = 0, []
    while conv == 0:
        info = ctx.lm_iterate()
        tr.append((info.cost, int(info.ntries), info.lambda_))
        conv = ctx.lm_advance(info.cost, 0)
    res = ctx.lm_end()
    compared = 0
    for i, ((c, nt, lam), r) in enumerate(zip(tr, tr_ref)):
        # once successive costs agree to 1e-9 the accept / reject and termination decisions are rounding-level ties
        if i > 0 and abs(tr_ref[i - 1].cost - r.cost) <= 1e-9 * r.cost:
            break
        assert c == pytest.approx(r.cost, rel=tol), (name, i, c, r.cost)
        assert nt == r.ntries, (name, i, nt, r.ntries)                     # linear solves of the iteration
        if it != 0:
            assert lam == pytest.approx(r.lambda_, rel=1e-6), (name, i)    # lambda / trust radius / step size after the iteration
        compared += 1
    assert compared >= 2, compared
    assert res.bestcost == pytest.approx(res_ref.bestcost, rel=1e-8)
    assert ctx.cost(0) == res.bestcost
    ref = P.variables()
    cams = ctx.get_variables(pkg.capi.VAR_EUCLID6, p.ncam, 6)
    pts = ctx.get_variables(pkg.capi.VAR_EUCLID3, p.npt, 3)
    assert np.max(np.abs(cams.ravel() - ref[:6 * p.ncam])) <= 1e-7 and np.max(np.abs(pts.ravel() - ref[6 * p.ncam:])) <= 1e-7
    ctx.close()


def test_python_api_iterators(pkg):
    # test/functional.jl:57-96 style: every iterator through optimize!(problem, NLLSOptions(iterator=...))
    p, mask = _problem(pkg, seed=5)
    costs = {}
    for it in (pkg.newton, pkg.levenbergmarquardt, pkg.dogleg, pkg.gradientdescent):
        prob = pkg.NLLSProblem()
        prob.addvariables(p.cameras)
        prob.addvariables(p.points)
        aos = p.costs_aos()
        prob.addcosts(pkg.AffineReprojection, aos)
        res = pkg.optimize(prob, pkg.NLLSOptions(iterator=it, maxiters=200 if it == pkg.gradientdescent else 30), mask)
        assert res.bestcost <= res.startcost
        assert pkg.cost(prob) == res.bestcost
        costs[it] = res.bestcost
    assert costs[pkg.newton] == pytest.approx(costs[pkg.levenbergmarquardt], rel=1e-6)
    assert costs[pkg.dogleg] == pytest.approx(costs[pkg.levenbergmarquardt], rel=1e-6)


def test_lm_step_is_iterate_plus_advance(pkg):
    """nlls_lm_step (one outer iteration with the null callback, what nlls_optimize runs and bench.py times) must reproduce the split
    nlls_lm_iterate / nlls_lm_advance(cost, 0) loop bit for bit, and both must end where nlls_optimize ends."""
    p, mask = _problem(pkg, seed=5)
    opts = pkg.NLLSOptions(maxiters=7)
    runs = []
    for mode in ("split", "step", "optimize"):
        ctx = cuda_context(pkg, p, pkg.capi.ROBUST_HUBER, (0.05,))
        tr = []
        if mode == "optimize":
            res = ctx.optimize(opts.c())
        else:
            ctx.lm_begin(opts.c())
            conv = 0
            while conv == 0:
                if mode == "split":
                    info = ctx.lm_iterate()
                    conv = ctx.lm_advance(info.cost, 0)
                else:
                    info, conv = ctx.lm_step()
                tr.append((info.cost, int(info.ntries), info.lambda_, conv))
            res = ctx.lm_end()
        runs.append((tr, res.bestcost, res.niterations, res.termination, ctx.get_variables(pkg.capi.VAR_EUCLID6, p.ncam, 6)))
        ctx.close()
    a, b = np.array(runs[0][0]), np.array(runs[1][0])
    assert a.shape == b.shape and len(a) == runs[0][2]
    assert np.array_equal(a[:, 1], b[:, 1]) and np.array_equal(a[:, 3], b[:, 3])        # inner tries, termination words
    assert np.allclose(a[:, 0], b[:, 0], rtol=1e-10, atol=0)
    # lambda: while the cost still moves; once successive costs agree to 1e-9 the gain ratio q is a quotient of rounding errors
    moving = np.concatenate([[True], np.abs(np.diff(a[:, 0])) > 1e-9 * a[1:, 0]])
    assert np.allclose(a[moving, 2], b[moving, 2], rtol=1e-6, atol=0)
    for r in runs[1:]:
        assert abs(r[1] - runs[0][1]) <= 1e-10 * runs[0][1] and r[2] == runs[0][2] and r[3] == runs[0][3]
        assert np.allclose(r[4], runs[0][4], rtol=1e-8, atol=1e-10)
