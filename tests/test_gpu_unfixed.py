"""`unfixed` masks (optimize!(problem, options, unfixed), src/optimize.jl:5-20, varflags dispatch src/cost.jl:36-51) and
optimizesingles! (src/optimize.jl:60-76,183-205; test/optimizeba.jl:61-62), CUDA path against the oracle."""
import numpy as np
import pytest

from helpers import cuda_context, oracle_problem, relerr

pytestmark = pytest.mark.gpu


def _ba(pkg, ncam, npt, prop, seed=1, pn=3e-3, cn=0.0):
    rng = np.random.default_rng(seed)
    p = pkg.synthetic.create_ba_problem(ncam, npt, prop, rng)
    pkg.synthetic.perturb_ba_problem(p, pn, cn, rng)
    return p


@pytest.mark.parametrize("shape", [(3, 5, 1.0), (10, 50, 0.3)])
def test_optimizesingles_points(pkg, orc, shape):
    # test/optimizeba.jl:52-62: perturb the landmarks only, optimise every landmark on its own -> cost < 1e-15
    p = _ba(pkg, *shape)
    P = oracle_problem(orc, p)
    idx = np.arange(p.ncam + 1, p.ncam + p.npt + 1)
    it_ref = P.optimizesingles(idx)
    assert P.cost() < 1e-15
    ctx = cuda_context(pkg, p)
    it = ctx.optimize_singles(pkg.capi.VAR_EUCLID3, pkg.NLLSOptions().c())
    assert ctx.cost(0) < 1e-15                                          # :62
    assert it == it_ref                                                  # same number of LM iterations, summed over the landmarks
    pts = ctx.get_variables(pkg.capi.VAR_EUCLID3, p.npt, 3)
    pts_ref = P.variables()[6 * p.ncam:].reshape(p.npt, 3)
    assert np.max(np.abs(pts - pts_ref)) <= 1e-9
    cams = ctx.get_variables(pkg.capi.VAR_EUCLID6, p.ncam, 6)
    assert np.array_equal(cams, p.cameras)                               # cameras untouched
    ctx.close()


def test_optimizesingles_robust_noisy(pkg, orc):
    # noisy measurements + Huber: per-landmark optima are not zero-cost; final variables and cost against the oracle
    rng = np.random.default_rng(4)
    p = pkg.synthetic.create_bal_shaped(30, 2000, 9000, rng, noise=0.01, outlier_frac=0.05)
    pkg.synthetic.perturb_ba_problem(p, 1e-2, 0.0, rng)
    kern = (1, 0.02, False, 1.0)
    P = oracle_problem(orc, p, kernel=kern)
    it_ref = P.optimizesingles(np.arange(p.ncam + 1, p.ncam + p.npt + 1))
    ctx = cuda_context(pkg, p, 1, (0.02,))
    it = ctx.optimize_singles(pkg.capi.VAR_EUCLID3, pkg.NLLSOptions().c())
    assert abs(it - it_ref) <= 0.01 * it_ref                             # (a rounding-level tie may cost one landmark one iteration)
    assert ctx.cost(0) == pytest.approx(P.cost(), rel=1e-10)
    pts = ctx.get_variables(pkg.capi.VAR_EUCLID3, p.npt, 3)
    pts_ref = P.variables()[6 * p.ncam:].reshape(p.npt, 3)
    assert np.max(np.abs(pts - pts_ref)) <= 1e-7
    ctx.close()


@pytest.mark.parametrize("which", ["points", "cameras", "mixed"])
def test_unfixed_mask_trajectory(pkg, orc, which):
    # optimize!(problem, options, unfixed): only the unfixed variables move; same LM trajectory as the oracle's masked system
    rng = np.random.default_rng(2)
    p = pkg.synthetic.create_bal_shaped(20, 600, 3000, rng, noise=0.01)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    n = p.ncam + p.npt
    mask = np.zeros(n, dtype=np.uint8)
    if which == "points":
        mask[p.ncam:] = 1
    elif which == "cameras":
        mask[:p.ncam] = 1
    else:
        mask[:] = rng.random(n) < 0.6
        mask[0] = 0
        mask[-1] = 1
    P = oracle_problem(orc, p)
    P.set_unfixed(mask)
    res_ref, tr_ref = P.optimize(orc.Options(maxiters=8))
    ctx = cuda_context(pkg, p)
    ctx.set_unfixed(mask)
    c0 = ctx.linearize()
    assert ctx.dof == P.dof
    g = ctx.gradient()
    P2 = oracle_problem(orc, p)
    P2.set_unfixed(mask)
    assert P2.linearize() == pytest.approx(c0, rel=1e-10)
    assert relerr(g, P2.grad()) <= 1e-12                                 # linsystem.b over the unfixed variables only
    ctx.lm_begin(pkg.NLLSOptions(maxiters=8).c())
    conv, tr = 0, []
    while conv == 0:
        info = ctx.lm_iterate()
        tr.append((info.cost, int(info.ntries)))
        conv = ctx.lm_advance(info.cost, 0)
    res = ctx.lm_end()
    compared = 0
    for i, ((c, nt), r) in enumerate(zip(tr, tr_ref)):
        if i > 0 and abs(tr_ref[i - 1].cost - r.cost) <= 1e-9 * r.cost:
            break                                                        # converged: the remaining decisions are rounding-level ties
        assert nt == r.ntries
        assert c == pytest.approx(r.cost, rel=1e-9)
        compared += 1
    assert compared >= 1   # (cameras fixed: the affine problem is linear in the landmarks and converges in one step)
    cams = ctx.get_variables(pkg.capi.VAR_EUCLID6, p.ncam, 6)
    pts = ctx.get_variables(pkg.capi.VAR_EUCLID3, p.npt, 3)
    fixed_c, fixed_p = mask[:p.ncam] == 0, mask[p.ncam:] == 0
    assert np.array_equal(cams[fixed_c], p.cameras[fixed_c]) and np.array_equal(pts[fixed_p], p.points[fixed_p])   # fixed variables: bitwise untouched
    ref = P.variables()
    assert np.max(np.abs(cams.ravel() - ref[:6 * p.ncam])) <= 1e-7 and np.max(np.abs(pts.ravel() - ref[6 * p.ncam:])) <= 1e-7
    assert res.bestcost == pytest.approx(res_ref.bestcost, rel=1e-8)
    with pytest.raises(pkg.capi.NLLSError):
        ctx.hessian_blocks()                                             # no reference-layout read-back under a mask
    ctx.set_unfixed(None)
    assert ctx.dof == 6 * p.ncam + 3 * p.npt
    ctx.close()


def test_python_api_unfixed_and_singles(pkg):
    rng = np.random.default_rng(1)
    p = pkg.synthetic.create_ba_problem(3, 5, 1.0, rng)
    pkg.synthetic.perturb_ba_problem(p, 3e-3, 0.0, rng)
    prob = pkg.NLLSProblem()
    for c in p.cameras:
        prob.addvariable(pkg.EuclideanVector(c))
    for x in p.points:
        prob.addvariable(pkg.EuclideanVector(x))
    for z, ci, pi in zip(p.z, p.cam_idx, p.pt_idx):
        prob.addcost(pkg.AffineReprojection(z, ci, pi))
    cams0 = [np.array(v) for v in prob.variables[:3]]
    mask = np.array([len(np.asarray(v)) == 3 for v in prob.variables])
    res = pkg.optimize(prob, pkg.NLLSOptions(), mask)                    # landmarks only, jointly
    assert res.bestcost < 1e-15 and pkg.cost(prob) == res.bestcost
    assert all(np.array_equal(np.asarray(a), b) for a, b in zip(prob.variables[:3], cams0))
    # and one at a time (test/optimizeba.jl:61-62)
    for i in range(3, 8):
        prob.variables[i] = pkg.EuclideanVector(np.asarray(prob.variables[i]) + rng.standard_normal(3) * 3e-3)
    pkg.optimizesingles(prob, pkg.NLLSOptions(), pkg.capi.VAR_EUCLID3)
    assert pkg.cost(prob) < 1e-15
