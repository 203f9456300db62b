"""Irregular points on the CUDA path against the oracle: tracks longer than any tile of the tile kernels / Schur plans holds, and
several costs on one (camera, point) pair.  The reference accepts both — updatesymA! / updateb! just accumulate
(src/linearsystem.jl:132-175) — so the drop-in must too (missing 7 / next 9)."""
import os

import numpy as np
import pytest

from helpers import cuda_context, oracle_problem, relerr

pytestmark = pytest.mark.gpu

TOL_H = 1e-12
TOL_COST = 1e-10


def _problem(pkg, ncam, npt, nobs, long_tracks, dup_frac, seed=5):
    """BAL-shaped affine problem; `long_tracks` points additionally see a run of that many cameras; a share `dup_frac` of the costs
    is duplicated with an independent measurement (two costs on one (camera, point) pair)."""
    rng = np.random.default_rng(seed)
    S = pkg.synthetic
    p = S.create_bal_shaped(ncam, npt, nobs, rng, noise=0.01, outlier_frac=0.02)
    cam, pt, z = [p.cam_idx], [p.pt_idx], [p.z]
    have = set(zip(p.cam_idx.tolist(), p.pt_idx.tolist()))
    for i, k in enumerate(long_tracks):
        l = int((i + 1) * npt / (len(long_tracks) + 1))                       # spread over the point range
        c0 = int(rng.integers(0, ncam - k + 1))
        cs = np.array([c for c in range(c0, c0 + k) if (c + 1, l + 1 + ncam) not in have], dtype=np.int64)
        zz = S.project_affine(p.cameras[cs], np.repeat(p.points[l][None], cs.size, 0)) + rng.standard_normal((cs.size, 2)) * 0.01
        cam.append(cs + 1); pt.append(np.full(cs.size, l + 1 + ncam, dtype=np.int64)); z.append(zz)
    cam, pt, z = np.concatenate(cam), np.concatenate(pt), np.concatenate(z)
    if dup_frac > 0:
        sel = rng.choice(cam.size, size=max(1, int(dup_frac * cam.size)), replace=False)
        zz = z[sel] + rng.standard_normal((sel.size, 2)) * 0.01
        cam, pt, z = np.concatenate([cam, cam[sel]]), np.concatenate([pt, pt[sel]]), np.concatenate([z, zz])
    order = np.lexsort((pt, cam))                                             # camera-major cost order, like the reference test
    q = S.BAProblem(p.cameras, p.points, cam[order], pt[order], z[order])
    S.perturb_ba_problem(q, 1e-3, 1e-3, rng)
    return q


CASES = {
    "long": dict(long_tracks=(240, 300, 333, 257), dup_frac=0.0),            # > 232 (Schur tiles) and > 256 (the largest point tile)
    "dups": dict(long_tracks=(), dup_frac=0.01),
    "both": dict(long_tracks=(300, 90), dup_frac=0.01),
}


@pytest.mark.parametrize("schur", ["auto", "v2", "v4", "v5"])
@pytest.mark.parametrize("case", list(CASES))
def test_irregular_points_linearize_solve_and_lm(pkg, orc, case, schur):
    p = _problem(pkg, 340, 4000, 20000, **CASES[case])
    kern = (orc.RK_HUBER, 0.05, False, 1.0)
    P = oracle_problem(orc, p, kernel=kern)
    c_ref = P.linearize()
    g_ref = P.grad()
    lam = 1e-3
    x_ref = P.solve(lam)
    if schur != "auto":
        os.environ["NLLS_B200_SCHUR"] = schur
    try:
        ctx = cuda_context(pkg, p, pkg.capi.ROBUST_HUBER, (0.05,))
        c = ctx.linearize()
        assert abs(c - c_ref) <= TOL_COST * abs(c_ref)
        assert relerr(ctx.gradient(), g_ref) <= TOL_H
        if CASES[case]["dup_frac"] == 0.0:                                   # (one block per cost on the device: no reference-layout read-back with duplicates)
            assert relerr(ctx.hessian_blocks(), P.hess_data()) <= TOL_H
        else:
            with pytest.raises(pkg.capi.NLLSError):
                ctx.hessian_blocks()
        ctx.solve(lam)
        assert relerr(ctx.step(), x_ref) <= 1e-9
        # a few LM iterations: same accept / reject sequence, costs to 1e-10
        res_o, tr_o = P.optimize(orc.Options(maxiters=4, maxtime=1e5))
        res = ctx.optimize(pkg.NLLSOptions(maxiters=4, maxtime=1e5).c())
        assert int(res.niterations) == int(res_o.niterations)
        assert int(res.costcomputations) == int(res_o.costcomputations)
        assert abs(res.bestcost - res_o.bestcost) <= TOL_COST * abs(res_o.bestcost)
        ctx.close()
    finally:
        os.environ.pop("NLLS_B200_SCHUR", None)


def test_irregular_points_pinhole(pkg, orc):
    """SO(3) / pinhole cameras: the Schur tiles hold 128 observations, tracks of 150 and 260 are irregular there."""
    rng = np.random.default_rng(9)
    S = pkg.synthetic
    base = S.create_bal_shaped_pinhole(300, 3000, 15000, rng, noise=0.3)
    cam, pt, z = [base.cam_idx], [base.pt_idx], [base.z]
    have = set(zip(base.cam_idx.tolist(), base.pt_idx.tolist()))
    for l, k in ((700, 150), (2100, 260)):
        cs = np.array([c for c in range(10, 10 + k) if (c + 1, l + 1 + base.ncam) not in have], dtype=np.int64)
        zz = S.project_pinhole(base.cameras[cs], np.repeat(base.points[l][None], cs.size, 0)) + rng.standard_normal((cs.size, 2)) * 0.3
        cam.append(cs + 1); pt.append(np.full(cs.size, l + 1 + base.ncam, dtype=np.int64)); z.append(zz)
    cam, pt, z = np.concatenate(cam), np.concatenate(pt), np.concatenate(z)
    order = np.lexsort((pt, cam))
    p = S.BAProblem(base.cameras, base.points, cam[order], pt[order], z[order])
    S.perturb_pinhole_problem(p, 1e-3, 1e-4, rng)
    P = orc.Problem()
    P.add_variables(orc.VT_PINHOLE, p.cameras)
    P.add_variables(orc.VT_EUCLID, p.points)
    P.add_costs(orc.RT_PINHOLE_BA, np.stack([p.cam_idx, p.pt_idx], 1), p.z, kernel=(orc.RK_HUBER, 1.5, False, 1.0))
    c_ref = P.linearize()
    lam = 1e-2
    x_ref = P.solve(lam)
    capi = pkg.capi
    ctx = capi.Context(0)
    ctx.set_variables(capi.VAR_PINHOLE, p.cameras, first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
    ctx.set_costs(capi.RES_PINHOLE_BA, p.costs_aos(), capi.ROBUST_HUBER, (1.5,))
    c = ctx.linearize()
    assert abs(c - c_ref) <= TOL_COST * abs(c_ref)
    assert relerr(ctx.gradient(), P.grad()) <= 1e-11
    ctx.solve(lam)
    assert relerr(ctx.step(), x_ref) <= 1e-9
    ctx.close()
