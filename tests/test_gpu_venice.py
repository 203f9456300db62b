"""C4 — the configuration the headline number is quoted on (Venice-shape 1778 / 993 923 / 5 001 946, Huber(0.03), bench.py's own
generator and seed) — compared with the CPU oracle at full size: one linearisation (every Hessian block, gradient, cost), one
damped solve, and three LM iterations (same inner-try counts, per-iteration cost, lambda).  The oracle needs ~4 s per
linearisation and ~10 s per full-system LDL' solve at this size, so the whole test is about two minutes of CPU time."""
import numpy as np
import pytest

import bench
from helpers import relerr

pytestmark = pytest.mark.gpu


def _blockwise_relerr_ba(H, H_ref, ncam, k, dc=6):
    """max over ALL blocks of the reference BSM (cameras-first BA layout) of |a - b|_max / |b|_max, vectorised."""
    wb = 3 * dc
    nobs = int(k.sum())
    obs_start = np.concatenate([[0], np.cumsum(k)])
    npt = len(k)
    hB = dc * dc * ncam
    cam_starts = np.arange(ncam, dtype=np.int64) * (dc * dc)
    w_starts = hB + wb * np.arange(nobs, dtype=np.int64) + 9 * np.repeat(np.arange(npt, dtype=np.int64), k)
    v_starts = hB + wb * obs_start[1:].astype(np.int64) + 9 * np.arange(npt, dtype=np.int64)
    starts = np.sort(np.concatenate([cam_starts, w_starts, v_starts]))
    assert starts[-1] + 9 == len(H_ref)
    d = np.maximum.reduceat(np.abs(H - H_ref), starts)
    a = np.maximum.reduceat(np.abs(H_ref), starts)
    assert np.all(a > 0)
    return float(np.max(d / a)), len(starts)


def test_venice_oracle_parity(pkg, orc):
    p = bench.make_problem(pkg, "venice")
    capi = pkg.capi
    P = bench.oracle_problem(p, orc)
    c_ref = P.linearize()
    ctx = capi.Context(0)
    ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
    ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos(), capi.ROBUST_HUBER, (bench.HUBER_WIDTH,))
    c = ctx.linearize()
    assert abs(c - c_ref) <= 1e-10 * abs(c_ref), (c, c_ref)
    H, H_ref = ctx.hessian_blocks(), P.hess_data()
    assert H.shape == H_ref.shape
    k = np.bincount(p.pt_idx - p.ncam - 1, minlength=p.npt)
    worst, nblocks = _blockwise_relerr_ba(H, H_ref, p.ncam, k)
    assert nblocks == p.ncam + p.npt + p.nobs
    assert worst <= 1e-12, worst                                      # every one of the 5 997 647 blocks
    del H
    g, g_ref = ctx.gradient(), P.grad()
    assert relerr(g, g_ref) <= 1e-12
    assert relerr(g[:6 * p.ncam], g_ref[:6 * p.ncam]) <= 1e-12        # camera part (sums of ~2 800 terms each) on its own scale
    # one damped solve at the first LM damping value (initlambda: 1e-6 max |H_ii|, src/iterators.jl:131-137)
    diag_c = H_ref[:36 * p.ncam].reshape(p.ncam, 6, 6)[:, np.arange(6), np.arange(6)]
    lam = 1e-6 * max(float(np.max(np.abs(diag_c))), 0.0)
    x_ref = P.solve(lam)
    ctx.solve(lam)
    assert relerr(ctx.step(), x_ref) <= 1e-9
    del H_ref
    # three LM iterations from the same start: same accept/reject sequence, cost <= 1e-10, lambda
    res_ref, tr_ref = P.optimize(orc.Options(maxiters=3, maxtime=1e5))
    ctx.lm_begin(pkg.NLLSOptions(maxiters=3, maxtime=1e5).c())
    conv, tr = 0, []
    while conv == 0:
        info = ctx.lm_iterate()
        tr.append((info.cost, int(info.ntries), info.lambda_))
        conv = ctx.lm_advance(info.cost, 0)
    res = ctx.lm_end()
    assert len(tr) == len(tr_ref) == 3
    for (cst, nt, lm), r in zip(tr, tr_ref):
        assert nt == r.ntries
        assert cst == pytest.approx(r.cost, rel=1e-10)
        assert lm == pytest.approx(r.lambda_, rel=1e-6)
    assert res.bestcost == pytest.approx(res_ref.bestcost, rel=1e-10)
    assert res.termination == res_ref.termination == (1 << 8)
    assert ctx.cost(0) == res.bestcost
    ctx.close()


def test_bench_arms_agree_on_cost_trace(pkg):
    """bench.py's two arms (the CUDA path and --impl reference = the oracle) run the same LM trajectory on the same generated
    problem: their printed cost traces must agree to 1e-10 per iteration (Ladybug-shape, Huber; 8 iterations — the span over which
    the oracle agrees with itself under a second elimination order, profiles/r2_oracle_order_drift.json)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for impl in ("ours", "reference"):
        cmd = [sys.executable, os.path.join(root, "bench.py"), "--workload", "ladybug", "--steps", "5", "--warmup", "3", "--impl", impl, "--no-cpu-baseline"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[impl] = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = outs["ours"]["cost_trace"], outs["reference"]["cost_trace"]
    assert len(a) == len(b) == 8
    for x, y in zip(a, b):
        assert x == pytest.approx(y, rel=1e-10)
    assert outs["ours"]["config"] == outs["reference"]["config"]
