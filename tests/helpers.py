"""Shared helpers for the parity tests: build the oracle problem and the CUDA context from one synthetic BAProblem."""
import numpy as np


def oracle_problem(orc, p, kernel=None, order="cams_first"):
    """kernel: (kind, width, scaled, height) for the oracle."""
    P = orc.Problem()
    if order == "cams_first":
        P.add_variables(orc.VT_EUCLID, p.cameras)
        P.add_variables(orc.VT_EUCLID, p.points)
        vi = np.stack([p.cam_idx, p.pt_idx], 1)
    else:
        raise ValueError(order)
    P.add_costs(orc.RT_AFFINE_BA, vi, p.z, kernel=kernel or (orc.RK_NONE, 0.0, False, 1.0))
    return P


def cuda_context(pkg, p, robust=0, kparams=(), device=0):
    capi = pkg.capi
    ctx = capi.Context(device)
    ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, p.points, first_index=p.ncam + 1)
    ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos(), robust, kparams)
    return ctx


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def blockwise_relerr(a, b, starts, sizes):
    """max over blocks of |a_blk - b_blk|_max / |b_blk|_max  (per-block relative error, blocks with zero norm use abs)."""
    worst = 0.0
    for s, n in zip(starts, sizes):
        x, y = a[s:s + n], b[s:s + n]
        den = np.max(np.abs(y))
        e = np.max(np.abs(x - y))
        worst = max(worst, e / den if den > 0 else e)
    return worst


def densify(ctx, blocksizes):
    """Dense symmetric image of the library's block-sparse Hessian using nlls_get_hessian_index."""
    data = ctx.hessian_blocks()
    rb, cb, st = ctx.hessian_index()
    off = np.concatenate([[0], np.cumsum(blocksizes)])
    n = off[-1]
    M = np.zeros((n, n))
    for r, c, s in zip(rb, cb, st):
        nr, nc = blocksizes[r - 1], blocksizes[c - 1]
        blk = data[s - 1:s - 1 + nr * nc].reshape(nc, nr).T
        M[off[r - 1]:off[r - 1] + nr, off[c - 1]:off[c - 1] + nc] = blk
        if r != c:
            M[off[c - 1]:off[c - 1] + nc, off[r - 1]:off[r - 1] + nr] = blk.T
    return M
