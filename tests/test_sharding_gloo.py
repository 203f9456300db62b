"""N > 1 host-side logic on CPU: points are sharded across ranks (gloo, world_size 2); every rank linearises only its own
residual blocks (here with the CPU oracle standing in for the device kernels), camera blocks / gradient / cost are combined
with an all-reduce, and the result must equal the unsharded linearisation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_package
    import bench
    from oracle import oracle as orc
    pkg = load_package()
    rng = np.random.default_rng(3)
    p = pkg.synthetic.create_bal_shaped(20, 600, 2600, rng, noise=0.01, outlier_frac=0.05)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    kern = (orc.RK_HUBER, 0.03, False, 1.0)
    pts_sel, obs_sel = bench.shard_by_point(p, rank, world)
    # every observation of an owned point is local, and nothing else
    pl = p.pt_idx - p.ncam - 1
    assert np.array_equal(np.isin(pl, pts_sel), obs_sel)
    # local problem: all cameras (replicated) + owned points, renumbered locally for the oracle
    remap = -np.ones(p.npt, dtype=np.int64)
    remap[pts_sel] = np.arange(len(pts_sel))
    P = orc.Problem()
    P.add_variables(orc.VT_EUCLID, p.cameras)
    P.add_variables(orc.VT_EUCLID, p.points[pts_sel])
    P.add_costs(orc.RT_AFFINE_BA, np.stack([p.cam_idx[obs_sel], p.ncam + 1 + remap[pl[obs_sel]]], 1), p.z[obs_sel], kernel=kern)
    c = P.linearize()
    # dense images: a camera without local observations has no diagonal block in the local BlockSparseMatrix, so the
    # block-sparse layouts of the shards differ from the full one; the dense images are directly comparable
    Hd = P.hess_dense()
    g = P.grad()
    nc = 6 * p.ncam
    cam = torch.from_numpy(np.concatenate([Hd[:nc, :nc].ravel(), g[:nc], [c]]))
    dist.all_reduce(cam)                      # camera blocks, camera gradient and the cost are sums over ranks
    counts = torch.tensor([len(pts_sel), int(obs_sel.sum()), int(pts_sel[0])])
    gathered = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts)
    Pf = orc.Problem()
    Pf.add_variables(orc.VT_EUCLID, p.cameras)
    Pf.add_variables(orc.VT_EUCLID, p.points)
    Pf.add_costs(orc.RT_AFFINE_BA, np.stack([p.cam_idx, p.pt_idx], 1), p.z, kernel=kern)
    cf = Pf.linearize()
    Hf, gf = Pf.hess_dense(), Pf.grad()
    # the point rows are rank-local: they must equal the corresponding rows of the full system without any communication
    r0, r1 = nc + 3 * int(pts_sel[0]), nc + 3 * (int(pts_sel[-1]) + 1)
    rows_ok = np.array_equal(Hd[nc:, :nc], Hf[r0:r1, :nc]) and np.array_equal(Hd[nc:, nc:], Hf[r0:r1, r0:r1]) and np.array_equal(g[nc:], gf[r0:r1])
    ok = torch.tensor([int(rows_ok)])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        res = {
            "cam_err": float(np.max(np.abs(cam.numpy()[:nc * nc].reshape(nc, nc) - Hf[:nc, :nc])) / np.max(np.abs(Hf[:nc, :nc]))),
            "gcam_err": float(np.max(np.abs(cam.numpy()[nc * nc:nc * nc + nc] - gf[:nc])) / np.max(np.abs(gf))),
            "cost_err": abs(float(cam[-1]) - cf) / cf,
            "rows_equal": bool(ok.item()),
            "npts": [int(x[0]) for x in gathered], "nobs": [int(x[1]) for x in gathered], "total": (p.npt, p.nobs),
        }
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res["cam_err"] < 1e-13 and res["gcam_err"] < 1e-13 and res["cost_err"] < 1e-13
    assert res["rows_equal"]
    assert sum(res["npts"]) == res["total"][0] and sum(res["nobs"]) == res["total"][1]
    assert abs(res["nobs"][0] - res["nobs"][1]) < 0.05 * res["total"][1]     # balanced by observation count


def test_shard_partition_properties():
    sys.path.insert(0, ROOT)
    import bench
    from conftest import load_package
    pkg = load_package()
    rng = np.random.default_rng(0)
    p = pkg.synthetic.create_bal_shaped(49, 7776, 31843, rng)
    for world in (1, 2, 4, 8):
        seen_pts = np.zeros(p.npt, dtype=int)
        seen_obs = np.zeros(p.nobs, dtype=int)
        for r in range(world):
            pts, obs = bench.shard_by_point(p, r, world)
            assert len(pts) > 0 and np.all(np.diff(pts) == 1)        # contiguous, non-empty
            seen_pts[pts] += 1
            seen_obs[obs] += 1
        assert np.all(seen_pts == 1) and np.all(seen_obs == 1)        # a partition: every block owned exactly once
