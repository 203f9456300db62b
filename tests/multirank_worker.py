"""Worker of tests/test_gpu_multirank.py (launched with torch.distributed.run, one rank per GPU): runs K LM iterations of a
BAL-shaped problem sharded by point through the C ABI + NCCL and lets rank 0 write the trajectory and the final variables."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def build_problem(pkg, shape, seed=0):
    rng = np.random.default_rng(seed)
    if shape in pkg.synthetic.SHAPES:
        dims = pkg.synthetic.SHAPES[shape]
    else:
        dims = tuple(int(v) for v in shape.split(","))
    p = pkg.synthetic.create_bal_shaped(*dims, rng, noise=bench.NOISE, outlier_frac=bench.OUTLIERS)
    pkg.synthetic.perturb_ba_problem(p, bench.PERTURB, bench.PERTURB, rng)
    return p


def run(pkg, p, rank, world, device, iters, maxtime, comm=None, terminate_at=None):
    capi = pkg.capi
    pts_sel, obs_sel = bench.shard_by_point(p, rank, world)
    ctx = capi.Context(device)
    if comm is not None:
        ctx.comm_init(rank, world, comm)
    ctx.set_variables(capi.VAR_EUCLID6, p.cameras, first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, p.points[pts_sel], first_index=p.ncam + 1 + int(pts_sel[0]))
    ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos()[obs_sel], capi.ROBUST_HUBER, (bench.HUBER_WIDTH,))
    ctx.lm_begin(pkg.NLLSOptions(maxiters=iters, maxtime=maxtime).c())
    trace, conv = [], 0
    while conv == 0:
        info = ctx.lm_iterate()
        term = 0
        if terminate_at is not None and len(trace) + 1 == terminate_at[0]:
            term = terminate_at[1]
        conv = ctx.lm_advance(info.cost, term)
        trace.append((info.cost, int(info.ntries), info.lambda_, conv))
    res = ctx.lm_end()
    cams = ctx.get_variables(capi.VAR_EUCLID6, p.ncam, 6)
    pts = ctx.get_variables(capi.VAR_EUCLID3, len(pts_sel), 3)
    ctx.close()
    return np.array(trace), res, cams, pts, pts_sel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="ladybug")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="trajectory", choices=["trajectory", "maxtime", "callback"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    uid = [pkg.capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    p = build_problem(pkg, args.shape)
    maxtime, term = 1e5, None
    if args.mode == "maxtime":
        maxtime = 0.0 if rank == world - 1 else 1e5      # only ONE rank's clock says "time is up" (ranks must still agree)
    if args.mode == "callback":
        term = (2, 5 if rank == world - 1 else 0)        # only one rank's callback asks to stop, at iteration 2
    trace, res, cams, pts, pts_sel = run(pkg, p, rank, world, local, args.iters, maxtime, uid[0], term)
    gathered = [None] * world
    dist.all_gather_object(gathered, (trace, res.termination, res.niterations, res.bestcost, cams, pts, int(pts_sel[0])))
    if rank == 0:
        allpts = np.zeros((p.npt, 3))
        for g in gathered:
            allpts[g[6]:g[6] + len(g[5])] = g[5]
        np.savez(args.out, trace=gathered[0][0], traces=np.stack([g[0] for g in gathered]), termination=np.array([g[1] for g in gathered]),
                 niterations=np.array([g[2] for g in gathered]), bestcost=np.array([g[3] for g in gathered]),
                 cams=np.stack([g[4] for g in gathered]), points=allpts)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
