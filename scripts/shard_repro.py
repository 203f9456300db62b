"""Debugging aid: run rank r's shard of an N-rank job on ONE GPU (no NCCL) — local kernels see exactly the rank's inputs.
   python scripts/shard_repro.py final 8 [rank ...]"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
pkg = load_package(); capi = pkg.capi
wl, world = sys.argv[1], int(sys.argv[2])
ranks = [int(a) for a in sys.argv[3:]] or list(range(world))
p = bench.make_problem(pkg, wl)
for r in ranks:
    pts_sel, obs_sel = bench.shard_by_point(p, r, world)
    ctx = capi.Context(0)
    ctx.set_variables(capi.VAR_EUCLID6, np.ascontiguousarray(p.cameras), first_index=1)
    ctx.set_variables(capi.VAR_EUCLID3, np.ascontiguousarray(p.points[pts_sel]), first_index=p.ncam + 1 + int(pts_sel[0]))
    ctx.set_costs(capi.RES_AFFINE_BA, p.costs_aos()[obs_sel], capi.ROBUST_HUBER, (bench.HUBER_WIDTH,))
    try:
        ctx.prepare()
        ctx.lm_begin(pkg.NLLSOptions(maxiters=10, maxtime=1e5).c())
        for it in range(2):
            info = ctx.lm_iterate(); ctx.lm_advance(info.cost, 0)
        print(f"rank {r}/{world}: points {len(pts_sel)} obs {int(obs_sel.sum())} cost {info.cost:.6f} ok", flush=True)
    except Exception as e:
        print(f"rank {r}/{world}: points {len(pts_sel)} obs {int(obs_sel.sum())} FAILED {e}", flush=True)
        break
    ctx.close()
