import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import load_package
from helpers import *
pkg = load_package()
from oracle import oracle as orc
import test_gpu_parity as T
for kname in ["none","huber","huber2o"]:
    ok, rid, kp = T.KERNELS[kname]
    p = T._bal(pkg, *pkg.synthetic.SHAPES["ladybug"], noise=0.01, outlier_frac=0.05 if kname != "none" else 0.0)
    res, tr, res_ref, tr_ref, ctx, P = T._compare_trajectories(pkg, orc, p, ok, rid, kp, maxiters=30)
    print(kname, len(tr), len(tr_ref), res.bestcost, res_ref.bestcost, hex(res.termination), hex(res_ref.termination))
    for i,(a,b) in enumerate(zip(tr,tr_ref)):
        print(i, "%.3e"%abs(a[0]/b.cost-1), a[1], b.ntries, "%.3e %.3e"%(a[2], b.lambda_), "%.15g"%b.cost)
