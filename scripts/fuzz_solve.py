"""Randomised parity sweep of the damped solve (development aid): many small / medium BA problems of varied shape — banded and
scattered visibility, short and very long tracks, one to thousands of Schur tiles, forced v2 / v4 / automatic Schur path —
each compared with the oracle's full-system solve.   python scripts/fuzz_solve.py [ncases] [seed]"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
from helpers import oracle_problem, cuda_context, relerr
pkg = load_package()
from oracle import oracle as orc


def random_problem(rng):
    kind = rng.choice(["banded", "scattered", "longtracks", "tiny"])
    if kind == "tiny":
        ncam, npt = int(rng.integers(2, 6)), int(rng.integers(1, 12))
        p = pkg.synthetic.create_scattered(ncam, npt, 2, ncam, rng)
    elif kind == "scattered":
        ncam, npt = int(rng.integers(5, 60)), int(rng.integers(20, 1500))
        p = pkg.synthetic.create_scattered(ncam, npt, 2, int(rng.integers(2, min(ncam, 14) + 1)), rng)
    elif kind == "longtracks":
        ncam, npt = int(rng.integers(40, 230)), int(rng.integers(5, 300))
        p = pkg.synthetic.create_scattered(ncam, npt, max(2, ncam // 3), ncam, rng)
    else:
        ncam, npt = int(rng.integers(3, 400)), int(rng.integers(10, 30000))
        nobs = int(npt * rng.uniform(2.0, min(8.0, ncam)))
        p = pkg.synthetic.create_bal_shaped(ncam, npt, max(nobs, 2 * npt), rng, noise=0.01, outlier_frac=0.03)
    pkg.synthetic.perturb_ba_problem(p, 1e-3, 1e-3, rng)
    return kind, p


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    worst = 0.0
    for case in range(ncases):
        kind, p = random_problem(rng)
        robust = int(rng.integers(0, 3))
        ok = [(0, 0.0, False, 1.0), (1, 0.02, False, 1.0), (2, 0.02, False, 1.0)][robust]
        kp = () if robust == 0 else (0.02,)
        lam = float(10.0 ** rng.uniform(-4, 1))
        P = oracle_problem(orc, p, kernel=ok)
        c_ref = P.linearize()
        x_ref = P.solve(lam)
        for schur in ("v2", "v4", None):
            if schur: os.environ["NLLS_B200_SCHUR"] = schur
            else: os.environ.pop("NLLS_B200_SCHUR", None)
            try:
                ctx = cuda_context(pkg, p, robust, kp)
                c = ctx.linearize()
                Hd = ctx.hessian_blocks()
                eh = relerr(Hd, P.hess_data()) if P.is_sparse else 0.0   # (dense oracle systems are compared in tests/test_gpu_parity.py)
                ctx.solve(lam)
                ex = relerr(ctx.step(), x_ref)
                ctx.close()
            except Exception as e:
                print(f"case {case} {kind} cams {p.ncam} pts {p.npt} obs {p.nobs} schur {schur}: EXCEPTION {e}", flush=True)
                continue
            worst = max(worst, ex)
            flag = "" if (ex <= 1e-8 and eh <= 1e-12 and abs(c - c_ref) <= 1e-10 * abs(c_ref)) else "   <-- MISMATCH"
            print(f"case {case} {kind} cams {p.ncam} pts {p.npt} obs {p.nobs} kmax {np.bincount(p.pt_idx - p.ncam - 1).max()} robust {robust} lam {lam:.2g} "
                  f"schur {schur}: H {eh:.1e} x {ex:.1e}{flag}", flush=True)
    os.environ.pop("NLLS_B200_SCHUR", None)
    print("worst step error", worst)


if __name__ == "__main__":
    main()
