"""Synthetic problem generators (numpy only; host-side test/bench data, no solver code).

The Julia RNG stream is not reproducible outside Julia, so problems are regenerated from the reference's
*construction rules* with numpy's default_rng:

* create_ba_problem    — test/optimizeba.jl:6-36 (affine camera, banded visibility, noise-free measurements,
                         camera-major cost order); perturb_ba_problem — test/optimizeba.jl:38-47.
* create_bal_shaped    — BAL-*shaped* problems (Ladybug / Venice / Final sizes) built with the same rules:
                         point l has centre camera linspace(2, ncam-1, npt)[l] and is seen by its k_l nearest
                         cameras, k_l >= 2, sum k_l = nobs; measurements = projection + noise (+ outliers).
* create_adaptive_problem — examples/adaptivekernel.jl:20-30 / test/adaptivecost.jl:33-38.
"""
import numpy as np

CAM_OFFSET = np.array([1.0, 0.0, 0.0, 0.0, 1.0, 0.0])
LM_OFFSET = np.array([-0.5, -0.5, 10.0])

SHAPES = {  # (ncam, npt, nobs)  BASELINE.md §4
    "ladybug": (49, 7776, 31843),
    "venice": (1778, 993923, 5001946),
    "final": (13682, 4456117, 28987644),
}


def project_affine(cams, pts):
    """generatemeasurement(pose, X) = (pose[1:3].X, pose[4:6].X)   (test/optimizeba.jl:4)."""
    return np.stack([np.einsum("ij,ij->i", cams[:, 0:3], pts), np.einsum("ij,ij->i", cams[:, 3:6], pts)], axis=1)


class BAProblem:
    """Plain container: cameras (ncam, dc), points (npt, 3), costs in storage order.
    cam_idx / pt_idx are 1-based *global variable indices* (cameras first, then points), as varindices returns."""

    def __init__(self, cameras, points, cam_idx, pt_idx, z):
        self.cameras = np.ascontiguousarray(cameras, dtype=np.float64)
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int64)
        self.pt_idx = np.ascontiguousarray(pt_idx, dtype=np.int64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)

    @property
    def ncam(self):
        return self.cameras.shape[0]

    @property
    def npt(self):
        return self.points.shape[0]

    @property
    def nobs(self):
        return self.z.shape[0]

    def costs_aos(self):
        """The memory image of Vector{SimpleError2{2,Float64,EV6,EV3}}: 32 B per cost =
        measurement (2 x f64) then varind (2 x i64)   (src/residual.jl:4-7)."""
        aos = np.zeros(self.nobs, dtype=np.dtype([("z", "<f8", 2), ("varind", "<i8", 2)]))
        aos["z"] = self.z
        aos["varind"][:, 0] = self.cam_idx
        aos["varind"][:, 1] = self.pt_idx
        return aos

    def copy(self):
        return BAProblem(self.cameras.copy(), self.points.copy(), self.cam_idx, self.pt_idx, self.z)


def create_ba_problem(ncameras, nlandmarks, propvisible, rng):
    """test/optimizeba.jl:6-36."""
    cams = rng.standard_normal((ncameras, 6)) + CAM_OFFSET
    pts = rng.random((nlandmarks, 3)) + LM_OFFSET
    centres = np.linspace(2, ncameras - 1, nlandmarks)
    vis = np.abs(np.arange(1, ncameras + 1)[:, None] - centres[None, :])
    thresh = np.sort(vis.ravel())[int(np.ceil(vis.size * propvisible)) - 1]
    vis = vis <= thresh
    cam_l, lm_l = np.nonzero(vis)  # row-major nonzero == camera-major cost order (:24-32)
    z = project_affine(cams[cam_l], pts[lm_l])
    return BAProblem(cams, pts, cam_l + 1, lm_l + 1 + ncameras, z)


def perturb_ba_problem(problem, pointnoise, posenoise, rng):
    """test/optimizeba.jl:38-47."""
    problem.cameras = problem.cameras + rng.standard_normal(problem.cameras.shape) * posenoise
    problem.points = problem.points + rng.standard_normal(problem.points.shape) * pointnoise
    return problem


def create_bal_shaped(ncam, npt, nobs, rng, noise=0.01, outlier_frac=0.0, outlier_scale=50.0, camera_major=True):
    """BAL-shaped synthetic BA with the reference test's affine camera model."""
    assert nobs >= 2 * npt and ncam >= 2
    cams = rng.standard_normal((ncam, 6)) + CAM_OFFSET
    pts = rng.random((npt, 3)) + LM_OFFSET
    # track lengths k_l >= 2 with sum == nobs
    mean_extra = nobs / npt - 2.0
    k = 2 + rng.poisson(mean_extra, npt)
    k = np.minimum(k, ncam)
    diff = int(nobs - k.sum())
    while diff != 0:
        if diff > 0:
            cand = np.nonzero(k < ncam)[0]
            sel = rng.choice(cand, size=min(diff, cand.size), replace=False)
            k[sel] += 1
        else:
            cand = np.nonzero(k > 2)[0]
            sel = rng.choice(cand, size=min(-diff, cand.size), replace=False)
            k[sel] -= 1
        diff = int(nobs - k.sum())
    centres = np.linspace(2, ncam - 1, npt)
    start = np.clip(np.ceil(centres - k / 2.0).astype(np.int64), 1, ncam - k + 1)  # 1-based first camera of the window
    obs_start = np.concatenate([[0], np.cumsum(k)])
    pt_l = np.repeat(np.arange(npt, dtype=np.int64), k)
    cam_l = (np.arange(nobs, dtype=np.int64) - obs_start[pt_l]) + start[pt_l] - 1  # 0-based camera
    if camera_major:
        order = np.lexsort((pt_l, cam_l))
        pt_l, cam_l = pt_l[order], cam_l[order]
    z = project_affine(cams[cam_l], pts[pt_l])
    z += rng.standard_normal(z.shape) * noise
    if outlier_frac > 0:
        nout = int(round(outlier_frac * nobs))
        sel = rng.choice(nobs, size=nout, replace=False)
        z[sel] += rng.standard_normal((nout, 2)) * (noise * outlier_scale)
    return BAProblem(cams, pts, cam_l + 1, pt_l + 1 + ncam, z)


def create_scattered(ncam, npt, kmin, kmax, rng, noise=0.01):
    """Affine BA whose points see RANDOM camera subsets (no locality at all): every camera pair is coupled, so the reduced
    camera system is dense and consecutive points share no Schur blocks — the opposite corner from create_bal_shaped."""
    cams = rng.standard_normal((ncam, 6)) + CAM_OFFSET
    pts = rng.random((npt, 3)) + LM_OFFSET
    cam_l, pt_l = [], []
    for l in range(npt):
        k = int(rng.integers(kmin, kmax + 1))
        cs = np.sort(rng.choice(ncam, size=min(k, ncam), replace=False))
        cam_l.append(cs)
        pt_l.append(np.full(cs.size, l, dtype=np.int64))
    cam_l = np.concatenate(cam_l).astype(np.int64)
    pt_l = np.concatenate(pt_l)
    order = np.lexsort((pt_l, cam_l))  # camera-major cost order, like the reference test
    pt_l, cam_l = pt_l[order], cam_l[order]
    z = project_affine(cams[cam_l], pts[pt_l]) + rng.standard_normal((cam_l.size, 2)) * noise
    return BAProblem(cams, pts, cam_l + 1, pt_l + 1 + ncam, z)


def create_shape(name, rng, **kw):
    ncam, npt, nobs = SHAPES[name]
    return create_bal_shaped(ncam, npt, nobs, rng, **kw)


def create_adaptive_problem(ninliers, noutliers, rng, inliersigma=1.0, outliersigma=10.0, offset=1.0):
    """examples/adaptivekernel.jl:20-30: data = offset + [N(0, s_in) x ninliers ; N(0, s_out) x noutliers];
    start ContaminatedGaussian(0.5, 5.0, 0.6), mean 0."""
    data = offset + np.concatenate([rng.standard_normal(ninliers) * inliersigma, rng.standard_normal(noutliers) * outliersigma])
    return {"data": data, "start_kernel": (0.5, 5.0, 0.6), "start_mean": 0.0}


# ------------------------------------------------------------------------------------------------
# Pinhole / SO(3) cameras (repo-defined residual, BAL convention: P = R X + t, p = -P.xy / P.z, r = f (1 + k1 |p|^2 + k2 |p|^4) p - z)
# ------------------------------------------------------------------------------------------------
def so3_exp(w):
    """Rodrigues formula, vectorised: w (n, 3) -> R (n, 3, 3)."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=1)
    K = np.zeros((len(w), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -w[:, 2], w[:, 1]
    K[:, 1, 0], K[:, 1, 2] = w[:, 2], -w[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -w[:, 1], w[:, 0]
    small = th < 1e-5
    ths = np.where(small, 1.0, th)
    A = np.where(small, 1.0 - th ** 2 / 6.0, np.sin(ths) / ths)
    B = np.where(small, 0.5 - th ** 2 / 24.0, (1.0 - np.cos(ths)) / ths ** 2)
    return np.eye(3)[None] + A[:, None, None] * K + B[:, None, None] * (K @ K)


def pinhole_cameras(rod, t, f, k1, k2):
    """Stored form of NLLS_VAR_PINHOLE: R column-major (9), t (3), f, k1, k2."""
    R = so3_exp(rod)
    return np.concatenate([R.transpose(0, 2, 1).reshape(len(rod), 9), t, f[:, None], k1[:, None], k2[:, None]], axis=1)


def project_pinhole(cams, pts):
    R = cams[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    P = np.einsum("nij,nj->ni", R, pts) + cams[:, 9:12]
    pxy = -P[:, :2] / P[:, 2:3]
    n2 = np.sum(pxy * pxy, axis=1)
    s = cams[:, 12] * (1.0 + n2 * (cams[:, 13] + cams[:, 14] * n2))
    return s[:, None] * pxy


def create_bal_shaped_pinhole(ncam, npt, nobs, rng, noise=0.5, outlier_frac=0.0, outlier_scale=20.0):
    """The same visibility structure as create_bal_shaped with 9-DoF pinhole cameras (15 stored doubles)."""
    shape = create_bal_shaped(ncam, npt, nobs, rng, noise=0.0)
    cams = pinhole_cameras(rng.standard_normal((ncam, 3)) * 0.05, np.array([0.0, 0.0, -10.0]) + rng.standard_normal((ncam, 3)) * 0.1,
                           500.0 + 20.0 * rng.standard_normal(ncam), 1e-2 * rng.standard_normal(ncam), 1e-3 * rng.standard_normal(ncam))
    pts = rng.uniform(-1.0, 1.0, (npt, 3))
    cl, pl = shape.cam_idx - 1, shape.pt_idx - ncam - 1
    z = project_pinhole(cams[cl], pts[pl]) + rng.standard_normal((shape.nobs, 2)) * noise
    if outlier_frac > 0:
        nout = int(round(outlier_frac * shape.nobs))
        sel = rng.choice(shape.nobs, size=nout, replace=False)
        z[sel] += rng.standard_normal((nout, 2)) * (noise * outlier_scale)
    return BAProblem(cams, pts, shape.cam_idx, shape.pt_idx, z)


def perturb_pinhole_problem(problem, pointnoise, rotnoise, rng):
    """Landmarks += N(0, pointnoise); camera rotations <- Exp(N(0, rotnoise)) R  (the minimal update of NLLS_VAR_PINHOLE)."""
    problem.points = problem.points + rng.standard_normal(problem.points.shape) * pointnoise
    E = so3_exp(rng.standard_normal((problem.ncam, 3)) * rotnoise)
    R = problem.cameras[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    cams = problem.cameras.copy()
    cams[:, :9] = (E @ R).transpose(0, 2, 1).reshape(-1, 9)
    problem.cameras = cams
    return problem
