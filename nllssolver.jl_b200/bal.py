"""Reader / writer for the "Bundle Adjustment in the Large" text format (Agarwal et al.; grail.cs.washington.edu/projects/bal):

    <num_cameras> <num_points> <num_observations>
    <camera_index> <point_index> <x> <y>            x num_observations   (0-based indices)
    <9 camera parameters, one per line>             x num_cameras        Rodrigues vector (3), translation (3), f, k1, k2
    <3 point coordinates, one per line>             x num_points

The camera model is the one NLLS_RES_PINHOLE_BA implements (P = R X + t, p = -P.xy / P.z, r = f (1 + k1 |p|^2 + k2 |p|^4) p - z).
Host-side data plumbing only: the reader returns a synthetic.BAProblem whose cameras are in the stored form
of NLLS_VAR_PINHOLE (rotation matrix, translation, f, k1, k2), ready for set_variables / set_costs or the NLLSProblem mirror.
Plain and bz2-compressed files are accepted.
"""
import bz2
import io

import numpy as np

from . import synthetic


def _open(path, mode="rt"):
    return bz2.open(path, mode) if str(path).endswith(".bz2") else open(path, mode)


def read_bal(path):
    """-> synthetic.BAProblem (cameras (ncam, 15) stored pinhole form, points (npt, 3), 1-based global variable indices)."""
    with _open(path) as f:
        tok = np.array(f.read().split(), dtype=np.float64)
    ncam, npt, nobs = int(tok[0]), int(tok[1]), int(tok[2])
    need = 3 + 4 * nobs + 9 * ncam + 3 * npt
    if tok.size != need:
        raise ValueError(f"BAL file {path}: expected {need} numbers, found {tok.size}")
    obs = tok[3:3 + 4 * nobs].reshape(nobs, 4)
    cam_i, pt_i, z = obs[:, 0].astype(np.int64), obs[:, 1].astype(np.int64), obs[:, 2:4].copy()
    if cam_i.min() < 0 or cam_i.max() >= ncam or pt_i.min() < 0 or pt_i.max() >= npt:
        raise ValueError("BAL file: observation index out of range")
    cams9 = tok[3 + 4 * nobs:3 + 4 * nobs + 9 * ncam].reshape(ncam, 9)
    pts = tok[3 + 4 * nobs + 9 * ncam:].reshape(npt, 3).copy()
    cams = synthetic.pinhole_cameras(cams9[:, 0:3], cams9[:, 3:6], cams9[:, 6], cams9[:, 7], cams9[:, 8])
    return synthetic.BAProblem(cams, pts, cam_i + 1, pt_i + 1 + ncam, z)


def so3_log(R):
    """Rotation matrices (n, 3, 3) -> Rodrigues vectors (n, 3)."""
    R = np.asarray(R, dtype=np.float64)
    c = np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    v = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1) * 0.5   # sin(th) * axis
    s = np.sin(th)
    small = th < 1e-7
    k = np.where(small, 1.0 + th ** 2 / 6.0, th / np.where(small, 1.0, s))
    return v * k[:, None]


def write_bal(path, problem):
    """Inverse of read_bal (rotations go through the logarithm: angles below pi)."""
    R = problem.cameras[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    cams9 = np.concatenate([so3_log(R), problem.cameras[:, 9:12], problem.cameras[:, 12:15]], axis=1)
    buf = io.StringIO()
    buf.write(f"{problem.ncam} {problem.npt} {problem.nobs}\n")
    ci, pi = problem.cam_idx - 1, problem.pt_idx - 1 - problem.ncam
    for c, p, (x, y) in zip(ci, pi, problem.z):
        buf.write(f"{c} {p} {x:.17e} {y:.17e}\n")
    for v in cams9.ravel():
        buf.write(f"{v:.17e}\n")
    for v in problem.points.ravel():
        buf.write(f"{v:.17e}\n")
    with _open(path, "wt") as f:
        f.write(buf.getvalue())
