"""Pins for the two derivative sets no reference test holds (SURVEY §8c "parity unpinned"):

* autorobustifydkernel of ContaminatedGaussian — 4 gradient + 10 distinct Hessian entries of
      x -> robustify(update(kernel, x[1:3]), cost + x[4])          (src/autodiff.jl:164-165, src/robust.jl:15)
  with update(::ZeroToInfScalar / ::ZeroToOneScalar) of src/variable.jl:18-32 and robustify of src/robustadaptive.jl:25;
* computeresjac of the repo-defined pinhole residual: J = d r(update(cam, d[0:9]), X + d[9:12]) / d d at 0 (src/autodiff.jl:81-93).

Both oracle routines are exact forward-mode arithmetic written by the same author as the device code, so they are checked
here against something that shares no derivation with them: the DEFINITION above written once in mpmath and differentiated
NUMERICALLY at 60 significant digits (mp.diff, central differences of order 2 with h ~ 1e-15 relative to 60 digits => the
truncation error is ~1e-30).  The device kernels are compared with the oracle to <= 1e-12 in the GPU tests, which closes the chain."""
import itertools

import mpmath as mp
import numpy as np
import pytest

mp.mp.dps = 60
FLOATMIN = mp.mpf(2) ** -1022


def _rho_updated(kernel, cost, x):
    """robustify(update(kernel, x[0:3]), cost + x[3]) — no re-sort of the sigmas under differentiation (src/robustadaptive.jl:13)."""
    a0, b0, w0 = (mp.mpf(float(v)) for v in kernel)
    a = (a0 if a0 > 0 else FLOATMIN) * mp.e ** x[0]                        # src/variable.jl:22
    b = (b0 if b0 > 0 else FLOATMIN) * mp.e ** x[1]
    v = (w0 if w0 > 0 else FLOATMIN) * mp.e ** x[2]                        # src/variable.jl:30
    w = v / (1 + (v - w0))                                                 # :31
    c = mp.mpf(float(cost)) + x[3]
    s1sq, s2sq = a * a, b * b                                              # src/robustadaptive.jl:16-18
    return c * (s2sq / 2) - mp.log(w * a * mp.e ** (c * (s2sq - s1sq) / 2) + (1 - w) * b)   # :25


CG_CASES = [((1.0, 10.0, 0.8), 0.0), ((1.0, 10.0, 0.8), 0.49), ((1.0, 10.0, 0.8), 25.0), ((0.5, 5.0, 0.6), 1.3 ** 2),
            ((2.0, 2.5, 0.3), 4.0), ((0.1, 3.0, 0.95), 0.01), ((1.0, 50.0, 0.5), 900.0)]


@pytest.mark.parametrize("sig,cost", CG_CASES)
def test_contaminated_gaussian_kernel_derivatives_all_entries(orc, sig, cost):
    k = orc.cg_make(*sig)                                                  # (1/s1, 1/s2, w), narrowest first
    val, g, H = orc.cg_robustifydkernel(k, cost)
    f = lambda *x: _rho_updated(k, cost, x)
    zero = (0, 0, 0, 0)
    assert float(f(*zero)) == pytest.approx(val, rel=1e-13, abs=1e-13)
    g_fd = np.array([float(mp.diff(f, zero, tuple(int(i == j) for j in range(4)))) for i in range(4)])
    H_fd = np.zeros((4, 4))
    for i, j in itertools.combinations_with_replacement(range(4), 2):
        order = [0, 0, 0, 0]
        order[i] += 1
        order[j] += 1
        H_fd[i, j] = H_fd[j, i] = float(mp.diff(f, zero, tuple(order)))
    gs = max(1.0, np.max(np.abs(g_fd)))
    hs = max(1.0, np.max(np.abs(H_fd)))
    assert np.max(np.abs(g - g_fd)) <= 1e-12 * gs, (g, g_fd)               # all 4 gradient entries
    assert np.max(np.abs(H - H_fd)) <= 1e-12 * hs, (H, H_fd)               # all 10 distinct Hessian entries (+ symmetry)
    assert np.array_equal(H, H.T)
    # the closed forms the survey derived for the w re-parameterisation: dw'/dx3 = w (1 - w), d2 = w (1 - w)(1 - 2 w)
    w = sig[2]
    wf = lambda t: (mp.mpf(w) * mp.e ** t) / (1 + (mp.mpf(w) * mp.e ** t - mp.mpf(w)))
    assert float(mp.diff(wf, 0)) == pytest.approx(w * (1 - w), rel=1e-13)
    assert float(mp.diff(wf, 0, 2)) == pytest.approx(w * (1 - w) * (1 - 2 * w), rel=1e-12, abs=1e-14)


def _so3_exp(w):
    th = mp.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2)
    K = mp.matrix([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th == 0:
        return mp.eye(3)
    return mp.eye(3) + (mp.sin(th) / th) * K + ((1 - mp.cos(th)) / th ** 2) * (K * K)


def _pinhole_residual(cam, X, z, d):
    """r(update(cam, d[0:9]), X + d[9:12]): R <- Exp(d[0:3]) R, (t, f, k1, k2) += d[3:9]; BAL projection (repo-defined)."""
    R = mp.matrix(3, 3)
    for i in range(3):
        for j in range(3):
            R[i, j] = mp.mpf(float(cam[i + 3 * j]))                        # column-major storage
    R = _so3_exp(d[0:3]) * R
    t = [mp.mpf(float(cam[9 + i])) + d[3 + i] for i in range(3)]
    f, k1, k2 = (mp.mpf(float(cam[12 + i])) + d[6 + i] for i in range(3))
    Xv = mp.matrix([mp.mpf(float(X[i])) + d[9 + i] for i in range(3)])
    P = R * Xv
    P0, P1, P2 = P[0] + t[0], P[1] + t[1], P[2] + t[2]
    px, py = -P0 / P2, -P1 / P2
    n2 = px * px + py * py
    s = f * (1 + n2 * (k1 + k2 * n2))
    return [s * px - mp.mpf(float(z[0])), s * py - mp.mpf(float(z[1]))]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pinhole_resjac_against_numeric_differentiation(orc, seed):
    rng = np.random.default_rng(seed)
    cam = orc.make_pinhole(rng.standard_normal(3) * (0.05 if seed < 3 else 1.5), np.array([0.0, 0.0, -10.0]) + rng.standard_normal(3) * 0.3,
                           500.0 + 20.0 * rng.standard_normal(), 1e-2 * rng.standard_normal(), 1e-3 * rng.standard_normal())
    X = rng.uniform(-1.0, 1.0, 3)
    z = rng.standard_normal(2) * 20.0
    r, J = orc.resjac(orc.RT_PINHOLE_BA, z, [(orc.VT_PINHOLE, cam), (orc.VT_EUCLID, X)])
    assert J.shape == (2, 12)
    zero = [mp.mpf(0)] * 12
    r_mp = _pinhole_residual(cam, X, z, zero)
    assert np.allclose(r, [float(v) for v in r_mp], rtol=1e-12, atol=1e-10)
    J_fd = np.zeros((2, 12))
    for k in range(12):
        for i in range(2):
            fk = lambda tt, k=k, i=i: _pinhole_residual(cam, X, z, [tt if q == k else mp.mpf(0) for q in range(12)])[i]
            J_fd[i, k] = float(mp.diff(fk, 0))
    scale = np.max(np.abs(J_fd), axis=0)                                   # per column: the DoF have very different units
    assert np.all(np.abs(J - J_fd) <= 1e-11 * np.maximum(scale, 1e-300)), (np.abs(J - J_fd) / scale)
    # update(::pinhole) itself: rotation composed on the left, remaining parameters additive
    x = rng.standard_normal(9) * 0.1
    upd = orc.update_variable(orc.VT_PINHOLE, cam, x)
    E = _so3_exp([mp.mpf(float(v)) for v in x[:3]])
    R0 = mp.matrix(3, 3)
    for i in range(3):
        for j in range(3):
            R0[i, j] = mp.mpf(float(cam[i + 3 * j]))
    R1 = E * R0
    exp_upd = np.array([float(R1[i % 3, i // 3]) for i in range(9)] + [cam[9 + i] + x[3 + i] for i in range(6)])
    assert np.allclose(upd, exp_upd, rtol=1e-13, atol=1e-14)
