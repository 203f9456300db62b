"""Host logic of the Schur v5 kernel (csrc/schur5_plan.hpp) on the CPU: the plan is executed by a plain C++ emulator
(tests/native/schur5_plan_check.cpp, built here with g++) and must reproduce the brute-force Schur complement — entry streams,
window coordinates, band ranges, masks, FLUSH addressing into the tile-sparse reduced system, outlier bookkeeping."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def chk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("s5") / "libs5check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "native", "schur5_plan_check.cpp")])
    L = C.CDLL(so)
    L.schur5_plan_check.restype = C.c_double
    L.schur5_plan_check.argtypes = [C.c_int, C.c_uint, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return L


@pytest.mark.parametrize("dc", [6, 9])
@pytest.mark.parametrize("case", [
    (30, 900, 5.0, 0, 4),      # banded, several CTAs, several super-tiles per CTA
    (13, 400, 4.0, 0, 1),      # one CTA, window close to the camera range's end
    (60, 3000, 6.5, 7, 9),     # long tracks (outliers by width) + every 7th point with a gap in its camera list (outliers by shape)
    (5, 50, 3.0, 0, 3),        # fewer cameras than a window
    (200, 2500, 2.5, 0, 16),   # short tracks, fast camera drift
])
def test_plan_reproduces_brute_force(chk, dc, case):
    nA, nB, kmean, scat, ncta = case
    for seed in (1, 2):
        st = (C.c_double * 8)()
        err = chk.schur5_plan_check(dc, seed, nA, nB, kmean, scat, ncta, st)
        assert err >= 0, f"plan check failed structurally: code {err}"
        assert err <= 1e-13, err
        if scat == 0 and kmean <= 5.0 and dc == 6:
            assert st[1] <= 0.08, f"outlier share {st[1]}"              # banded data: (almost) everything on the fast path
