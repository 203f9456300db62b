"""Host logic of the reduced camera system (csrc/reduced_plan.hpp) on the CPU: elimination orders, symbolic factorisation, levels and
the storage order by rank ownership (tests/native/reduced_plan_check.cpp, built here with g++)."""
import ctypes as C
import math
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def chk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("red") / "libredcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "native", "reduced_plan_check.cpp")])
    L = C.CDLL(so)
    L.reduced_plan_check.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_double)]
    L.reduced_owner_check.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    return L


@pytest.mark.parametrize("NT,nlong", [(149, 0), (149, 10), (149, 70), (1141, 0), (40, 5), (9, 2), (3, 0), (1, 0)])
def test_graph_order_is_valid_and_shallow(chk, NT, nlong):
    out = (C.c_double * 8)()
    for seed in (1, 2):
        assert chk.reduced_plan_check(NT, nlong, seed, 0, out) == 0
        nlev, ntiles, maxrows, same, mono = (int(v) for v in out[:5])
        assert same == 1 and mono == 1                      # symbolic fill == dense boolean elimination; levels respect every dependency
        if nlong == 0 and NT > 2:                            # block tridiagonal: a balanced elimination tree
            assert nlev <= math.ceil(math.log2(NT + 1)) + 1, (NT, nlev)
            assert maxrows <= 2
    # the Venice-shape case of DESIGN.md: 149 columns and a few second sub-diagonal tiles (~70 long tracks falling on a handful of tile
    # pairs): 8 levels where the band order (half-bandwidth 2 everywhere, leaves eliminated one after the other) needs 12
    if NT == 149 and nlong > 0:
        graph_levels = nlev
        assert chk.reduced_plan_check(NT, nlong, 2, 1, out) == 0 and int(out[3]) == 1 and int(out[4]) == 1
        assert graph_levels < int(out[0])
        if nlong == 10:
            assert graph_levels == 8 and int(out[0]) == 12


@pytest.mark.parametrize("NT,nranks", [(149, 2), (149, 8), (40, 3), (12, 8), (5, 2)])
def test_ownership_order(chk, NT, nranks):
    out = (C.c_double * 8)()
    assert chk.reduced_owner_check(NT, nranks, out) == 0
    block, shared, fill, slots = (int(v) for v in out[:4])
    assert slots == nranks * block + shared + fill
    assert shared <= 3 * (nranks - 1) + 1                    # only the tiles around a band boundary are shared
