"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/liboracle.so, the CPU restatement of the reference's LM hot path
(see oracle/nlls_oracle.hpp for the file:line map and the parity-pinning statement).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# ids shared with oracle/nlls_oracle.hpp
VT_EUCLID, VT_CONTAMGAUSS, VT_PINHOLE = 0, 1, 2
IT_NEWTON, IT_LM, IT_DOGLEG, IT_GD = 0, 1, 2, 3   # src/structs.jl:4
RT_AFFINE_BA, RT_PINHOLE_BA, RT_ADAPTIVE_OFFSET, RT_ROSENBROCK_A, RT_ROSENBROCK_B = 1, 2, 3, 4, 5
RK_NONE, RK_HUBER, RK_HUBER2O, RK_GEMANMCCLURE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("nlls_oracle.cpp", "oracle_capi.cpp", "nlls_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_robustify.restype = C.c_double
        L.orc_robustify.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]
        L.orc_robustifydcost.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _dp]
        L.orc_cg_make.argtypes = [C.c_double, C.c_double, C.c_double, _dp]
        L.orc_cg_robustify.restype = C.c_double
        L.orc_cg_robustify.argtypes = [_dp, C.c_double]
        L.orc_cg_robustifydcost.argtypes = [_dp, C.c_double, _dp]
        L.orc_cg_robustifydkernel.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.orc_update_variable.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp]
        L.orc_make_pinhole.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_resjac.argtypes = [C.c_int, _dp, C.c_int, _i32p, _i32p, _dp, _i32p, _i32p, _dp, _dp]
        L.orc_problem_new.restype = C.c_void_p
        L.orc_problem_free.argtypes = [C.c_void_p]
        L.orc_set_elimination_order.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_unfixed.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int64]
        L.orc_optimizesingles.restype = C.c_int64
        L.orc_optimizesingles.argtypes = [C.c_void_p, C.c_void_p, _ip, C.c_int64]
        L.orc_add_variables.restype = C.c_int64
        L.orc_add_variables.argtypes = [C.c_void_p, C.c_int, C.c_int64, _dp, C.c_int]
        L.orc_add_costs.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_double, C.c_int, C.c_double]
        for f in ("orc_num_variables", "orc_variables_len", "orc_dof", "orc_hess_len"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_is_sparse.argtypes = [C.c_void_p]
        for f in ("orc_get_variables", "orc_set_variables", "orc_get_hess_data", "orc_get_grad", "orc_get_step", "orc_get_hess_dense"):
            getattr(L, f).argtypes = [C.c_void_p, _dp]
        L.orc_cost.restype = C.c_double
        L.orc_cost.argtypes = [C.c_void_p]
        L.orc_linearize.restype = C.c_double
        L.orc_linearize.argtypes = [C.c_void_p]
        L.orc_solve.argtypes = [C.c_void_p, C.c_double, _dp]
        L.orc_optimize.restype = C.c_int64
        L.orc_optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_set_callback.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_callback.restype = None
        L.orc_em_optimize.argtypes = [_dp, _dp, C.c_int64, C.c_int]
        L.orc_em_optimize.restype = None
        L.orc_bsm_new.restype = C.c_void_p
        L.orc_bsm_new.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _i32p, _i32p]
        L.orc_bsm_free.argtypes = [C.c_void_p]
        for f in ("orc_bsm_nnz", "orc_bsm_m", "orc_bsm_n"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_bsm_start.restype = C.c_int64
        L.orc_bsm_start.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_bsm_setblock.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp, C.c_int64]
        for f in ("orc_bsm_data", "orc_bsm_todense", "orc_bsm_symmetrifyfull"):
            getattr(L, f).argtypes = [C.c_void_p, _dp]
        L.orc_bsm_uniformscaling.argtypes = [C.c_void_p, C.c_double]
        L.orc_bsm_sparse.restype = C.c_int64
        L.orc_bsm_sparse.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, C.c_int64]
        L.orc_rle.restype = C.c_int64
        L.orc_rle.argtypes = [_ip, C.c_int64, _ip]
        L.orc_solve_dense.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_solve_sparse.argtypes = [C.c_int64, _ip, _ip, _dp, _dp, _dp]
        L.orc_fast_bAb_dense.restype = C.c_double
        L.orc_fast_bAb_dense.argtypes = [C.c_int, _dp, _dp]
        L.orc_fast_bAb_sparse.restype = C.c_double
        L.orc_fast_bAb_sparse.argtypes = [C.c_int64, _ip, _ip, _dp, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


class Options(C.Structure):
    """NLLSOptions (src/structs.jl:22-35); callback_terminate emulates a callback's `terminate`."""
    _fields_ = [("reldcost", C.c_double), ("absdcost", C.c_double), ("dstep", C.c_double),
                ("maxfails", C.c_int64), ("maxiters", C.c_int64), ("maxtime_ns", C.c_uint64),
                ("callback_terminate", C.c_int64), ("iterator", C.c_int64)]

    def __init__(self, maxiters=100, reldcost=1e-15, absdcost=1e-15, dstep=1e-15, maxfails=3, maxtime=30.0, callback_terminate=0,
                 iterator=1):
        super().__init__(reldcost, absdcost, dstep, maxfails, maxiters, int(round(maxtime * 1e9)), callback_terminate, iterator)


class Result(C.Structure):
    """NLLSResult (src/structs.jl:37-50)."""
    _fields_ = [(n, C.c_double) for n in ("startcost", "bestcost", "timetotal", "timeinit", "timecost", "timegradient", "timesolver")] + \
               [(n, C.c_int64) for n in ("termination", "niterations", "costcomputations", "gradientcomputations", "linearsolvers")]


class IterRecord(C.Structure):
    _fields_ = [("cost", C.c_double), ("lambda_", C.c_double), ("maxstep", C.c_double), ("ntries", C.c_int64)]


# ---------------------------------------------------------------------------------------------
def robustify(kind, width, cost, scaled=False, height=1.0):
    return lib().orc_robustify(kind, width, int(scaled), height, cost)


def robustifydcost(kind, width, cost, scaled=False, height=1.0):
    out = np.zeros(3)
    lib().orc_robustifydcost(kind, width, int(scaled), height, cost, _p(out))
    return out


def cg_make(s1, s2, w):
    out = np.zeros(3)
    lib().orc_cg_make(s1, s2, w, _p(out))
    return out


def cg_robustify(p3, cost):
    p3 = _d(p3)
    return lib().orc_cg_robustify(_p(p3), cost)


def cg_robustifydcost(p3, cost):
    p3 = _d(p3)
    out = np.zeros(3)
    lib().orc_cg_robustifydcost(_p(p3), cost, _p(out))
    return out


def cg_robustifydkernel(p3, cost):
    p3 = _d(p3)
    val = np.zeros(1)
    g = np.zeros(4)
    H = np.zeros(16)
    lib().orc_cg_robustifydkernel(_p(p3), cost, _p(val), _p(g), _p(H))
    return val[0], g, H.reshape(4, 4).T


def update_variable(vtype, v, x):
    v = _d(v)
    x = _d(x)
    out = np.zeros_like(v)
    lib().orc_update_variable(vtype, _p(v), len(v), _p(x), _p(out))
    return out


def make_pinhole(rod, t, f, k1, k2):
    out = np.zeros(15)
    rod = _d(rod)
    t = _d(t)
    lib().orc_make_pinhole(_p(rod), _p(t), f, k1, k2, _p(out))
    return out


def resjac(rtype, data, variables):
    """variables: list of (vtype, values). Returns (r[m], J[m, P])."""
    d = np.zeros(4)
    d[:len(data)] = data
    nd = len(variables)
    vt = np.array([v[0] for v in variables], dtype=np.int32)
    vn = np.array([len(v[1]) for v in variables], dtype=np.int32)
    vv = np.zeros((nd, 16))
    for i, v in enumerate(variables):
        vv[i, :len(v[1])] = v[1]
    m = C.c_int(0)
    P = C.c_int(0)
    r = np.zeros(4)
    J = np.zeros(64)
    lib().orc_resjac(rtype, _p(d), nd, vt.ctypes.data_as(_i32p), vn.ctypes.data_as(_i32p), _p(vv), C.byref(m), C.byref(P), _p(r), _p(J))
    return r[:m.value].copy(), J[:m.value * P.value].reshape(P.value, m.value).T.copy()


def rle(sortedints):
    s = np.ascontiguousarray(sortedints, dtype=np.int64)
    out = np.zeros(int(s[-1]) + 2, dtype=np.int64)
    n = lib().orc_rle(_pi(s), len(s), _pi(out))
    return out[:n]


def solve_dense(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = _d(b)
    x = np.zeros_like(b)
    how = lib().orc_solve_dense(len(b), A.ctypes.data_as(_dp), _p(b), _p(x))
    return x, how


def solve_sparse(A_csc, b):
    """A_csc: scipy.sparse CSC, full symmetric."""
    A = A_csc.tocsc()
    A.sort_indices()
    colptr = (A.indptr + 1).astype(np.int64)
    rowval = (A.indices + 1).astype(np.int64)
    nz = _d(A.data)
    b = _d(b)
    x = np.zeros_like(b)
    lib().orc_solve_sparse(A.shape[0], _pi(colptr), _pi(rowval), _p(nz), _p(b), _p(x))
    return x


def fast_bAb_dense(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = _d(b)
    return lib().orc_fast_bAb_dense(len(b), A.ctypes.data_as(_dp), _p(b))


def fast_bAb_sparse(A_csc, b):
    A = A_csc.tocsc()
    A.sort_indices()
    colptr = (A.indptr + 1).astype(np.int64)
    rowval = (A.indices + 1).astype(np.int64)
    nz = _d(A.data)
    b = _d(b)
    return lib().orc_fast_bAb_sparse(A.shape[0], _pi(colptr), _pi(rowval), _p(nz), _p(b))


class BSM:
    """BlockSparseMatrix{Float64}(sparsitytransposed, rowblocksizes, colblocksizes) (src/BlockSparseMatrix.jl:30-47).
    `pattern` is the dense boolean (nrowblocks x ncolblocks) block pattern."""

    def __init__(self, pattern, rbs, cbs):
        pattern = np.asarray(pattern) > 0
        nrb, ncb = pattern.shape
        colptr = [1]
        rowval = []
        for r in range(nrb):
            cols = np.nonzero(pattern[r])[0] + 1
            rowval.extend(cols.tolist())
            colptr.append(len(rowval) + 1)
        self._colptr = np.array(colptr, dtype=np.int64)
        self._rowval = np.array(rowval if rowval else [0], dtype=np.int64)
        self._rbs = np.array(rbs, dtype=np.int32)
        self._cbs = np.array(cbs, dtype=np.int32)
        self.h = lib().orc_bsm_new(nrb, ncb, _pi(self._colptr), _pi(self._rowval), self._rbs.ctypes.data_as(_i32p), self._cbs.ctypes.data_as(_i32p))
        self.m = lib().orc_bsm_m(self.h)
        self.n = lib().orc_bsm_n(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bsm_free(self.h)
            self.h = None

    def nnz(self):
        return lib().orc_bsm_nnz(self.h)

    def validblock(self, i, j):
        return lib().orc_bsm_start(self.h, i, j) != 0

    def setblock(self, i, j, block):
        """block: 2-D array (rows x cols); stored column-major (src/BlockSparseMatrix.jl:102-105)."""
        v = np.asfortranarray(np.atleast_2d(np.asarray(block, dtype=np.float64))).ravel(order="F").copy()
        lib().orc_bsm_setblock(self.h, i, j, _p(v), len(v))

    def data(self):
        out = np.zeros(self.nnz())
        lib().orc_bsm_data(self.h, _p(out))
        return out

    def todense(self):
        out = np.zeros(self.m * self.n)
        lib().orc_bsm_todense(self.h, _p(out))
        return out.reshape(self.n, self.m).T

    def symmetrifyfull(self):
        out = np.zeros(self.m * self.m)
        lib().orc_bsm_symmetrifyfull(self.h, _p(out))
        return out.reshape(self.m, self.m).T

    def uniformscaling(self, k):
        lib().orc_bsm_uniformscaling(self.h, k)

    def sparse(self, symmetrify=False):
        import scipy.sparse as sp
        cap = 2 * self.nnz() + 16
        ncols = self.m if symmetrify else self.n
        colptr = np.zeros(ncols + 1, dtype=np.int64)
        rowval = np.zeros(cap, dtype=np.int64)
        vals = np.zeros(cap)
        nz = lib().orc_bsm_sparse(self.h, int(symmetrify), _pi(colptr), _pi(rowval), _p(vals), cap)
        assert nz >= 0
        return sp.csc_matrix((vals[:nz], rowval[:nz] - 1, colptr - 1), shape=(self.m, ncols))


class Problem:
    """Mirror of NLLSProblem + addvariable!/addcost!/cost/optimize! for the oracle (src/problem.jl, src/optimize.jl)."""

    def __init__(self):
        self.h = lib().orc_problem_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_free(self.h)
            self.h = None

    def set_elimination_order(self, mode):
        """0: default order of the sparse LDL'; 1: a second exact order (ties reversed) — used to measure solver-vs-solver drift."""
        lib().orc_set_elimination_order(self.h, mode)

    def set_unfixed(self, mask=None):
        """optimize!(problem, options, unfixed): boolean vector over the variables (None: all unfixed)."""
        if mask is None:
            lib().orc_set_unfixed(self.h, None, 0)
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            lib().orc_set_unfixed(self.h, m.ctypes.data_as(C.POINTER(C.c_ubyte)), len(m))

    def optimizesingles(self, indices, options=None):
        """optimizesingles!(problem, options, indices) with 1-based variable indices (src/optimize.jl:60-76)."""
        options = options or Options()
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        return lib().orc_optimizesingles(self.h, C.byref(options), _pi(idx), len(idx))

    def add_variables(self, vtype, values):
        """values: (n, nstore) array. Returns the 1-based index of the first variable added."""
        v = _d(np.atleast_2d(values))
        return lib().orc_add_variables(self.h, vtype, v.shape[0], _p(v), v.shape[1])

    def add_costs(self, rtype, varind, data, kernel=(RK_NONE, 0.0, False, 1.0)):
        vi = np.ascontiguousarray(np.atleast_2d(varind), dtype=np.int64)
        d = _d(np.atleast_2d(data))
        assert vi.shape[0] == d.shape[0]
        kind, width, scaled, height = kernel
        lib().orc_add_costs(self.h, rtype, vi.shape[0], vi.shape[1], _pi(vi), d.shape[1], _p(d), kind, width, int(scaled), height)

    def variables(self):
        out = np.zeros(lib().orc_variables_len(self.h))
        lib().orc_get_variables(self.h, _p(out))
        return out

    def set_variables(self, flat):
        flat = _d(flat)
        assert len(flat) == lib().orc_variables_len(self.h)
        lib().orc_set_variables(self.h, _p(flat))

    def cost(self):
        return lib().orc_cost(self.h)

    def linearize(self):
        """zero! + costgradhess! (src/optimize.jl:118). Returns cost."""
        return lib().orc_linearize(self.h)

    @property
    def dof(self):
        return lib().orc_dof(self.h)

    @property
    def is_sparse(self):
        return bool(lib().orc_is_sparse(self.h))

    def hess_data(self):
        out = np.zeros(lib().orc_hess_len(self.h))
        lib().orc_get_hess_data(self.h, _p(out))
        return out

    def hess_dense(self):
        n = self.dof
        out = np.zeros(n * n)
        lib().orc_get_hess_dense(self.h, _p(out))
        return out.reshape(n, n).T

    def grad(self):
        out = np.zeros(self.dof)
        lib().orc_get_grad(self.h, _p(out))
        return out

    def solve(self, lam):
        out = np.zeros(self.dof)
        lib().orc_solve(self.h, lam, _p(out))
        return out

    def set_callback(self, kind):
        """0: nullcallback; 1: the EM callback of test/adaptivecost.jl:15-25 (refit the adaptive kernel of varnext, recompute the cost)."""
        lib().orc_set_callback(self.h, int(kind))

    def optimize(self, options=None, maxtrace=4096):
        options = options or Options()
        res = Result()
        trace = (IterRecord * maxtrace)()
        n = lib().orc_optimize(self.h, C.byref(options), C.byref(res), trace, maxtrace)
        return res, [trace[i] for i in range(min(n, maxtrace))]


def em_optimize(kernel3, squarederrors, maxiters=10):
    """optimize(kernel::ContaminatedGaussian, squarederrors, maxiters) (src/robustadaptive.jl:48-73): kernel3 = stored
    (invsigma1, invsigma2, w) -> refitted stored triple."""
    k = _d(np.array(kernel3, dtype=np.float64).copy())
    sq = _d(squarederrors)
    lib().orc_em_optimize(_p(k), _p(sq), len(sq), int(maxiters))
    return k
