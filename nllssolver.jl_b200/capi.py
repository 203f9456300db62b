"""ctypes binding of include/nlls_b200.h — the same symbols the Julia glue (julia/NLLSsolverB200.jl) ccalls.

There is no fallback: if libnlls_b200.so is missing or no CUDA device is present, every compute entry point
raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libnlls_b200.so")

# enums of nlls_b200.h
OK, ERR_INVALID, ERR_NO_KERNEL, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_NO_DEVICE = range(7)
VAR_SCALAR, VAR_EUCLID3, VAR_EUCLID6, VAR_CONTAMGAUSS, VAR_PINHOLE = 1, 3, 6, 100, 101
RES_AFFINE_BA, RES_PINHOLE_BA, RES_ADAPTIVE_OFFSET = 1, 2, 3
ROBUST_NONE, ROBUST_HUBER, ROBUST_HUBER2O, ROBUST_GEMANMCCLURE, ROBUST_SCALED = 0, 1, 2, 3, 16
ITER_NEWTON, ITER_LM, ITER_DOGLEG, ITER_GD = 0, 1, 2, 3
TIME_LINEARIZE, TIME_LIN_POINT, TIME_LIN_CAM, TIME_COST, TIME_SCHUR, TIME_SOLVE_REDUCED, TIME_BACKSUB, TIME_TRY, TIME_MEMSET_H, TIME_LIN_LOOP = range(10)

EXPORTS = [
    "nlls_create", "nlls_destroy", "nlls_last_error", "nlls_version", "nlls_comm_unique_id", "nlls_comm_init",
    "nlls_set_variables", "nlls_set_costs", "nlls_add_costs", "nlls_set_unfixed", "nlls_optimize_singles", "nlls_prepare", "nlls_linearize", "nlls_cost", "nlls_adaptive_em", "nlls_solve", "nlls_update",
    "nlls_lm_begin", "nlls_lm_iterate", "nlls_lm_advance", "nlls_lm_step", "nlls_lm_end", "nlls_optimize", "nlls_get_variables", "nlls_dof",
    "nlls_get_gradient", "nlls_get_step", "nlls_hessian_len", "nlls_get_hessian_blocks", "nlls_hessian_nblocks",
    "nlls_get_hessian_index", "nlls_time_kernels", "nlls_timer_start", "nlls_timer_stop", "nlls_kernel_launches", "nlls_algorithmic_bytes", "nlls_algorithmic_flops",
]


class Options(C.Structure):
    """nlls_options == NLLSOptions (src/structs.jl:22-35)."""
    _fields_ = [("reldcost", C.c_double), ("absdcost", C.c_double), ("dstep", C.c_double), ("maxfails", C.c_int64),
                ("maxiters", C.c_int64), ("maxtime_ns", C.c_uint64), ("iterator", C.c_int32), ("reserved", C.c_int32)]


class Result(C.Structure):
    """nlls_result == NLLSResult (src/structs.jl:37-50)."""
    _fields_ = [(n, C.c_double) for n in ("startcost", "bestcost", "timetotal", "timeinit", "timecost", "timegradient", "timesolver")] + \
               [(n, C.c_int64) for n in ("termination", "niterations", "costcomputations", "gradientcomputations", "linearsolvers")]


class IterInfo(C.Structure):
    _fields_ = [("cost", C.c_double), ("lambda_", C.c_double), ("maxstep", C.c_double), ("stepnorm", C.c_double),
                ("ntries", C.c_int64), ("accepted", C.c_int64)]


class NLLSError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nlls_b200 error {code}: {msg}")
        self.code = code


_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("NLLS_B200_LIB", SO)   # development aid: a variant build of the same library (build.py: extra_flags / out)
        if not os.path.exists(so):
            raise RuntimeError(f"{so} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(so)
        vp = C.c_void_p
        L.nlls_create.argtypes = [C.POINTER(vp), C.c_int]
        L.nlls_destroy.argtypes = [vp]
        L.nlls_last_error.argtypes = [vp]
        L.nlls_last_error.restype = C.c_char_p
        L.nlls_comm_unique_id.argtypes = [C.c_void_p]
        L.nlls_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_void_p]
        L.nlls_set_variables.argtypes = [vp, C.c_int, _dp, C.c_int64, C.c_int64, C.c_int64, _ip]
        L.nlls_set_costs.argtypes = [vp, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, _dp, C.c_int, C.c_int64]
        L.nlls_adaptive_em.argtypes = [vp, C.c_int, C.c_int]
        L.nlls_add_costs.argtypes = [vp, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, _dp, C.c_int]
        L.nlls_prepare.argtypes = [vp]
        L.nlls_set_unfixed.argtypes = [vp, C.POINTER(C.c_ubyte), C.c_int64]
        L.nlls_optimize_singles.argtypes = [vp, C.c_int, C.POINTER(Options), _ip]
        L.nlls_linearize.argtypes = [vp, _dp]
        L.nlls_cost.argtypes = [vp, C.c_int, _dp]
        L.nlls_solve.argtypes = [vp, C.c_double]
        L.nlls_update.argtypes = [vp]
        L.nlls_lm_begin.argtypes = [vp, C.POINTER(Options)]
        L.nlls_lm_iterate.argtypes = [vp, C.POINTER(IterInfo)]
        L.nlls_lm_advance.argtypes = [vp, C.c_double, C.c_int64, _ip]
        L.nlls_lm_step.argtypes = [vp, C.c_void_p, _ip]
        L.nlls_lm_end.argtypes = [vp, C.POINTER(Result)]
        L.nlls_optimize.argtypes = [vp, C.POINTER(Options), C.POINTER(Result)]
        L.nlls_get_variables.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_int64, C.c_int64]
        for f in ("nlls_dof", "nlls_hessian_len", "nlls_hessian_nblocks", "nlls_kernel_launches"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int64
        L.nlls_get_gradient.argtypes = [vp, _dp]
        L.nlls_get_step.argtypes = [vp, _dp]
        L.nlls_get_hessian_blocks.argtypes = [vp, _dp]
        L.nlls_get_hessian_index.argtypes = [vp, _ip, _ip, _ip]
        L.nlls_time_kernels.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp]
        L.nlls_algorithmic_bytes.argtypes = [vp, C.c_int, _dp]
        L.nlls_algorithmic_flops.argtypes = [vp, C.c_int, _dp]
        L.nlls_timer_start.argtypes = [vp]
        L.nlls_timer_stop.argtypes = [vp, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """Thin object wrapper over nlls_ctx*: one per NLLSProblem."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().nlls_create(C.byref(self.h), device)
        if rc != OK:
            self.h = None
            raise NLLSError(rc, "nlls_create failed (no CUDA device?) — there is no CPU fallback")

    def close(self):
        if getattr(self, "h", None):
            lib().nlls_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise NLLSError(rc, lib().nlls_last_error(self.h).decode())

    # ---- definition
    def comm_init(self, rank, nranks, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._ck(lib().nlls_comm_init(self.h, rank, nranks, buf))

    def set_variables(self, vartype, values, first_index=1, indices=None):
        v = _d(np.atleast_2d(values))
        idx = None
        if indices is not None:
            idx = np.ascontiguousarray(indices, dtype=np.int64)
            assert len(idx) == v.shape[0]
        self._ck(lib().nlls_set_variables(self.h, vartype, v.ctypes.data_as(_dp), v.shape[0], v.shape[1], first_index,
                                          idx.ctypes.data_as(_ip) if idx is not None else None))

    def set_costs(self, restype, aos, robust=ROBUST_NONE, kparams=(), kernel_var=0):
        aos = np.ascontiguousarray(aos)
        kp = _d(list(kparams)) if len(kparams) else None
        self._ck(lib().nlls_set_costs(self.h, restype, aos.ctypes.data_as(C.c_void_p), aos.dtype.itemsize, aos.shape[0], robust,
                                      kp.ctypes.data_as(_dp) if kp is not None else None, len(kparams), kernel_var))

    def add_costs(self, restype, aos, robust=ROBUST_NONE, kparams=()):
        """A further cost set of the same residual struct with its own robust kernel (addcost! with a second residual type)."""
        aos = np.ascontiguousarray(aos)
        kp = _d(list(kparams)) if len(kparams) else None
        self._ck(lib().nlls_add_costs(self.h, restype, aos.ctypes.data_as(C.c_void_p), aos.dtype.itemsize, aos.shape[0], robust,
                                      kp.ctypes.data_as(_dp) if kp is not None else None, len(kparams)))

    def prepare(self):
        self._ck(lib().nlls_prepare(self.h))

    def adaptive_em(self, which=1, maxiters=10):
        """optimize(kernel, squarederrors, maxiters) (src/robustadaptive.jl:48-73) on the kernel variable of buffer `which`."""
        self._ck(lib().nlls_adaptive_em(self.h, which, maxiters))

    def set_unfixed(self, mask=None):
        """optimize!(problem, options, unfixed): boolean vector by variable position (None: all variables are optimised)."""
        if mask is None:
            self._ck(lib().nlls_set_unfixed(self.h, None, 0))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            self._ck(lib().nlls_set_unfixed(self.h, m.ctypes.data_as(C.POINTER(C.c_ubyte)), len(m)))

    def optimize_singles(self, vartype, options):
        """optimizesingles!(problem, options, type): returns the summed iteration count."""
        it = C.c_int64()
        self._ck(lib().nlls_optimize_singles(self.h, vartype, C.byref(options), C.byref(it)))
        return it.value

    # ---- the six operations
    def linearize(self):
        c = C.c_double()
        self._ck(lib().nlls_linearize(self.h, C.byref(c)))
        return c.value

    def cost(self, which=0):
        c = C.c_double()
        self._ck(lib().nlls_cost(self.h, which, C.byref(c)))
        return c.value

    def solve(self, lam):
        self._ck(lib().nlls_solve(self.h, lam))

    def update(self):
        self._ck(lib().nlls_update(self.h))

    def lm_begin(self, options):
        self._ck(lib().nlls_lm_begin(self.h, C.byref(options)))

    def lm_iterate(self):
        info = IterInfo()
        self._ck(lib().nlls_lm_iterate(self.h, C.byref(info)))
        return info

    def lm_advance(self, cost, terminate=0):
        conv = C.c_int64()
        self._ck(lib().nlls_lm_advance(self.h, cost, terminate, C.byref(conv)))
        return conv.value

    def lm_step(self):
        """One outer iteration with the null callback: (info, converged)."""
        info, conv = IterInfo(), C.c_int64()
        self._ck(lib().nlls_lm_step(self.h, C.byref(info), C.byref(conv)))
        return info, conv.value

    def lm_end(self):
        r = Result()
        self._ck(lib().nlls_lm_end(self.h, C.byref(r)))
        return r

    def optimize(self, options):
        r = Result()
        self._ck(lib().nlls_optimize(self.h, C.byref(options), C.byref(r)))
        return r

    # ---- read-back
    def get_variables(self, vartype, n, nstore, which=0, out=None):
        out = np.zeros((n, nstore)) if out is None else out
        self._ck(lib().nlls_get_variables(self.h, vartype, which, out.ctypes.data_as(_dp), n, nstore))
        return out

    @property
    def dof(self):
        return lib().nlls_dof(self.h)

    def gradient(self):
        out = np.zeros(self.dof)
        self._ck(lib().nlls_get_gradient(self.h, out.ctypes.data_as(_dp)))
        return out

    def step(self):
        out = np.zeros(self.dof)
        self._ck(lib().nlls_get_step(self.h, out.ctypes.data_as(_dp)))
        return out

    def hessian_blocks(self):
        out = np.zeros(lib().nlls_hessian_len(self.h))
        self._ck(lib().nlls_get_hessian_blocks(self.h, out.ctypes.data_as(_dp)))
        return out

    def hessian_index(self):
        n = lib().nlls_hessian_nblocks(self.h)
        rb, cb, st = (np.zeros(n, dtype=np.int64) for _ in range(3))
        self._ck(lib().nlls_get_hessian_index(self.h, rb.ctypes.data_as(_ip), cb.ctypes.data_as(_ip), st.ctypes.data_as(_ip)))
        return rb, cb, st

    # ---- measurement
    def time_kernels(self, which, reps=10, flush_l2=True):
        ms = C.c_double()
        self._ck(lib().nlls_time_kernels(self.h, which, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def timer_start(self):
        self._ck(lib().nlls_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(lib().nlls_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def kernel_launches(self):
        return lib().nlls_kernel_launches(self.h)

    def algorithmic_bytes(self, which):
        b = C.c_double()
        self._ck(lib().nlls_algorithmic_bytes(self.h, which, C.byref(b)))
        return b.value

    def algorithmic_flops(self, which):
        b = C.c_double()
        self._ck(lib().nlls_algorithmic_flops(self.h, which, C.byref(b)))
        return b.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().nlls_comm_unique_id(buf)
    if rc != OK:
        raise NLLSError(rc, "ncclGetUniqueId failed")
    return buf.raw
