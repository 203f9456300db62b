"""Host-side mirror of the reference's user API for the LM hot path (the Julia toolchain is absent in this image, so the
host side above the C ABI is Python; julia/NLLSsolverB200.jl holds the equivalent ccall glue).

Mirrors, with the same names and argument meaning (Julia's trailing `!` dropped):
    NLLSProblem, addvariable!, addcost!            src/problem.jl:5-25,90-122
    NLLSOptions, NLLSResult                        src/structs.jl:22-50
    optimize!(problem, options, unfixed, callback) src/optimize.jl:57
    cost(problem)                                  src/cost.jl:10
    nullcallback, printoutcallback, storecostscallback / CostTrajectory   src/callbacks.jl
Residual / variable / robust-kernel classes carry the id of their registered sm_100a kernel; a problem containing a
residual type without one is rejected with an error — there is no CPU fallback.
"""
import sys
import time

import numpy as np

from . import capi

# ------------------------------------------------------------------------------------------------
# Variables (src/variable.jl, src/robustadaptive.jl)
# ------------------------------------------------------------------------------------------------


class EuclideanVector(np.ndarray):
    """EuclideanVector{N,Float64} (src/variable.jl:8-10). N in {3, 6} has a registered update kernel."""

    def __new__(cls, values):
        return np.asarray(values, dtype=np.float64).reshape(-1).view(cls)


class PinholeCamera:
    """Repo-defined 9-DoF camera (the reference ships none): R (3x3), t, f, k1, k2."""

    def __init__(self, R, t, f, k1=0.0, k2=0.0):
        self.R = np.asarray(R, dtype=np.float64).reshape(3, 3)
        self.t = np.asarray(t, dtype=np.float64).reshape(3)
        self.f, self.k1, self.k2 = float(f), float(k1), float(k2)

    def stored(self):
        return np.concatenate([self.R.ravel(order="F"), self.t, [self.f, self.k1, self.k2]])

    @classmethod
    def from_stored(cls, v):
        return cls(np.asarray(v[:9]).reshape(3, 3, order="F"), v[9:12], v[12], v[13], v[14])


class ContaminatedGaussian:
    """ContaminatedGaussian(sigma1, sigma2, w): adaptive robust kernel *and* 3-DoF variable (src/robustadaptive.jl:3-23).
    Stored as (1/sigma1, 1/sigma2, w) with the narrowest Gaussian first; w is not flipped by the re-sort (:12-15)."""

    def __init__(self, s1, s2, w):
        a, b = 1.0 / float(s1), 1.0 / float(s2)
        if not (a >= b):
            a, b = b, a
        self.invsigma1, self.invsigma2, self.w = a, b, float(w)

    def stored(self):
        return np.array([self.invsigma1, self.invsigma2, self.w])

    @classmethod
    def from_stored(cls, v):
        k = cls.__new__(cls)
        k.invsigma1, k.invsigma2, k.w = float(v[0]), float(v[1]), float(v[2])
        return k

    def params(self):                                                      # src/robustadaptive.jl:23
        return np.array([1.0 / self.invsigma1, 1.0 / self.invsigma2, self.w])


def _vartype(v):
    if isinstance(v, PinholeCamera):
        return capi.VAR_PINHOLE, v.stored()
    if isinstance(v, ContaminatedGaussian):
        return capi.VAR_CONTAMGAUSS, v.stored()
    if isinstance(v, (float, int)):
        return capi.VAR_SCALAR, np.array([float(v)])
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 3:
        return capi.VAR_EUCLID3, a
    if a.size == 6:
        return capi.VAR_EUCLID6, a
    raise TypeError(f"variable of size {a.size} has no registered update kernel")


# ------------------------------------------------------------------------------------------------
# Robust kernels (src/robust.jl)
# ------------------------------------------------------------------------------------------------
class NoRobust:
    id, params = capi.ROBUST_NONE, ()


class HuberKernel:
    id = capi.ROBUST_HUBER

    def __init__(self, width):
        self.params = (float(width),)


class Huber2oKernel(HuberKernel):
    id = capi.ROBUST_HUBER2O


class GemanMcclureKernel:
    id = capi.ROBUST_GEMANMCCLURE

    def __init__(self, width):
        self.params = (float(width),)


class Scaled:
    def __init__(self, robust, height):
        self.id = robust.id | capi.ROBUST_SCALED
        self.params = ((robust.params[0] if robust.params else 0.0), float(height))


# ------------------------------------------------------------------------------------------------
# Residuals (src/residual.jl:4-14)
# ------------------------------------------------------------------------------------------------
COST_DTYPE = np.dtype([("z", "<f8", 2), ("varind", "<i8", 2)])  # memory image of SimpleError2{2,Float64,..}


class SimpleError2:
    """SimpleError2{N,T,V1,V2}(measurement, vi1, vi2)  (src/residual.jl:4-9)."""
    restype = None  # no registered kernel for the generic type
    robustkernel = NoRobust()

    def __init__(self, measurement, vi1, vi2):
        self.measurement = np.asarray(measurement, dtype=np.float64).reshape(2)
        self.varind = (int(vi1), int(vi2))


class SimpleError3:
    """SimpleError3{N,T,V1,V2,V3}(measurement, vi1, vi2, vi3)  (src/residual.jl:16-27): a measurement-error residual over three
    variables.  The reference ships the type without any `generatemeasurement`; no sm_100a kernel is registered for it here, so a
    problem that holds one is rejected with NLLS_ERR_NO_KERNEL when it is optimised (no CPU fallback)."""
    restype = None
    robustkernel = NoRobust()
    ndeps = 3

    def __init__(self, measurement, vi1, vi2, vi3):
        self.measurement = np.asarray(measurement, dtype=np.float64).ravel()
        self.varind = (int(vi1), int(vi2), int(vi3))


class SimpleError4:
    """SimpleError4{N,T,V1,V2,V3,V4}(measurement, vi1, vi2, vi3, vi4)  (src/residual.jl:29-38); see SimpleError3."""
    restype = None
    robustkernel = NoRobust()
    ndeps = 4

    def __init__(self, measurement, vi1, vi2, vi3, vi4):
        self.measurement = np.asarray(measurement, dtype=np.float64).ravel()
        self.varind = (int(vi1), int(vi2), int(vi3), int(vi4))


class AffineReprojection(SimpleError2):
    """SimpleError2{2,Float64,EuclideanVector{6},EuclideanVector{3}} with generatemeasurement(pose, X) =
    (pose[1:3].X, pose[4:6].X)  (test/optimizeba.jl:4)."""
    restype = capi.RES_AFFINE_BA


class PinholeReprojection(SimpleError2):
    """Pinhole (BAL convention) reprojection error of a PinholeCamera and a 3-D point (repo-defined)."""
    restype = capi.RES_PINHOLE_BA


ADAPTIVE_DTYPE = np.dtype([("data", "<f8"), ("varind", "<i8")])  # memory image of SimpleResidual (test/adaptivecost.jl:3-6)


class OffsetResidual:
    """AbstractAdaptiveResidual r = mean - data (examples/adaptivekernel.jl:9-18, test/adaptivecost.jl:3-13):
    varindices = (kernel variable, mean variable); the ContaminatedGaussian kernel is the first variable of the block."""
    restype = capi.RES_ADAPTIVE_OFFSET
    robustkernel = NoRobust()

    def __init__(self, data, varind, kernelind=1):
        self.data = float(data)
        self.varind = (int(kernelind), int(varind))


def robustified(base, kernel):
    """A residual type whose robustkernel(res) returns `kernel` (README.md:26-35)."""
    return type(f"{base.__name__}_{type(kernel).__name__}", (base,), {"robustkernel": kernel})


# ------------------------------------------------------------------------------------------------
# Options / result (src/structs.jl)
# ------------------------------------------------------------------------------------------------
newton, levenbergmarquardt, dogleg, gradientdescent = capi.ITER_NEWTON, capi.ITER_LM, capi.ITER_DOGLEG, capi.ITER_GD


class NLLSOptions:
    def __init__(self, maxiters=100, reldcost=1e-15, absdcost=1e-15, dstep=1e-15, maxfails=3, maxtime=30.0,
                 iterator=levenbergmarquardt, callback=None, iteratordata=None):
        self.reldcost, self.absdcost, self.dstep = reldcost, absdcost, dstep
        self.maxfails, self.maxiters = maxfails, maxiters
        self.maxtime = int(round(maxtime * 1e9))  # ns, like the reference constructor (src/structs.jl:33-35)
        self.iterator, self.callback, self.iteratordata = iterator, callback, iteratordata

    def c(self):
        return capi.Options(self.reldcost, self.absdcost, self.dstep, self.maxfails, self.maxiters, self.maxtime, self.iterator, 0)


class NLLSResult:
    FIELDS = ("startcost", "bestcost", "timetotal", "timeinit", "timecost", "timegradient", "timesolver", "termination",
              "niterations", "costcomputations", "gradientcomputations", "linearsolvers")

    def __init__(self, r):
        for f in self.FIELDS:
            setattr(self, f, getattr(r, f))

    def __repr__(self):
        return (f"NLLSsolver optimization took {self.timetotal:f} seconds and {self.niterations} iterations to reduce the cost from "
                f"{self.startcost:e} to {self.bestcost:e}, using {self.costcomputations} cost computations, "
                f"{self.gradientcomputations} gradient computations and {self.linearsolvers} linear solves; termination {self.termination:#x}")


# ------------------------------------------------------------------------------------------------
# Problem (src/problem.jl)
# ------------------------------------------------------------------------------------------------
class NLLSProblem:
    def __init__(self, device=0):
        self.variables = []      # problem.variables: python objects, 1-based indices via addvariable
        self.costs = {}          # VectorRepo: {cost type: list of costs or structured array}
        self._ctx = None
        self._dirty = True
        self._device = device
        self._shard = None       # (rank, nranks, uid) for multi-GPU

    # -- construction
    def addvariable(self, variable):
        """addvariable!(problem, variable) -> 1-based index  (src/problem.jl:114-122)."""
        _vartype(variable)  # validates nvars > 0 / registered type
        self.variables.append(variable)
        self._dirty = True
        return len(self.variables)

    def addvariables(self, array2d):
        """Bulk addvariable! of EuclideanVectors (rows). Returns the index of the first one."""
        a = np.asarray(array2d, dtype=np.float64)
        first = len(self.variables) + 1
        self.variables.extend(EuclideanVector(r) for r in a)
        self._dirty = True
        return first

    def addcost(self, cost):
        """addcost!(problem, cost)  (src/problem.jl:90-107)."""
        if not isinstance(cost, (SimpleError2, SimpleError3, SimpleError4, OffsetResidual)):
            raise TypeError("unsupported cost")
        if isinstance(cost, OffsetResidual):                               # src/problem.jl:97
            assert isinstance(self.variables[cost.varind[0] - 1], ContaminatedGaussian), "adaptive residual: first variable must be the kernel"
        for vi in cost.varind:
            assert 1 <= vi <= len(self.variables), "Problem with varindices()"
        lst = self.costs.setdefault(type(cost), [])
        if isinstance(lst, np.ndarray):
            raise TypeError("cost type was bulk-loaded; use addcosts")
        lst.append(cost)
        self._dirty = True

    def addcosts(self, costtype, aos):
        """Bulk addcost!: `aos` is the memory image of Vector{costtype} (COST_DTYPE / ADAPTIVE_DTYPE)."""
        aos = np.ascontiguousarray(aos, dtype=ADAPTIVE_DTYPE if costtype.restype == capi.RES_ADAPTIVE_OFFSET else COST_DTYPE)
        if costtype in self.costs and len(self.costs[costtype]):
            raise ValueError("bulk load into a non-empty cost vector")
        self.costs[costtype] = aos
        self._dirty = True

    def numcosts(self):
        return sum(len(v) for v in self.costs.values())

    # -- multi-GPU: this process owns a shard of the points; call before optimize
    def set_shard(self, rank, nranks, uid):
        self._shard = (rank, nranks, uid)
        self._dirty = True

    # -- device context
    def _gather_variables(self):
        groups = {}
        for i, v in enumerate(self.variables):
            vt, vals = _vartype(v)
            g = groups.setdefault(vt, ([], []))
            g[0].append(i + 1)
            g[1].append(vals)
        return {vt: (np.array(idx, dtype=np.int64), np.stack(vals)) for vt, (idx, vals) in groups.items()}

    def _one_cost_aos(self, ctype, lst):
        if getattr(ctype, "restype", None) is None:
            raise capi.NLLSError(capi.ERR_NO_KERNEL, f"residual type {ctype.__name__} has no registered sm_100a kernel (no CPU fallback)")
        if isinstance(lst, np.ndarray):
            aos = lst
        elif ctype.restype == capi.RES_ADAPTIVE_OFFSET:
            aos = np.zeros(len(lst), dtype=ADAPTIVE_DTYPE)
            aos["data"] = [c.data for c in lst]
            aos["varind"] = [c.varind[1] for c in lst]
            kinds = {c.varind[0] for c in lst}
            assert len(kinds) == 1, "all adaptive residuals must share one kernel variable"
            self._kernel_var = kinds.pop()
        else:
            aos = np.zeros(len(lst), dtype=COST_DTYPE)
            aos["z"] = np.stack([c.measurement for c in lst])
            aos["varind"] = np.array([c.varind for c in lst], dtype=np.int64)
        return aos

    def _cost_sets(self):
        """[(residual type, AoS image)] of the non-empty Vector{T}s of problem.costs (src/VectorRepo.jl:3), in insertion order.  Several
        residual types are summed (src/cost.jl:54) when they share one registered kernel family (same `restype`: same residual
        struct, same variable classes) and differ in their robustkernel(); anything else has no kernel."""
        live = [(t, c) for t, c in self.costs.items() if len(c)]
        if not live:
            raise capi.NLLSError(capi.ERR_INVALID, "no costs")
        sets = [(t, self._one_cost_aos(t, c)) for t, c in live]
        if len({t.restype for t, _ in sets}) != 1 or (len(sets) > 1 and sets[0][0].restype == capi.RES_ADAPTIVE_OFFSET):
            raise capi.NLLSError(capi.ERR_UNSUPPORTED, "residual types of one problem must share one registered kernel family (they may differ in the robust kernel)")
        return sets

    def _cost_aos(self):
        return self._cost_sets()[0]

    def context(self):
        """Create / refresh the device context (≙ makesymmvls + NLLSInternal, src/optimize.jl:16)."""
        if self._ctx is None:
            self._ctx = capi.Context(self._device)
            if self._shard is not None:
                self._ctx.comm_init(*self._shard)
            self._dirty = True
        groups = self._gather_variables()
        if self._dirty:
            sets = self._cost_sets()
            ctype, aos = sets[0]
            for vt, (idx, vals) in groups.items():
                self._ctx.set_variables(vt, vals, indices=idx)
            k = ctype.robustkernel
            kernel_var = 0
            if ctype.restype == capi.RES_ADAPTIVE_OFFSET:
                kernel_var = getattr(self, "_kernel_var", None) or next(i + 1 for i, v in enumerate(self.variables) if isinstance(v, ContaminatedGaussian))
            self._ctx.set_costs(ctype.restype, aos, k.id, k.params, kernel_var)
            for t2, aos2 in sets[1:]:
                self._ctx.add_costs(t2.restype, aos2, t2.robustkernel.id, t2.robustkernel.params)
            self._ctx.prepare()
            self._dirty = False
        else:
            for vt, (idx, vals) in groups.items():
                self._ctx.set_variables(vt, vals, indices=idx)
        self._groups = groups
        return self._ctx

    def _pull_variables(self, which=0):
        for vt, (idx, vals) in self._groups.items():
            out = self._ctx.get_variables(vt, len(idx), vals.shape[1], which)
            for i, row in zip(idx, out):
                old = self.variables[i - 1]
                if isinstance(old, PinholeCamera):
                    self.variables[i - 1] = PinholeCamera.from_stored(row)
                elif isinstance(old, ContaminatedGaussian):
                    self.variables[i - 1] = ContaminatedGaussian.from_stored(row)
                elif isinstance(old, (float, int)):
                    self.variables[i - 1] = float(row[0])
                else:
                    self.variables[i - 1] = EuclideanVector(row)


def cost(problem):
    """cost(problem)  (src/cost.jl:10)."""
    return problem.context().cost(0)


# ------------------------------------------------------------------------------------------------
# Callbacks (src/callbacks.jl) — stay on the host, fed by nlls_iterinfo
# ------------------------------------------------------------------------------------------------
class IterData:
    """What the reference passes to callbacks as `data::NLLSInternal` (the fields callbacks read, src/callbacks.jl)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.starttime = time.perf_counter_ns()
        self.iternum = 0
        self.startcost = self.bestcost = float("nan")

    @property
    def x(self):  # data.linsystem.x
        return self.ctx.step()


def nullcallback(cost, *unused):                                          # src/callbacks.jl:20
    return cost, 0


class CostTrajectory:                                                     # src/callbacks.jl:63-100
    def __init__(self):
        self.costs, self.times_ns, self.trajectory = [], [], []

    def empty(self):
        self.__init__()


def storecostscallback(store):                                            # src/callbacks.jl:102-133
    def cb(cost, problem, data, iteratedata):
        if isinstance(store, CostTrajectory):
            store.costs.append(cost)
            store.times_ns.append(time.perf_counter_ns() - data.starttime)
            store.trajectory.append(data.x)
        else:
            store.append(cost)
        return cost, 0
    return cb


def printoutcallback(cost, problem, data, iteratedata):                   # src/callbacks.jl:39-60
    if data.iternum == 1:
        print("iter      cost      cost_change    |step|    trust region")
        print(f"{0:4d} {data.startcost:12.4e}")
    print(f"{data.iternum:4d} {cost:12.4e} {data.bestcost - cost:12.4e} {iteratedata.stepnorm:12.4e} {1.0 / iteratedata.lambda_:12.4e}")
    return cost, 0


def emcallback(cost, problem, data, iteratedata):                         # test/adaptivecost.jl:15-25
    """The reference test's EM callback: refit the adaptive kernel of varnext from the squared residuals (optimize(kernel,
    squarederrors), src/robustadaptive.jl:48-73 — on the device), recompute the cost of varnext, count the evaluation."""
    data.ctx.adaptive_em(which=1)
    return data.ctx.cost(1), 0


# ------------------------------------------------------------------------------------------------
# optimize!
# ------------------------------------------------------------------------------------------------
def convertunfixed(unfixed, problem):
    """src/optimize.jl:19-22: nothing -> all; a type -> the variables of that type; an integer -> that variable alone; else as given."""
    if unfixed is None:
        return None
    n = len(problem.variables)
    if isinstance(unfixed, type):
        return np.array([isinstance(v, unfixed) for v in problem.variables], dtype=np.uint8)
    if isinstance(unfixed, (int, np.integer)) and not isinstance(unfixed, bool):
        m = np.zeros(n, dtype=np.uint8)
        m[int(unfixed) - 1] = 1
        return m
    m = np.asarray(unfixed).astype(np.uint8)
    assert len(m) == n, "unfixed must have one entry per variable"
    return m


def optimizesingles(problem, options=None, vartype=None):
    """optimizesingles!(problem, options, type)  (src/optimize.jl:60-76): every variable of `vartype` optimised on its own with all
    other variables fixed.  A batched CUDA kernel exists for the 3-D point type of the bundle-adjustment residuals; other types raise."""
    options = options or NLLSOptions()
    ctx = problem.context()
    ctx.set_unfixed(None)
    sample = next((v for v in problem.variables if isinstance(v, vartype)), None) if isinstance(vartype, type) else None
    vt = _vartype(sample)[0] if sample is not None else vartype
    ctx.optimize_singles(vt, options.c())
    problem._pull_variables(0)


def optimize(problem, options=None, unfixed=None, callback=nullcallback):
    """optimize!(problem, options, unfixed, callback)::NLLSResult  (src/optimize.jl:57).
    The LM loop runs in the CUDA library; variables are updated in place.  With a callback the loop is driven from
    here so that callback(cost, problem, data, iteratedata) -> (cost, terminate) runs exactly where the reference
    calls it (src/optimize.jl:128)."""
    options = options or NLLSOptions()
    assert len(problem.variables) > 0
    ctx = problem.context()
    ctx.set_unfixed(convertunfixed(unfixed, problem))
    copts = options.c()
    if callback is nullcallback:
        res = ctx.optimize(copts)
    else:
        data = IterData(ctx)
        ctx.lm_begin(copts)
        data.startcost = data.bestcost = ctx.cost(0)
        conv = 0
        while conv == 0:
            info = ctx.lm_iterate()
            data.iternum += 1
            cost_, terminate = callback(info.cost, problem, data, info)
            conv = ctx.lm_advance(cost_, int(terminate))
            data.bestcost = min(data.bestcost, cost_) if cost_ == cost_ else data.bestcost
        res = ctx.lm_end()
    problem._pull_variables(0)
    return NLLSResult(res)
