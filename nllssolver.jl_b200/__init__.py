"""nllssolver.jl_b200 — B200-native LM inner loop behind the NLLSsolver.jl problem API."""
from . import synthetic  # noqa: F401
from . import capi  # noqa: F401,E402
from . import bal  # noqa: F401,E402
from .api import (  # noqa: F401,E402
    NLLSProblem, NLLSOptions, NLLSResult, EuclideanVector, PinholeCamera, SimpleError2, SimpleError3, SimpleError4, AffineReprojection, PinholeReprojection,
    NoRobust, HuberKernel, Huber2oKernel, GemanMcclureKernel, Scaled, robustified, COST_DTYPE,
    ContaminatedGaussian, OffsetResidual, ADAPTIVE_DTYPE,
    optimize, optimizesingles, convertunfixed, cost, nullcallback, printoutcallback, storecostscallback, emcallback, CostTrajectory,
    newton, levenbergmarquardt, dogleg, gradientdescent,
)
