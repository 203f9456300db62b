"""nllssolver.jl_b200 — B200-native LM inner loop behind the NLLSsolver.jl problem API."""
from . import synthetic  # noqa: F401
