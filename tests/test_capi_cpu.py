"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares, and refuses to compute
without a CUDA device (no fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(pkg):
    hdr = open(os.path.join(ROOT, "include", "nlls_b200.h")).read()
    declared = set(re.findall(r"\b(nlls_[a-z0-9_]+)\s*\(", hdr))
    L = pkg.capi.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(pkg.capi.EXPORTS) == declared


def test_julia_glue_binds_declared_symbols(pkg):
    """julia/NLLSsolverB200.jl cannot run here (no Julia): at least every symbol it ccalls must be declared in the header and
    exported by the library, and the enum values it hard-codes must be the header's."""
    jl = open(os.path.join(ROOT, "julia", "NLLSsolverB200.jl")).read()
    hdr = open(os.path.join(ROOT, "include", "nlls_b200.h")).read()
    bound = set(re.findall(r"ccall\(\(:(nlls_[a-z0-9_]+)", jl))
    declared = set(re.findall(r"\b(nlls_[a-z0-9_]+)\s*\(", hdr))
    assert bound and bound <= declared, bound - declared
    L = pkg.capi.lib()
    assert all(hasattr(L, s) for s in bound)
    for jname, hname in [("VAR_PINHOLE", "NLLS_VAR_PINHOLE"), ("VAR_CONTAMGAUSS", "NLLS_VAR_CONTAMGAUSS"), ("RES_ADAPTIVE_OFFSET", "NLLS_RES_ADAPTIVE_OFFSET"),
                         ("RES_PINHOLE_BA", "NLLS_RES_PINHOLE_BA"), ("ROBUST_SCALED", "NLLS_ROBUST_SCALED"), ("ITER_DOGLEG", "NLLS_ITER_DOGLEG"),
                         ("ITER_GD", "NLLS_ITER_GD"), ("ROBUST_GEMANMCCLURE", "NLLS_ROBUST_GEMANMCCLURE")]:
        jv = int(re.search(rf"const {jname} = \w+\((\d+)\)", jl).group(1))
        hv = int(re.search(rf"{hname} = (\d+)", hdr).group(1))
        assert jv == hv, (jname, jv, hv)


def test_struct_sizes(pkg):
    import ctypes as C
    assert C.sizeof(pkg.capi.Options) == 56
    assert C.sizeof(pkg.capi.Result) == 96
    assert C.sizeof(pkg.capi.IterInfo) == 48
    assert pkg.COST_DTYPE.itemsize == 32  # SimpleError2{2,Float64,..}: 2 x f64 + 2 x i64 (src/residual.jl:4-7)


def test_no_device_no_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(pkg.capi.NLLSError) as e:
        pkg.capi.Context(0)
    assert e.value.code == pkg.capi.ERR_NO_DEVICE


def test_unregistered_residual_rejected(pkg):
    prob = pkg.NLLSProblem()
    a = prob.addvariable(pkg.EuclideanVector([0.0] * 6))
    b = prob.addvariable(pkg.EuclideanVector([0.0] * 3))
    prob.addcost(pkg.SimpleError2([0.0, 0.0], a, b))  # generic SimpleError2 without generatemeasurement: no kernel
    with pytest.raises(pkg.capi.NLLSError) as e:
        prob._cost_aos()
    assert e.value.code == pkg.capi.ERR_NO_KERNEL


def test_simpleerror3_and_4_are_rejected_without_a_kernel(pkg):
    """src/residual.jl:16-38: the reference exports the types; north star: residual types with no registered kernel are rejected."""
    for cls, nv in ((pkg.SimpleError3, 3), (pkg.SimpleError4, 4)):
        prob = pkg.NLLSProblem()
        idx = [prob.addvariable(pkg.EuclideanVector([0.0] * 3)) for _ in range(nv)]
        prob.addcost(cls([0.0, 0.0], *idx))
        assert prob.numcosts() == 1
        with pytest.raises(pkg.capi.NLLSError) as e:
            prob._cost_aos()
        assert e.value.code == pkg.capi.ERR_NO_KERNEL
    with pytest.raises(AssertionError):                       # varindices are checked like src/problem.jl:99-101
        prob.addcost(pkg.SimpleError3([0.0, 0.0], 1, 2, 99))


def test_unsupported_options(pkg):
    prob = pkg.NLLSProblem()
    prob.addvariable(pkg.EuclideanVector([0.0] * 6))
    with pytest.raises(pkg.capi.NLLSError):
        pkg.optimize(prob, pkg.NLLSOptions(iterator=pkg.dogleg))
    with pytest.raises(pkg.capi.NLLSError):
        pkg.optimize(prob, pkg.NLLSOptions(), unfixed=[True])


def test_bench_reference_arm(pkg):
    """bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line with the contract's keys, also for
    W = 0 and under a multi-rank launch where only rank 0 works."""
    import json
    import subprocess
    import sys
    for extra_env, args in [({}, ["--steps", "1", "--warmup", "0"]), ({}, ["--steps", "2", "--warmup", "1"]),
                            ({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ["--gpus", "2", "--steps", "1", "--warmup", "0"])]:
        env = dict(os.environ)
        env.update(extra_env)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ladybug"] + args,
                             capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if extra_env:
            assert not lines   # the other ranks exit 0 without work
            continue
        line = json.loads(lines[-1])
        assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "residual blocks/s"
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
        assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
