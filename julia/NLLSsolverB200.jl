# NLLSsolverB200.jl — thin ccall glue between NLLSsolver.jl's problem API and libnlls_b200 (include/nlls_b200.h).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia toolchain (SURVEY.md F1).  Every ccall below is
# mirrored 1:1 by nllssolver.jl_b200/capi.py (ctypes), which the GPU test-suite exercises; keep the two in sync.
#
# Usage (a maintainer of NLLSsolver.jl would add this as a package extension):
#     using NLLSsolver, NLLSsolverB200
#     NLLSsolverB200.register_residual(MyReprojectionError, NLLSsolverB200.RES_AFFINE_BA)
#     result = NLLSsolverB200.optimize!(problem, NLLSOptions())          # same signature as NLLSsolver.optimize!
module NLLSsolverB200

using NLLSsolver, StaticArrays

const LIB = get(ENV, "NLLS_B200_LIB", joinpath(@__DIR__, "..", "nllssolver.jl_b200", "libnlls_b200.so"))

# enums of nlls_b200.h
const OK = Cint(0); const ERR_NO_KERNEL = Cint(2)
const VAR_SCALAR = Cint(1); const VAR_EUCLID3 = Cint(3); const VAR_EUCLID6 = Cint(6); const VAR_CONTAMGAUSS = Cint(100); const VAR_PINHOLE = Cint(101)
const RES_AFFINE_BA = Cint(1); const RES_PINHOLE_BA = Cint(2)
const ROBUST_NONE = Cint(0); const ROBUST_HUBER = Cint(1); const ROBUST_HUBER2O = Cint(2); const ROBUST_GEMANMCCLURE = Cint(3); const ROBUST_SCALED = Cint(16)

struct COptions            # nlls_options  == NLLSOptions (src/structs.jl:22-35)
    reldcost::Cdouble; absdcost::Cdouble; dstep::Cdouble
    maxfails::Int64; maxiters::Int64; maxtime_ns::UInt64
    iterator::Int32; reserved::Int32
end
struct CResult             # nlls_result   == NLLSResult (src/structs.jl:37-50)
    startcost::Cdouble; bestcost::Cdouble; timetotal::Cdouble; timeinit::Cdouble; timecost::Cdouble; timegradient::Cdouble; timesolver::Cdouble
    termination::Int64; niterations::Int64; costcomputations::Int64; gradientcomputations::Int64; linearsolvers::Int64
end
mutable struct CIterInfo   # nlls_iterinfo
    cost::Cdouble; lambda::Cdouble; maxstep::Cdouble; stepnorm::Cdouble; ntries::Int64; accepted::Int64
    CIterInfo() = new(0, 0, 0, 0, 0, 0)
end

# ---- registry: concrete residual type => id of its hand-written sm_100a kernel --------------------------------------
const RESIDUAL_KERNELS = Dict{DataType, Cint}()
register_residual(::Type{T}, id) where T = (RESIDUAL_KERNELS[T] = Cint(id))

vartype(::Type{Float64}) = VAR_SCALAR
vartype(::Type{SVector{3, Float64}}) = VAR_EUCLID3
vartype(::Type{SVector{6, Float64}}) = VAR_EUCLID6
vartype(::Type{T}) where T = error("NLLSsolverB200: variable type $T has no registered update kernel (no CPU fallback)")

# robustkernel(res) => (id, params)                                                    src/robust.jl
kernelspec(::NLLSsolver.NoRobust) = (ROBUST_NONE, Float64[])
kernelspec(k::NLLSsolver.HuberKernel) = (NLLSsolver.dynamic(k.secondorder) ? ROBUST_HUBER2O : ROBUST_HUBER, [Float64(k.width)])
kernelspec(k::NLLSsolver.GemanMcclureKernel) = (ROBUST_GEMANMCCLURE, [sqrt(Float64(k.width_squared))])
function kernelspec(k::NLLSsolver.Scaled)
    id, p = kernelspec(k.robust)
    return (id | ROBUST_SCALED, [isempty(p) ? 0.0 : p[1], Float64(k.height)])
end
kernelspec(k) = error("NLLSsolverB200: robust kernel $(typeof(k)) has no registered device implementation")

check(ctx, rc) = rc == OK ? nothing : error("nlls_b200 error $rc: " * unsafe_string(ccall((:nlls_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

# ---- optimize! -------------------------------------------------------------------------------------------------------
"""
    NLLSsolverB200.optimize!(problem, options=NLLSOptions(), unfixed=nothing, callback=nullcallback; device=0)

Drop-in for `NLLSsolver.optimize!` (src/optimize.jl:57) on the LM path.  Problems whose residual types have no registered
kernel, `unfixed` masks and non-LM iterators are rejected with an error.
"""
function optimize!(problem::NLLSProblem, options::NLLSOptions=NLLSOptions(), unfixed=nothing, callback=NLLSsolver.nullcallback; device::Integer=0)
    unfixed === nothing || error("NLLSsolverB200: `unfixed` masks are not implemented")
    options.iterator == NLLSsolver.levenbergmarquardt || error("NLLSsolverB200: only the Levenberg-Marquardt iterator is implemented")
    ctxref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nlls_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ctxref, device)
    rc == OK || error("nlls_create failed ($rc): no CUDA device — there is no CPU fallback")
    ctx = ctxref[]
    try
        # problem.variables: one call per concrete variable type, 1-based positions preserved (src/problem.jl:8,119-121)
        groups = Dict{DataType, Vector{Int}}()
        for (i, v) in enumerate(problem.variables)
            push!(get!(groups, typeof(v), Int[]), i)
        end
        bufs = Dict{DataType, Matrix{Float64}}()
        for (T, idx) in groups
            n = length(problem.variables[idx[1]])
            buf = Matrix{Float64}(undef, n, length(idx))              # column-major: one variable per column = AoS of n doubles
            for (k, i) in enumerate(idx); buf[:, k] .= problem.variables[i]; end
            bufs[T] = buf
            idx64 = Int64.(idx)
            GC.@preserve buf idx64 check(ctx, ccall((:nlls_set_variables, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Int64}), ctx, vartype(T), buf, length(idx), n, 0, idx64))
        end
        # problem.costs.data[T]: the Vector{T} of isbits structs is handed over as is (src/VectorRepo.jl:3)
        for (T, vec) in problem.costs.data
            isempty(vec) && continue
            haskey(RESIDUAL_KERNELS, T) || error("NLLSsolverB200: residual type $T has no registered sm_100a kernel (no CPU fallback)")
            isbitstype(T) || error("NLLSsolverB200: residual type $T is not isbits")
            id, kp = kernelspec(NLLSsolver.robustkernel(vec[1]))
            GC.@preserve vec kp check(ctx, ccall((:nlls_set_costs, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cdouble}, Cint, Int64),
                ctx, RESIDUAL_KERNELS[T], vec, sizeof(T), length(vec), id, kp, length(kp), 0))
        end
        copts = Ref(COptions(options.reldcost, options.absdcost, options.dstep, options.maxfails, options.maxiters, options.maxtime, Int32(1), Int32(0)))
        cres = Ref{CResult}()
        if callback === NLLSsolver.nullcallback
            check(ctx, ccall((:nlls_optimize, LIB), Cint, (Ptr{Cvoid}, Ptr{COptions}, Ptr{CResult}), ctx, copts, cres))
        else
            # the loop of optimizeinternal! with the callback exactly where the reference calls it (src/optimize.jl:126-128)
            check(ctx, ccall((:nlls_lm_begin, LIB), Cint, (Ptr{Cvoid}, Ptr{COptions}), ctx, copts))
            info = CIterInfo(); conv = Ref{Int64}(0)
            while conv[] == 0
                check(ctx, ccall((:nlls_lm_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{CIterInfo}), ctx, info))
                cost, terminate = callback(info.cost, problem, info, info)::Tuple{Float64, Int}
                check(ctx, ccall((:nlls_lm_advance, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64, Ptr{Int64}), ctx, cost, terminate, conv))
            end
            check(ctx, ccall((:nlls_lm_end, LIB), Cint, (Ptr{Cvoid}, Ptr{CResult}), ctx, cres))
        end
        # variables are optimised in place (src/optimize.jl docstring)
        for (T, idx) in groups
            buf = bufs[T]
            GC.@preserve buf check(ctx, ccall((:nlls_get_variables, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Int64, Int64),
                ctx, vartype(T), 0, buf, length(idx), size(buf, 1)))
            for (k, i) in enumerate(idx); problem.variables[i] = T(buf[:, k]); end
        end
        r = cres[]
        return NLLSResult(r.startcost, r.bestcost, r.timetotal, r.timeinit, r.timecost, r.timegradient, r.timesolver,
                          r.termination, r.niterations, r.costcomputations, r.gradientcomputations, r.linearsolvers)
    finally
        ccall((:nlls_destroy, LIB), Cint, (Ptr{Cvoid},), ctx)
    end
end

end # module
