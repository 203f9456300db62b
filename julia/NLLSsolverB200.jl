# NLLSsolverB200.jl — thin ccall glue between NLLSsolver.jl's problem API and libnlls_b200 (include/nlls_b200.h).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia toolchain (SURVEY.md F1).  Every ccall below is
# mirrored 1:1 by nllssolver.jl_b200/capi.py (ctypes), which the GPU test-suite exercises; keep the two in sync
# (tests/test_capi_cpu.py checks that every symbol bound here is exported by the library and declared in the header).
#
# Usage (a maintainer of NLLSsolver.jl would add this as a package extension):
#     using NLLSsolver, NLLSsolverB200
#     NLLSsolverB200.register_residual(MyReprojectionError, NLLSsolverB200.RES_AFFINE_BA)
#     result = NLLSsolverB200.optimize!(problem, NLLSOptions())          # same signature as NLLSsolver.optimize!
module NLLSsolverB200

using NLLSsolver, StaticArrays, LinearAlgebra

const LIB = get(ENV, "NLLS_B200_LIB", joinpath(@__DIR__, "..", "nllssolver.jl_b200", "libnlls_b200.so"))

# enums of nlls_b200.h
const OK = Cint(0); const ERR_NO_KERNEL = Cint(2)
const VAR_SCALAR = Cint(1); const VAR_EUCLID3 = Cint(3); const VAR_EUCLID6 = Cint(6); const VAR_CONTAMGAUSS = Cint(100); const VAR_PINHOLE = Cint(101)
const RES_AFFINE_BA = Cint(1); const RES_PINHOLE_BA = Cint(2); const RES_ADAPTIVE_OFFSET = Cint(3)
const ROBUST_NONE = Cint(0); const ROBUST_HUBER = Cint(1); const ROBUST_HUBER2O = Cint(2); const ROBUST_GEMANMCCLURE = Cint(3); const ROBUST_SCALED = Cint(16)
const ITER_NEWTON = Int32(0); const ITER_LM = Int32(1); const ITER_DOGLEG = Int32(2); const ITER_GD = Int32(3)

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

# ---- the repo-defined SO(3) / pinhole camera (NLLS_VAR_PINHOLE, NLLS_RES_PINHOLE_BA) ------------------------------------
# The reference ships no such variable (SURVEY F2); this is the Julia definition the device kernel implements, so that the same
# problem also runs on the reference's own CPU path.  Stored: R (column-major), t, f, k1, k2; 9 DoF; update = Exp(w) R on the left.
struct PinholeCamera
    R::SMatrix{3, 3, Float64, 9}; t::SVector{3, Float64}; f::Float64; k1::Float64; k2::Float64
end
NLLSsolver.nvars(::PinholeCamera) = static(9)
function so3exp(w::SVector{3, Float64})
    th = norm(w)
    K = @SMatrix [0.0 -w[3] w[2]; w[3] 0.0 -w[1]; -w[2] w[1] 0.0]
    a, b = th < 1e-5 ? (1 - th^2 / 6, 0.5 - th^2 / 24) : (sin(th) / th, (1 - cos(th)) / th^2)
    return SMatrix{3, 3, Float64}(I) + a * K + b * (K * K)
end
NLLSsolver.update(c::PinholeCamera, x, s=1) =      # same signature as the reference's update(var, updatevec, start=1), src/variable.jl
    PinholeCamera(so3exp(SVector(x[s], x[s+1], x[s+2])) * c.R, c.t + SVector(x[s+3], x[s+4], x[s+5]), c.f + x[s+6], c.k1 + x[s+7], c.k2 + x[s+8])
stored(c::PinholeCamera) = vcat(vec(c.R), c.t, c.f, c.k1, c.k2)
PinholeCamera(v::AbstractVector) = PinholeCamera(SMatrix{3, 3, Float64}(v[1:9]), SVector{3, Float64}(v[10:12]), v[13], v[14], v[15])
# BAL convention: P = R X + t, p = -P.xy / P.z, r = f (1 + k1 |p|^2 + k2 |p|^4) p - z.   Same isbits layout as SimpleError2 (32 B).
struct PinholeReprojectionError <: NLLSsolver.AbstractResidual
    measurement::SVector{2, Float64}
    varind::SVector{2, Int}
end
NLLSsolver.ndeps(::PinholeReprojectionError) = static(2)
NLLSsolver.nres(::PinholeReprojectionError) = static(2)
NLLSsolver.varindices(r::PinholeReprojectionError) = r.varind
NLLSsolver.getvars(r::PinholeReprojectionError, vars::Vector) = (vars[r.varind[1]]::PinholeCamera, vars[r.varind[2]]::NLLSsolver.EuclideanVector{3, Float64})
function NLLSsolver.computeresidual(r::PinholeReprojectionError, c::PinholeCamera, X)
    P = c.R * X + c.t
    p = -SVector(P[1], P[2]) / P[3]
    n2 = p' * p
    return c.f * (1 + n2 * (c.k1 + c.k2 * n2)) * p - r.measurement
end

# ---- registry: concrete residual type => id of its hand-written sm_100a kernel --------------------------------------
const RESIDUAL_KERNELS = Dict{DataType, Cint}(PinholeReprojectionError => RES_PINHOLE_BA)
register_residual(::Type{T}, id) where T = (RESIDUAL_KERNELS[T] = Cint(id))

vartype(::Type{Float64}) = VAR_SCALAR
vartype(::Type{SVector{3, Float64}}) = VAR_EUCLID3
vartype(::Type{SVector{6, Float64}}) = VAR_EUCLID6
vartype(::Type{NLLSsolver.ContaminatedGaussian{Float64}}) = VAR_CONTAMGAUSS
vartype(::Type{PinholeCamera}) = VAR_PINHOLE
vartype(::Type{T}) where T = error("NLLSsolverB200: variable type $T has no registered update kernel (no CPU fallback)")
# stored doubles of a variable, in the order nlls_vartype documents
storedvalues(v::Float64) = (v,)
storedvalues(v::SVector) = Tuple(v)
storedvalues(v::NLLSsolver.ContaminatedGaussian) = (v.invsigma1.val, v.invsigma2.val, v.w.val)
storedvalues(v::PinholeCamera) = Tuple(stored(v))
fromstored(::Type{Float64}, b) = b[1]
fromstored(::Type{T}, b) where T <: SVector = T(b)
fromstored(::Type{NLLSsolver.ContaminatedGaussian{Float64}}, b) = NLLSsolver.ContaminatedGaussian(1 / b[1], 1 / b[2], b[3])
fromstored(::Type{PinholeCamera}, b) = PinholeCamera(b)

# robustkernel(res) => (id, params)                                                    src/robust.jl
kernelspec(::NLLSsolver.NoRobust) = (ROBUST_NONE, Float64[])
kernelspec(k::NLLSsolver.HuberKernel) = (NLLSsolver.dynamic(k.secondorder) ? ROBUST_HUBER2O : ROBUST_HUBER, [Float64(k.width)])
kernelspec(k::NLLSsolver.GemanMcclureKernel) = (ROBUST_GEMANMCCLURE, [sqrt(Float64(k.width_squared))])
function kernelspec(k::NLLSsolver.Scaled)
    id, p = kernelspec(k.robust)
    return (id | ROBUST_SCALED, [isempty(p) ? 0.0 : p[1], Float64(k.height)])
end
kernelspec(k) = error("NLLSsolverB200: robust kernel $(typeof(k)) has no registered device implementation")

iteratorid(it) = it == NLLSsolver.newton ? ITER_NEWTON : it == NLLSsolver.levenbergmarquardt ? ITER_LM :
                 it == NLLSsolver.dogleg ? ITER_DOGLEG : it == NLLSsolver.gradientdescent ? ITER_GD :
                 error("NLLSsolverB200: iterator $it is not implemented")

check(ctx, rc) = rc == OK ? nothing : error("nlls_b200 error $rc: " * unsafe_string(ccall((:nlls_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

# What the reference's own callbacks read from `data` (src/callbacks.jl:39-60,102-107: bestcost, startcost, iternum, starttime,
# linsystem.x) — filled from the library after every iteration, so printoutcallback / storecostscallback work unchanged.
mutable struct LinSystemShim; x::Vector{Float64}; end
mutable struct DataShim
    bestcost::Float64; startcost::Float64; iternum::Int; starttime::UInt64; linsystem::LinSystemShim
end

# ---- problem upload ----------------------------------------------------------------------------------------------------
function upload!(ctx, problem::NLLSProblem)
    # problem.variables: one call per concrete variable type, 1-based positions preserved (src/problem.jl:8,119-121)
    groups = Dict{DataType, Vector{Int}}()
    for (i, v) in enumerate(problem.variables)
        push!(get!(groups, typeof(v), Int[]), i)
    end
    bufs = Dict{DataType, Matrix{Float64}}()
    for (T, idx) in groups
        n = length(storedvalues(problem.variables[idx[1]]))
        buf = Matrix{Float64}(undef, n, length(idx))              # column-major: one variable per column = AoS of n doubles
        for (k, i) in enumerate(idx); buf[:, k] .= storedvalues(problem.variables[i]); end
        bufs[T] = buf
        idx64 = Int64.(idx)
        GC.@preserve buf idx64 check(ctx, ccall((:nlls_set_variables, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Int64}), ctx, vartype(T), buf, length(idx), n, 0, idx64))
    end
    # problem.costs.data[T]: the Vector{T} of isbits structs is handed over as is (src/VectorRepo.jl:3).  The context holds ONE
    # cost set (one residual type): a second non-empty type is an error, not silently dropped.
    nsets = 0
    for (T, vec) in problem.costs.data
        isempty(vec) && continue
        (nsets += 1) == 1 || error("NLLSsolverB200: more than one residual type per problem is not supported by this build")
        haskey(RESIDUAL_KERNELS, T) || error("NLLSsolverB200: residual type $T has no registered sm_100a kernel (no CPU fallback)")
        isbitstype(T) || error("NLLSsolverB200: residual type $T is not isbits")
        id, kp, kernelvar = if RESIDUAL_KERNELS[T] == RES_ADAPTIVE_OFFSET
            # adaptive residuals: varindices(res)[1] is the kernel variable, shared by all costs of the type (src/robustadaptive.jl)
            (ROBUST_NONE, Float64[], Int64(NLLSsolver.varindices(vec[1])[1]))
        else
            (kernelspec(NLLSsolver.robustkernel(vec[1]))..., Int64(0))
        end
        GC.@preserve vec kp check(ctx, ccall((:nlls_set_costs, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cdouble}, Cint, Int64),
            ctx, RESIDUAL_KERNELS[T], vec, sizeof(T), length(vec), id, kp, length(kp), kernelvar))
    end
    return groups, bufs
end

function download!(ctx, problem::NLLSProblem, groups, bufs)
    # variables are optimised in place (src/optimize.jl docstring)
    for (T, idx) in groups
        buf = bufs[T]
        GC.@preserve buf check(ctx, ccall((:nlls_get_variables, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Int64, Int64),
            ctx, vartype(T), 0, buf, length(idx), size(buf, 1)))
        for (k, i) in enumerate(idx); problem.variables[i] = fromstored(T, view(buf, :, k)); end
    end
end

function withcontext(f, device)
    ctxref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nlls_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ctxref, device)
    rc == OK || error("nlls_create failed ($rc): no CUDA device — there is no CPU fallback")
    try
        return f(ctxref[])
    finally
        ccall((:nlls_destroy, LIB), Cint, (Ptr{Cvoid},), ctxref[])
    end
end

coptions(options::NLLSOptions) = Ref(COptions(options.reldcost, options.absdcost, options.dstep, options.maxfails, options.maxiters, options.maxtime,
                                              iteratorid(options.iterator), Int32(0)))
result(r::CResult) = NLLSResult(r.startcost, r.bestcost, r.timetotal, r.timeinit, r.timecost, r.timegradient, r.timesolver,
                                r.termination, r.niterations, r.costcomputations, r.gradientcomputations, r.linearsolvers)

# ---- optimize! -------------------------------------------------------------------------------------------------------
"""
    NLLSsolverB200.optimize!(problem, options=NLLSOptions(), unfixed=nothing, callback=nullcallback; device=0)

Drop-in for `NLLSsolver.optimize!` (src/optimize.jl:5-57).  `unfixed`: `nothing` (all variables), a `BitVector` mask, an index
(that variable only) or a variable type (optimizesingles!-style dispatch is `optimizesingles!` below).  Problems whose residual
types have no registered kernel are rejected with an error.
"""
function optimize!(problem::NLLSProblem, options::NLLSOptions=NLLSOptions(), unfixed=nothing, callback=NLLSsolver.nullcallback; device::Integer=0)
    withcontext(device) do ctx
        groups, bufs = upload!(ctx, problem)
        if unfixed !== nothing
            mask = unfixed isa Integer ? UInt8[i == unfixed for i in 1:length(problem.variables)] :
                   unfixed isa DataType ? UInt8[typeof(v) == unfixed for v in problem.variables] : UInt8.(collect(unfixed))
            GC.@preserve mask check(ctx, ccall((:nlls_set_unfixed, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), ctx, mask, length(mask)))
        end
        copts = coptions(options)
        cres = Ref{CResult}()
        if callback === NLLSsolver.nullcallback
            check(ctx, ccall((:nlls_optimize, LIB), Cint, (Ptr{Cvoid}, Ptr{COptions}, Ptr{CResult}), ctx, copts, cres))
        else
            # the loop of optimizeinternal! with the callback exactly where the reference calls it (src/optimize.jl:126-128)
            check(ctx, ccall((:nlls_lm_begin, LIB), Cint, (Ptr{Cvoid}, Ptr{COptions}), ctx, copts))
            dof = ccall((:nlls_dof, LIB), Int64, (Ptr{Cvoid},), ctx)
            c0 = Ref{Cdouble}(0)
            check(ctx, ccall((:nlls_cost, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), ctx, 0, c0))
            data = DataShim(c0[], c0[], 0, Base.time_ns(), LinSystemShim(zeros(dof)))
            info = CIterInfo(); conv = Ref{Int64}(0)
            while conv[] == 0
                data.iternum += 1
                check(ctx, ccall((:nlls_lm_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{CIterInfo}), ctx, info))
                x = data.linsystem.x
                GC.@preserve x check(ctx, ccall((:nlls_get_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx, x))
                # iteratedata: the trust-region quantity the reference's callbacks print (LevMarData.lambda)
                cost, terminate = callback(info.cost, problem, data, info.lambda)::Tuple{Float64, Int}
                check(ctx, ccall((:nlls_lm_advance, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64, Ptr{Int64}), ctx, cost, terminate, conv))
                data.bestcost = min(data.bestcost, cost)
            end
            check(ctx, ccall((:nlls_lm_end, LIB), Cint, (Ptr{Cvoid}, Ptr{CResult}), ctx, cres))
        end
        download!(ctx, problem, groups, bufs)
        return result(cres[])
    end
end

"""
    NLLSsolverB200.optimizesingles!(problem, options, type; device=0)

`optimizesingles!` (src/optimize.jl:60-76,183-205): every variable of `type` on its own with all others fixed — a batch of
independent small Levenberg-Marquardt solves on the device (registered for the point type of the bundle-adjustment residuals).
"""
function optimizesingles!(problem::NLLSProblem, options::NLLSOptions, type::DataType; device::Integer=0)
    withcontext(device) do ctx
        groups, bufs = upload!(ctx, problem)
        iters = Ref{Int64}(0)
        check(ctx, ccall((:nlls_optimize_singles, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{COptions}, Ptr{Int64}), ctx, vartype(type), coptions(options), iters))
        download!(ctx, problem, groups, bufs)
        return iters[]
    end
end

end # module
