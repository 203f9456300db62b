# baseline/run_reference.jl — times the UNMODIFIED reference (ojwoodford/NLLSsolver.jl) on the bench workload's construction rules.
#
# Not runnable in the build image (no Julia toolchain, SURVEY F1); bench.py's `--impl reference` arm times the C++ restatement
# (oracle/) instead and will prefer this script if a `julia` binary with NLLSsolver installed ever appears on PATH
# (SURVEY §8d: "Ship baseline/run_reference.jl that times optimize! with NLLSResult.time{gradient,cost,solver}").
#
#   julia --threads=auto baseline/run_reference.jl [ncam npt nobs] [iters]
#
# Prints one JSON line: residual blocks/s through full LM iterations + the reference's own timers (src/structs.jl:42-44).
using NLLSsolver, StaticArrays, Random, LinearAlgebra, Printf

# test/optimizeba.jl:4 — the affine camera of the reference's own bundle-adjustment test
NLLSsolver.generatemeasurement(pose::SVector{6, Float64}, X::SVector{3, Float64}) = SVector(dot(@view(pose[1:3]), X), dot(@view(pose[4:6]), X))
NLLSsolver.robustkernel(::SimpleError2{2, Float64, SVector{6, Float64}, SVector{3, Float64}}) = HuberKernel(0.03)

function create_bal_shaped(ncam, npt, nobs; noise=0.01, outlier_frac=0.02, outlier_scale=50.0, rng=MersenneTwister(0))
    # the construction rules of nllssolver.jl_b200/synthetic.py: point l has centre camera linspace(2, ncam-1, npt)[l] and is seen
    # by its k_l nearest cameras, k_l >= 2, sum k_l = nobs (the RNG stream differs from numpy's: shapes match, values do not)
    problem = NLLSProblem(Union{SVector{6, Float64}, SVector{3, Float64}}, SimpleError2{2, Float64, SVector{6, Float64}, SVector{3, Float64}})
    for _ in 1:ncam; addvariable!(problem, SVector{6}(randn(rng, 6)) + SVector(1.0, 0, 0, 0, 1, 0)); end
    for _ in 1:npt; addvariable!(problem, SVector{3}(rand(rng, 3)) + SVector(-0.5, -0.5, 10.0)); end
    k = fill(2, npt); extra = nobs - 2npt
    while extra > 0; l = rand(rng, 1:npt); if k[l] < ncam; k[l] += 1; extra -= 1; end; end
    centres = range(2, ncam - 1, length=npt)
    for l in 1:npt
        c0 = clamp(ceil(Int, centres[l] - k[l] / 2), 1, ncam - k[l] + 1)
        for c in c0:c0 + k[l] - 1
            z = NLLSsolver.generatemeasurement(problem.variables[c], problem.variables[ncam + l]) + noise * SVector{2}(randn(rng, 2))
            rand(rng) < outlier_frac && (z += noise * outlier_scale * SVector{2}(randn(rng, 2)))
            addcost!(problem, SimpleError2{2, Float64, SVector{6, Float64}, SVector{3, Float64}}(z, c, ncam + l))
        end
    end
    for i in 1:ncam; problem.variables[i] += 1e-3 * SVector{6}(randn(rng, 6)); end
    for i in ncam + 1:ncam + npt; problem.variables[i] += 1e-3 * SVector{3}(randn(rng, 3)); end
    return problem
end

function main()
    ncam, npt, nobs = length(ARGS) >= 3 ? parse.(Int, ARGS[1:3]) : (1778, 993923, 5001946)
    iters = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 2
    problem = create_bal_shaped(ncam, npt, nobs)
    optimize!(deepcopy(problem), NLLSOptions(maxiters=1, maxtime=1e9))          # compile
    t0 = time_ns()
    res = optimize!(problem, NLLSOptions(maxiters=iters, maxtime=1e9, iterator=NLLSsolver.levenbergmarquardt))
    dt = (time_ns() - t0) * 1e-9 - res.timeinit
    @printf("{\"impl\": \"reference\", \"kind\": \"reference\", \"value\": %.6e, \"unit\": \"residual blocks/s\", \"lm_iters_per_sec\": %.6e, \"threads\": %d, \"timegradient\": %.4f, \"timesolver\": %.4f, \"timecost\": %.4f, \"niterations\": %d, \"bestcost\": %.12e}\n",
            nobs * res.niterations / dt, res.niterations / dt, Threads.nthreads(), res.timegradient, res.timesolver, res.timecost, res.niterations, res.bestcost)
end
main()
