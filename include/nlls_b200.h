/* nlls_b200.h — C ABI of the B200-native Levenberg-Marquardt inner loop for NLLSsolver.jl.
 *
 * The reference (ojwoodford/NLLSsolver.jl v4.0.3, pure Julia) has no FFI; its seam is the Julia dispatch
 * between the outer optimiser (src/optimize.jl, src/iterators.jl) and the objective / linear-system layer
 * (src/cost.jl, src/residual.jl, src/linearsystem.jl, src/linearsolver.jl).  Each entry point below names the
 * reference call it replaces.  The Julia glue (julia/NLLSsolverB200.jl) binds exactly these symbols with
 * ccall; the same symbols are driven from Python through ctypes (nllssolver.jl_b200/capi.py).
 *
 * Conventions
 *  - every function returns an int status (NLLS_OK == 0); nothing unwinds through the ABI;
 *  - numeric trouble (NaN/Inf cost or step, singular system) is NOT an error: it is reported through the
 *    termination word exactly as the reference does (src/optimize.jl:151-161);
 *  - all host pointers are plain arrays owned by the caller and are copied during the call;
 *  - variable indices are the reference's 1-based positions in problem.variables (src/problem.jl:119-121);
 *  - a context is not thread-safe: one in-flight call per context; calls are synchronous on return;
 *  - there is NO CPU fallback: a residual type / robust kernel without a registered sm_100a kernel is
 *    rejected with NLLS_ERR_NO_KERNEL.
 */
#ifndef NLLS_B200_H
#define NLLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nlls_ctx nlls_ctx;

enum nlls_status {
    NLLS_OK = 0,
    NLLS_ERR_INVALID = 1,     /* bad argument / call order */
    NLLS_ERR_NO_KERNEL = 2,   /* residual type, variable type or robust kernel has no registered CUDA kernel */
    NLLS_ERR_UNSUPPORTED = 3, /* structurally valid but outside what this build handles (see nlls_last_error) */
    NLLS_ERR_CUDA = 4,
    NLLS_ERR_NCCL = 5,
    NLLS_ERR_NO_DEVICE = 6
};

/* Variable types (src/variable.jl, src/robustadaptive.jl). */
enum nlls_vartype {
    NLLS_VAR_EUCLID3 = 3,          /* EuclideanVector{3,Float64}: 3 stored, 3 DoF, update v + x   src/variable.jl:8-10 */
    NLLS_VAR_EUCLID6 = 6,          /* EuclideanVector{6,Float64}                                                      */
    NLLS_VAR_SCALAR = 1,           /* Float64                                                      src/variable.jl:4-5 */
    NLLS_VAR_CONTAMGAUSS = 100,    /* ContaminatedGaussian: (invsigma1, invsigma2, w) stored, 3 DoF  src/robustadaptive.jl:3-22 */
    NLLS_VAR_PINHOLE = 101         /* repo-defined camera: R (9, column-major), t (3), f, k1, k2 = 15 stored, 9 DoF   */
};

/* Residual types with a registered fused kernel. */
enum nlls_restype {
    NLLS_RES_AFFINE_BA = 1,        /* SimpleError2{2,Float64,EV6,EV3} + affine generatemeasurement  src/residual.jl:4-14, test/optimizeba.jl:4.
                                      AoS element (32 B): double z[2]; int64 varind[2] = (camera, point). */
    NLLS_RES_PINHOLE_BA = 2,       /* same AoS element; camera is NLLS_VAR_PINHOLE (repo-defined, parity-unpinned)   */
    NLLS_RES_ADAPTIVE_OFFSET = 3   /* AbstractAdaptiveResidual r = mean - data  examples/adaptivekernel.jl:9-18, test/adaptivecost.jl:3-13.
                                      AoS element (16 B): double data; int64 varind (the mean variable); kernel variable index given separately. */
};

/* Fixed robust kernels (src/robust.jl:7-77): robustkernel(res) of the residual type. */
enum nlls_robust {
    NLLS_ROBUST_NONE = 0,          /* NoRobust                        kparams: -            */
    NLLS_ROBUST_HUBER = 1,         /* HuberKernel(w)                  kparams: w            */
    NLLS_ROBUST_HUBER2O = 2,       /* Huber2oKernel(w)                kparams: w            */
    NLLS_ROBUST_GEMANMCCLURE = 3,  /* GemanMcclureKernel(w)           kparams: w            */
    NLLS_ROBUST_SCALED = 16        /* OR-ed in: Scaled(kernel, height) kparams: w, height   */
};

/* NLLSOptions (src/structs.jl:22-35). iterator: all four of src/structs.jl:4 (Dogleg and gradient descent: BA residuals, one rank). */
enum nlls_iterator { NLLS_ITER_NEWTON = 0, NLLS_ITER_LM = 1, NLLS_ITER_DOGLEG = 2, NLLS_ITER_GD = 3 };
typedef struct nlls_options {
    double reldcost;
    double absdcost;
    double dstep;
    int64_t maxfails;
    int64_t maxiters;
    uint64_t maxtime_ns;
    int32_t iterator;
    int32_t reserved;
} nlls_options;

/* NLLSResult (src/structs.jl:37-50); times in seconds. */
typedef struct nlls_result {
    double startcost, bestcost, timetotal, timeinit, timecost, timegradient, timesolver;
    int64_t termination, niterations, costcomputations, gradientcomputations, linearsolvers;
} nlls_result;

/* What a per-iteration Julia callback needs (src/callbacks.jl:39-60,102-107). */
typedef struct nlls_iterinfo {
    double cost;      /* value returned by iterate!                                          */
    double lambda;    /* LevMarData.lambda after the iteration                               */
    double maxstep;   /* maximum(abs, x)                                                     */
    double stepnorm;  /* norm(x)                                                             */
    int64_t ntries;   /* inner LM tries (= linear solves) of this outer iteration            */
    int64_t accepted; /* 1: !(cost > bestcost); 0: returned only because max|x| < dstep      */
} nlls_iterinfo;

/* ---- lifetime ---------------------------------------------------------------------------------------- */
int nlls_create(nlls_ctx** ctx, int device);            /* one context per NLLSProblem; owns streams + device buffers */
int nlls_destroy(nlls_ctx* ctx);
const char* nlls_last_error(nlls_ctx* ctx);             /* Julia exceptions / @assert messages                         */
int nlls_version(void);

/* ---- multi-GPU (one process per GPU; residual blocks sharded by point) -------------------- */
int nlls_comm_unique_id(void* id128);                   /* rank 0 creates; the host broadcasts the 128 bytes           */
int nlls_comm_init(nlls_ctx* ctx, int rank, int nranks, const void* id128);

/* ---- problem definition ------------------------------------------------------------------------------ */
/* problem.variables (src/problem.jl:8), one call per concrete variable type.  `aos` holds n variables of
 * `stride` doubles each (stored values, see nlls_vartype).  indices == NULL: variables occupy the 1-based
 * positions first_index .. first_index+n-1, otherwise indices[i] is the position of variable i (ascending).
 * Calling it again with the same (type, n, indices) only refreshes the values (problem.variables changed). */
int nlls_set_variables(nlls_ctx* ctx, int vartype, const double* aos, int64_t n, int64_t stride,
                       int64_t first_index, const int64_t* indices);
/* problem.costs.data[T] (src/VectorRepo.jl:3): the AoS image of Vector{T}; see nlls_restype for the element
 * layout.  robust = nlls_robust id (| NLLS_ROBUST_SCALED), kparams as listed there.  kernel_var is the
 * 1-based index of the adaptive kernel variable for NLLS_RES_ADAPTIVE_OFFSET (ignored otherwise).
 * In a multi-rank run each rank passes only the costs of the points it owns. */
int nlls_set_costs(nlls_ctx* ctx, int restype, const void* aos, int64_t stride_bytes, int64_t n,
                   int robust, const double* kparams, int nkparams, int64_t kernel_var);
/* A further Vector{T} of problem.costs.data: costs of the same residual struct (same restype) whose type carries a different
 * robustkernel().  cost, gradient and Hessian are summed over all cost sets (src/cost.jl:54, src/VectorRepo.jl:64-69).  Up to 8 sets;
 * nlls_set_costs starts over with one set. */
int nlls_add_costs(nlls_ctx* ctx, int restype, const void* aos, int64_t stride_bytes, int64_t n,
                   int robust, const double* kparams, int nkparams);
/* optimize!(problem, options, unfixed) (src/optimize.jl:5-20): unfixed[i] != 0 <=> variable i + 1 is optimised, for i < n; variables
 * beyond n are unfixed; n == 0 (or unfixed == NULL) clears the mask.  Fixed variables keep their values, contribute to the cost, and are
 * left out of linsystem.b / x (nlls_dof, nlls_get_gradient, nlls_get_step); the Hessian read-back is not available under a mask. */
int nlls_set_unfixed(nlls_ctx* ctx, const uint8_t* unfixed, int64_t n);
/* optimizesingles!(problem, options, type) (src/optimize.jl:60-76,183-205): every variable of `vartype` on its own, all others fixed,
 * over the costs that depend on it — a batch of independent small LM solves.  Registered for the point type (NLLS_VAR_EUCLID3) of the
 * bundle-adjustment residuals.  iterations (may be NULL): summed iteration count. */
int nlls_optimize_singles(nlls_ctx* ctx, int vartype, const nlls_options* opts, int64_t* iterations);
/* makesymmvls (src/linearsystem.jl:91-124) + reordercostsforschur! (src/problem.jl:177-199): builds the
 * point-major collision-free scatter layout and uploads everything.  Called implicitly if needed. */
int nlls_prepare(nlls_ctx* ctx);

/* ---- the six L1/L2 operations the outer optimiser calls ---------------------------------- */
/* zero! + costgradhess!(linsystem, variables, costs)   src/optimize.jl:118,168-169, src/cost.jl:29-54      */
int nlls_linearize(nlls_ctx* ctx, double* cost);
/* cost(vars, costs)  src/cost.jl:11 ; which: 0 = variables, 1 = varnext, 2 = varbest                      */
int nlls_cost(nlls_ctx* ctx, int which, double* cost);
/* optimize(kernel::ContaminatedGaussian, squarederrors, maxiters) (src/robustadaptive.jl:48-73): Expectation-Maximisation refit of the
 * adaptive kernel variable of buffer `which` (as nlls_cost) from the squared residuals at that buffer's means — what the EM callback of
 * test/adaptivecost.jl:15-25 does between iterations (followed there by nlls_cost(ctx, 1, ..) for the new cost). */
int nlls_adaptive_em(nlls_ctx* ctx, int which, int maxiters);
/* uniformscaling!(H, lambda) + solve! + negate!   src/iterators.jl:149-152 ;  x = -(H + lambda I)^-1 g    */
int nlls_solve(nlls_ctx* ctx, double lambda);
/* update!(varnext, variables, linsystem)   src/iterators.jl:155, src/linearsystem.jl:206-213              */
int nlls_update(nlls_ctx* ctx);
/* The outer loop of optimizeinternal! (src/optimize.jl:109-180), split where the reference calls the user callback:
 *   nlls_lm_begin    setupiterator + first costgradhess!                                   (:109-121)
 *   nlls_lm_iterate  iterate!(::LevMarData, ...) — damp/solve/update/cost until accepted    (:126, src/iterators.jl:139-172)
 *                    (options.iterator == NLLS_ITER_NEWTON: iterate!(::NewtonData, ...) — one undamped try, :17-27)
 *   [the caller runs callback(cost, ...) -> (cost, terminate) here]                        (:128)
 *   nlls_lm_advance  best/fail bookkeeping, variables <-> varnext swap, termination word, re-linearisation when it is 0
 *                    (:130-171); `terminate` is OR-ed in << 16
 *   nlls_lm_end      restore varbest if needed, fill NLLSResult                            (:173-178)                    */
int nlls_lm_begin(nlls_ctx* ctx, const nlls_options* opts);
int nlls_lm_iterate(nlls_ctx* ctx, nlls_iterinfo* info);
int nlls_lm_advance(nlls_ctx* ctx, double cost, int64_t terminate, int64_t* converged);
/* nlls_lm_iterate + nlls_lm_advance(info->cost, 0): one outer iteration with the null callback (src/callbacks.jl:20), as nlls_optimize
 * runs it.  Multi-rank: nlls_lm_advance combines `terminate` over the ranks (maximum; one 8-byte all-reduce per iteration, so that one
 * rank's callback stops every rank at the same iteration); this call has no flag to exchange and no collective of its own. */
int nlls_lm_step(nlls_ctx* ctx, nlls_iterinfo* info, int64_t* converged);
int nlls_lm_end(nlls_ctx* ctx, nlls_result* result);
/* optimizeinternal! with nullcallback   src/optimize.jl:109-180                                           */
int nlls_optimize(nlls_ctx* ctx, const nlls_options* opts, nlls_result* result);

/* ---- read-back in the reference's layout (debug / parity / callbacks) --------------------------------- */
int nlls_get_variables(nlls_ctx* ctx, int vartype, int which, double* aos, int64_t n, int64_t stride);
int64_t nlls_dof(nlls_ctx* ctx);                        /* length of linsystem.b / x                                   */
int nlls_get_gradient(nlls_ctx* ctx, double* b);        /* linsystem.b  (+grad, src/linearsystem.jl:166), variable order */
int nlls_get_step(nlls_ctx* ctx, double* x);            /* linsystem.x                                                 */
int64_t nlls_hessian_len(nlls_ctx* ctx);                /* length(A.data) of the BlockSparseMatrix                     */
/* BlockSparseMatrix.data in reference order: block rows by variable index, blocks ascending by block column,
 * each block column-major (src/BlockSparseMatrix.jl:30-47,102-105).                */
int nlls_get_hessian_blocks(nlls_ctx* ctx, double* data);
/* 1-based start of every stored block, as (row block, column block, start) triples, row-major order.      */
int64_t nlls_hessian_nblocks(nlls_ctx* ctx);
int nlls_get_hessian_index(nlls_ctx* ctx, int64_t* rowblock, int64_t* colblock, int64_t* start);

/* ---- measurement hooks (bench.py) --------------------------------------------------------------------- */
/* Runs `reps` back-to-back linearisations (or cost evaluations / LM tries) on the context's stream bracketed
 * by CUDA events on that stream and returns the average milliseconds per call of the named kernel group.  */
enum nlls_timed { NLLS_TIME_LINEARIZE = 0, NLLS_TIME_LIN_POINT = 1, NLLS_TIME_LIN_CAM = 2, NLLS_TIME_COST = 3,
                  NLLS_TIME_SCHUR = 4, NLLS_TIME_SOLVE_REDUCED = 5, NLLS_TIME_BACKSUB = 6, NLLS_TIME_TRY = 7,
                  NLLS_TIME_MEMSET_H = 8 /* write-only probe: cudaMemset over the point rows of H (destroys H until the next linearise) */,
                  NLLS_TIME_LIN_LOOP = 9 /* the re-linearisation as the LM loop runs it: point pass + finalize of the camera partials the
                                            accepted try's cost evaluation left behind */ };
int nlls_time_kernels(nlls_ctx* ctx, int which, int reps, int flush_l2, double* ms_per_call);
/* CUDA events on the context's stream around an arbitrary sequence of calls (bench.py's timed region). */
int nlls_timer_start(nlls_ctx* ctx);
int nlls_timer_stop(nlls_ctx* ctx, double* ms);
int64_t nlls_kernel_launches(nlls_ctx* ctx);            /* kernels launched by this context so far                     */
int nlls_algorithmic_bytes(nlls_ctx* ctx, int which, double* bytes);  /* SURVEY §8d figures for this problem         */
/* FP64 operations per launch of NLLS_TIME_SCHUR / NLLS_TIME_SOLVE_REDUCED (multiply and add counted separately) */
int nlls_algorithmic_flops(nlls_ctx* ctx, int which, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* NLLS_B200_H */
