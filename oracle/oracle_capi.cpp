// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points for ctypes (oracle/oracle.py).
#include <cstring>

#include "nlls_oracle.hpp"

using namespace orc;

extern "C" {

// ---- robust kernels
double orc_robustify(int kind, double width, int scaled, double height, double cost) {
    RobustSpec k{kind, width, scaled, height};
    return robustify(k, cost);
}
void orc_robustifydcost(int kind, double width, int scaled, double height, double cost, double* out3) {
    RobustSpec k{kind, width, scaled, height};
    robustifydcost(k, cost, out3[0], out3[1], out3[2]);
}
// ContaminatedGaussian(sigma1, sigma2, w) user-facing constructor; out3 = (invsigma1, invsigma2, w)
void orc_cg_make(double s1, double s2, double w, double* out3) {
    Variable k = make_contaminated_gaussian(s1, s2, w);
    out3[0] = k.v[0]; out3[1] = k.v[1]; out3[2] = k.v[2];
}
static Variable cgvar(const double* p) {
    Variable k; k.type = VT_CONTAMGAUSS; k.nstore = 3; k.ndof = 3; k.v[0] = p[0]; k.v[1] = p[1]; k.v[2] = p[2];
    return k;
}
double orc_cg_robustify(const double* p3, double cost) { return cg_robustify(cgvar(p3), cost); }
void orc_cg_robustifydcost(const double* p3, double cost, double* out3) { cg_robustifydcost(cgvar(p3), cost, out3[0], out3[1], out3[2]); }
void orc_cg_robustifydkernel(const double* p3, double cost, double* val, double* g4, double* H16) {
    cg_robustifydkernel(cgvar(p3), cost, *val, g4, H16);
}

// ---- variables
static Variable mkvar(int type, const double* v, int nstore) {
    Variable var; var.type = type; var.nstore = nstore;
    var.ndof = (type == VT_EUCLID) ? nstore : (type == VT_CONTAMGAUSS ? 3 : 9);
    std::memcpy(var.v, v, sizeof(double) * nstore);
    return var;
}
void orc_update_variable(int type, const double* v, int nstore, const double* x, double* out) {
    Variable r = update(mkvar(type, v, nstore), x);
    std::memcpy(out, r.v, sizeof(double) * nstore);
}
void orc_make_pinhole(const double* rod, const double* t, double f, double k1, double k2, double* out15) {
    Variable c = make_pinhole(rod, t, f, k1, k2);
    std::memcpy(out15, c.v, sizeof(double) * 15);
}

// ---- residual + Jacobian for one cost block (tests)
void orc_resjac(int type, const double* data, int ndeps, const int* vtypes, const int* vnstore, const double* vvals /* ndeps x 16 */,
                int* m, int* P, double* r, double* J) {
    Cost c; c.type = type; c.ndeps = ndeps; std::memcpy(c.data, data, sizeof(double) * 4);
    Variable vars[4]; const Variable* vp[4];
    for (int i = 0; i < ndeps; ++i) { vars[i] = mkvar(vtypes[i], vvals + 16 * i, vnstore[i]); vp[i] = &vars[i]; }
    computeresjac(c, vp, *m, *P, r, J);
}

// ---- problem
void* orc_problem_new() { return new Problem(); }
void orc_problem_free(void* p) { delete (Problem*)p; }
// optimize!(problem, options, unfixed): mask[i] != 0 -> variable i + 1 is optimised; n == 0 clears the mask (all unfixed)
void orc_set_unfixed(void* p, const unsigned char* mask, int64_t n) {
    Problem* pr = (Problem*)p;
    pr->unfixed.assign(mask, mask + n);
    pr->lsready = false;
}
void orc_set_elimination_order(void* p, int mode) { Problem* pr = (Problem*)p; pr->elimination_order = mode; pr->lsready = false; }
int64_t orc_add_variables(void* p, int type, int64_t n, const double* v, int nstore) {
    Problem* pr = (Problem*)p;
    int64_t first = 0;
    for (int64_t i = 0; i < n; ++i) { int64_t idx = pr->addvariable(mkvar(type, v + i * nstore, nstore)); if (i == 0) first = idx; }
    return first;
}
void orc_add_costs(void* p, int type, int64_t n, int ndeps, const int64_t* vi, int ndata, const double* data,
                   int kind, double width, int scaled, double height) {
    Problem* pr = (Problem*)p;
    RobustSpec k{kind, width, scaled, height};
    for (int64_t i = 0; i < n; ++i) {
        Cost c; c.type = type; c.ndeps = ndeps;
        for (int d = 0; d < ndeps; ++d) c.vi[d] = vi[i * ndeps + d];
        for (int d = 0; d < ndata; ++d) c.data[d] = data[i * ndata + d];
        pr->addcost(c, k);
    }
}
int64_t orc_num_variables(void* p) { return (int64_t)((Problem*)p)->variables.size(); }
int64_t orc_variables_len(void* p) { int64_t n = 0; for (auto& v : ((Problem*)p)->variables) n += v.nstore; return n; }
void orc_get_variables(void* p, double* out) {
    for (auto& v : ((Problem*)p)->variables) { std::memcpy(out, v.v, sizeof(double) * v.nstore); out += v.nstore; }
}
void orc_set_variables(void* p, const double* in) {
    for (auto& v : ((Problem*)p)->variables) { std::memcpy(v.v, in, sizeof(double) * v.nstore); in += v.nstore; }
}
double orc_cost(void* p) { return ((Problem*)p)->cost(); }
double orc_linearize(void* p) {
    Problem* pr = (Problem*)p;
    if (!pr->lsready) pr->makesymmvls();
    pr->zero();
    return pr->costgradhess();
}
int64_t orc_dof(void* p) { return ((Problem*)p)->dof; }
int orc_is_sparse(void* p) { return ((Problem*)p)->sparse ? 1 : 0; }
int64_t orc_hess_len(void* p) { Problem* pr = (Problem*)p; return pr->sparse ? (int64_t)pr->A.data.size() : pr->dof * pr->dof; }
// sparse: BSM.data in reference layout; dense: dof x dof column-major with only the lower blocks filled
void orc_get_hess_data(void* p, double* out) {
    Problem* pr = (Problem*)p;
    const std::vector<double>& src = pr->sparse ? pr->A.data : pr->Adense;
    std::memcpy(out, src.data(), sizeof(double) * src.size());
}
void orc_get_grad(void* p, double* out) { Problem* pr = (Problem*)p; std::memcpy(out, pr->b.data(), sizeof(double) * pr->b.size()); }
void orc_get_step(void* p, double* out) { Problem* pr = (Problem*)p; std::memcpy(out, pr->x.data(), sizeof(double) * pr->x.size()); }
// full symmetric dense image of the Hessian (small problems only)
void orc_get_hess_dense(void* p, double* out) {
    Problem* pr = (Problem*)p;
    if (pr->sparse) { std::vector<double> d; pr->A.symmetrifyfull(d); std::memcpy(out, d.data(), sizeof(double) * d.size()); }
    else {
        int64_t n = pr->dof;
        for (int64_t c = 0; c < n; ++c) for (int64_t r = 0; r < n; ++r) out[r + n * c] = (r >= c) ? pr->Adense[(size_t)(r + n * c)] : pr->Adense[(size_t)(c + n * r)];
    }
}
// one damped solve of the current linear system: x = -(H + lambda I)^-1 g  (reference full-system path)
void orc_solve(void* p, double lambda, double* xout) {
    Problem* pr = (Problem*)p;
    pr->gethessian();
    int64_t n = pr->dof;
    if (pr->sparse) {
        std::vector<double> nz = pr->hessval;
        for (int64_t j = 0; j < n; ++j)
            for (int64_t q = pr->hess.colptr[j] - 1; q < pr->hess.colptr[j + 1] - 1; ++q)
                if (pr->hess.rowval[q] - 1 == j) nz[(size_t)q] += lambda;
        ldl_factor_solve(pr->hess, nz.data(), pr->ldl, pr->b.data(), xout);
    } else {
        std::vector<double> A = pr->Adense;
        for (int64_t j = 0; j < n; ++j) A[(size_t)(j + n * j)] += lambda;
        solve_dense((int)n, A.data(), pr->b.data(), xout);
    }
    for (int64_t j = 0; j < n; ++j) xout[j] = -xout[j];
}

void orc_set_callback(void* p, int kind) { ((Problem*)p)->callback_kind = kind; }
void orc_em_optimize(double* k3, const double* sq, int64_t n, int maxiters) { em_optimize(k3, sq, n, maxiters); }

struct orc_options { double reldcost, absdcost, dstep; int64_t maxfails, maxiters; uint64_t maxtime_ns; int64_t callback_terminate; int64_t iterator; };
struct orc_result { double startcost, bestcost, timetotal, timeinit, timecost, timegradient, timesolver;
                    int64_t termination, niterations, costcomputations, gradientcomputations, linearsolvers; };
struct orc_iterrecord { double cost, lambda, maxstep; int64_t ntries; };

int64_t orc_optimize(void* p, const orc_options* o, orc_result* r, orc_iterrecord* trace, int64_t maxtrace) {
    Problem* pr = (Problem*)p;
    Options opt;
    opt.reldcost = o->reldcost; opt.absdcost = o->absdcost; opt.dstep = o->dstep;
    opt.maxfails = o->maxfails; opt.maxiters = o->maxiters; opt.maxtime_ns = o->maxtime_ns;
    opt.callback_terminate = (int)o->callback_terminate;
    opt.iterator = (int)o->iterator;
    std::vector<IterRecord> tr;
    Result res = pr->optimize(opt, &tr);
    r->startcost = res.startcost; r->bestcost = res.bestcost; r->timetotal = res.timetotal; r->timeinit = res.timeinit;
    r->timecost = res.timecost; r->timegradient = res.timegradient; r->timesolver = res.timesolver;
    r->termination = res.termination; r->niterations = res.niterations; r->costcomputations = res.costcomputations;
    r->gradientcomputations = res.gradientcomputations; r->linearsolvers = res.linearsolvers;
    int64_t n = (int64_t)tr.size();
    for (int64_t i = 0; i < n && i < maxtrace; ++i) { trace[i].cost = tr[i].cost; trace[i].lambda = tr[i].lambda; trace[i].maxstep = tr[i].maxstep; trace[i].ntries = tr[i].ntries; }
    return n;
}

// ---- BlockSparseMatrix (golden tests)
void* orc_bsm_new(int64_t nrb, int64_t ncb, const int64_t* colptr, const int64_t* rowval, const int* rbs, const int* cbs) {
    BSM* b = new BSM();
    std::vector<int64_t> cp(colptr, colptr + nrb + 1), rv(rowval, rowval + (colptr[nrb] - 1));
    b->build(cp, rv, std::vector<int>(rbs, rbs + nrb), std::vector<int>(cbs, cbs + ncb));
    return b;
}
void orc_bsm_free(void* b) { delete (BSM*)b; }
int64_t orc_bsm_nnz(void* b) { return (int64_t)((BSM*)b)->data.size(); }
int64_t orc_bsm_m(void* b) { return ((BSM*)b)->m; }
int64_t orc_bsm_n(void* b) { return ((BSM*)b)->n; }
int64_t orc_bsm_start(void* b, int64_t i, int64_t j) { return ((BSM*)b)->start(i, j); }
void orc_bsm_setblock(void* b, int64_t i, int64_t j, const double* vals, int64_t n) {
    BSM* m = (BSM*)b; int64_t st = m->start(i, j) - 1;
    for (int64_t k = 0; k < n; ++k) m->data[(size_t)(st + k)] = vals[k];
}
void orc_bsm_data(void* b, double* out) { BSM* m = (BSM*)b; std::memcpy(out, m->data.data(), sizeof(double) * m->data.size()); }
void orc_bsm_todense(void* b, double* out) { std::vector<double> d; ((BSM*)b)->todense(d); std::memcpy(out, d.data(), sizeof(double) * d.size()); }
void orc_bsm_symmetrifyfull(void* b, double* out) { std::vector<double> d; ((BSM*)b)->symmetrifyfull(d); std::memcpy(out, d.data(), sizeof(double) * d.size()); }
void orc_bsm_uniformscaling(void* b, double k) { ((BSM*)b)->uniformscaling(k); }
// returns nnz; fills colptr (n+1), rowval, values gathered from data
int64_t orc_bsm_sparse(void* b, int symmetrify, int64_t* colptr, int64_t* rowval, double* vals, int64_t cap) {
    BSM* m = (BSM*)b;
    CSCIndex idx = makesparseindices(*m, symmetrify != 0);
    int64_t nz = (int64_t)idx.nzval.size();
    if (nz > cap) return -nz;
    for (size_t i = 0; i < idx.colptr.size(); ++i) colptr[i] = idx.colptr[i];
    for (int64_t i = 0; i < nz; ++i) { rowval[i] = idx.rowval[(size_t)i]; vals[i] = m->data[(size_t)(idx.nzval[(size_t)i] - 1)]; }
    return nz;
}

// ---- utils / solvers
int64_t orc_rle(const int64_t* sorted, int64_t n, int64_t* out) {
    std::vector<int64_t> r = runlengthencodesortedints(std::vector<int64_t>(sorted, sorted + n));
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int64_t)r.size();
}
int orc_solve_dense(int n, const double* A, const double* b, double* x) { return solve_dense(n, A, b, x); }
// full symmetric CSC (1-based colptr/rowval), natural ordering
int orc_solve_sparse(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nz, const double* b, double* x) {
    CSCIndex A; A.m = A.n = n; A.colptr.assign(colptr, colptr + n + 1); A.rowval.assign(rowval, rowval + (colptr[n] - 1));
    std::vector<int64_t> perm((size_t)n); for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = i;
    LDLSymbolic sym; ldl_analyze(A, perm, sym);
    return ldl_factor_solve(A, nz, sym, b, x) ? 0 : 1;
}
double orc_fast_bAb_dense(int n, const double* A, const double* b) { return fast_bAb_dense(n, A, b); }
double orc_fast_bAb_sparse(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nz, const double* b) {
    CSCIndex A; A.m = A.n = n; A.colptr.assign(colptr, colptr + n + 1); A.rowval.assign(rowval, rowval + (colptr[n] - 1));
    return fast_bAb_sparse(A, nz, b);
}

}  // extern "C"

extern "C" int64_t orc_optimizesingles(void* p, const orc_options* o, const int64_t* indices, int64_t n) {
    Problem* pr = (Problem*)p;
    Options opt;
    opt.reldcost = o->reldcost; opt.absdcost = o->absdcost; opt.dstep = o->dstep;
    opt.maxfails = o->maxfails; opt.maxiters = o->maxiters; opt.maxtime_ns = o->maxtime_ns;
    opt.callback_terminate = 0;
    opt.iterator = (int)o->iterator;
    return pr->optimizesingles(opt, std::vector<int64_t>(indices, indices + n));
}
