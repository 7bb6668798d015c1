// host_model.cpp -- host-side layer around the device pivot loop: model -> equality standard form ->
// phase-1/phase-2 problems -> ellp_b200_{primal,dual}_solve_with_initial, plus the MPS reader.
//
// This is index-heavy O(problem size) preprocessing that the reference also does once per solve on
// the host; it mirrors (reference paths):
//   Option<StandardForm>::from(Problem)   src/standard_form.rs:78-191
//   PrimalPhase1 / PrimalPhase2           src/solvers/primal/primal_problem.rs:80-291
//   DualPhase1 / DualPhase2               src/solvers/dual/dual_problem.rs:89-404
//   {Primal,Dual}SimplexSolver::solve     primal_simplex_solver.rs:32-93, dual_simplex_solver.rs:33-108
//   parse_mps                             src/parse_mps.rs:23-546 (deterministic file order)
// The dense factorizations used here (pivoted Householder QR for the rank test, partial/full pivot
// LU for the starting bases) follow the algorithms of nalgebra, the reference's un-vendored
// dependency (Cargo.toml:16).  All pivoting on the device goes through the C ABI.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "engine.hpp"

namespace {

constexpr double kTol = 0.0000000001;  // src/util.rs:1
const double kInf = std::numeric_limits<double>::infinity();

struct RefPanic : std::runtime_error { using std::runtime_error::runtime_error; };
struct RefError : std::runtime_error { using std::runtime_error::runtime_error; };
struct AbiFailure { int code; };

inline void require(bool ok, const char* what) { if (!ok) throw RefPanic(what); }

struct Limits { uint8_t kind; double lo, hi; };

struct ModelVar { int64_t id; double cost; Limits lim; };
struct ModelRow { std::vector<std::pair<int64_t, double>> terms; uint8_t op; double rhs; };
struct Model {
    std::vector<ModelVar> vars;
    std::vector<ModelRow> rows;
    // flat copies handed out through ellp_b200_model_desc
    std::vector<double> f_obj, f_lb, f_ub, f_coef, f_rhs;
    std::vector<uint8_t> f_kind, f_op;
    std::vector<int64_t> f_id, f_col;
    std::vector<int32_t> f_ptr;
};

Model model_from_desc(const ellp_problem_desc* p) {
    Model mdl;
    mdl.vars.resize(p->nvars);
    for (int i = 0; i < p->nvars; ++i) {
        ModelVar& v = mdl.vars[i];
        v.id = p->var_id ? p->var_id[i] : i;
        v.cost = p->obj[i];
        v.lim.kind = p->kind[i];
        v.lim.lo = p->lb ? p->lb[i] : 0.;
        v.lim.hi = p->ub ? p->ub[i] : 0.;
        if (v.lim.kind == ELLP_FIXED) v.lim.hi = v.lim.lo;
    }
    mdl.rows.resize(p->ncons);
    for (int r = 0; r < p->ncons; ++r) {
        ModelRow& row = mdl.rows[r];
        row.op = p->op[r];
        row.rhs = p->rhs[r];
        for (int k = p->row_ptr[r]; k < p->row_ptr[r + 1]; ++k) row.terms.emplace_back(p->col_id[k], p->coef[k]);
    }
    return mdl;
}

// column-major dense block
struct Dense {
    int nr = 0, nc = 0;
    std::vector<double> v;
    Dense() {}
    Dense(int r, int c) : nr(r), nc(c), v((size_t)r * c, 0.) {}
    double& at(int i, int j) { return v[(size_t)j * nr + i]; }
    double at(int i, int j) const { return v[(size_t)j * nr + i]; }
    double* colp(int j) { return v.data() + (size_t)j * nr; }
    const double* colp(int j) const { return v.data() + (size_t)j * nr; }
};

using SwapList = std::vector<std::pair<int, int>>;
template <class T> void apply_swaps(const SwapList& s, std::vector<T>& x) { for (auto& p : s) std::swap(x[p.first], x[p.second]); }
template <class T> void undo_swaps(const SwapList& s, std::vector<T>& x) { for (size_t k = s.size(); k-- > 0;) std::swap(x[s[k].first], x[s[k].second]); }

double sgn1(double x) { return std::isnan(x) ? x : (std::signbit(x) ? -1. : 1.); }  // f64::signum

double inner(const double* a, const double* b, int n) { double s = 0.; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }

// first largest |entry| of the trailing block [k.., k..], scanning columns then rows
void trailing_argmax(const Dense& M, int k, int* pr, int* pc) {
    double best = -1.;
    *pr = k; *pc = k;
    for (int j = k; j < M.nc; ++j) {
        const double* cj = M.colp(j);
        for (int i = k; i < M.nr; ++i) {
            const double a = std::fabs(cj[i]);
            if (a > best) { best = a; *pr = i; *pc = j; }
        }
    }
}

// Gaussian elimination of column k below the diagonal (reciprocal-scaled multipliers, no fused ops)
void eliminate_below(Dense& M, int k, double pivot) {
    const double inv = 1. / pivot;
    double* ck = M.colp(k);
    for (int i = k + 1; i < M.nr; ++i) ck[i] *= inv;
    for (int j = k + 1; j < M.nc; ++j) {
        double* cj = M.colp(j);
        const double f = -cj[k];
        for (int i = k + 1; i < M.nr; ++i) cj[i] = f * ck[i] + cj[i];
    }
}

// partial-pivot LU (row swaps recorded in order); pivot = first largest |a| in the column
struct RowPivotLU {
    Dense f;
    SwapList swaps;
    explicit RowPivotLU(Dense M) : f(std::move(M)) {
        const int steps = std::min(f.nr, f.nc);
        for (int k = 0; k < steps; ++k) {
            int p = k;
            double best = std::fabs(f.at(k, k));
            for (int i = k + 1; i < f.nr; ++i) { const double a = std::fabs(f.at(i, k)); if (a > best) { best = a; p = i; } }
            const double piv = f.at(p, k);
            if (piv == 0.) continue;
            if (p != k) { swaps.emplace_back(k, p); for (int j = 0; j < f.nc; ++j) std::swap(f.at(k, j), f.at(p, j)); }
            eliminate_below(f, k, piv);
        }
    }
    int steps() const { return std::min(f.nr, f.nc); }
    bool solve(std::vector<double>& b) const {  // A x = b
        const int n = f.nr;
        apply_swaps(swaps, b);
        for (int k = 0; k < n; ++k) { const double t = b[k]; const double* ck = f.colp(k); for (int i = k + 1; i < n; ++i) b[i] = -t * ck[i] + b[i]; }
        for (int k = n - 1; k >= 0; --k) {
            const double dg = f.at(k, k);
            if (dg == 0.) return false;
            const double t = b[k] / dg;
            b[k] = t;
            const double* ck = f.colp(k);
            for (int i = 0; i < k; ++i) b[i] = -t * ck[i] + b[i];
        }
        return true;
    }
    bool solve_transposed(std::vector<double>& b) const {  // A^T x = b
        const int n = f.nr;
        for (int k = 0; k < n; ++k) { const double dg = f.at(k, k); if (dg == 0.) return false; b[k] = (b[k] - inner(f.colp(k), b.data(), k)) / dg; }
        for (int k = n - 1; k >= 0; --k) b[k] = b[k] - inner(f.colp(k) + k + 1, b.data() + k + 1, n - k - 1);
        undo_swaps(swaps, b);
        return true;
    }
};

struct HostStdForm {
    int m = 0, n = 0;
    Dense A;
    std::vector<double> c, b;
    std::vector<Limits> lim;
    Model model;
};

struct HostPoint {
    std::vector<double> x, y, d;
    std::vector<int32_t> B, N;
    std::vector<uint8_t> Ns;
};

double sf_obj(const HostStdForm& sf, const std::vector<double>& x) { return inner(sf.c.data(), x.data(), (int)std::min(sf.c.size(), x.size())); }

double sf_dual_obj(const HostStdForm& sf, const std::vector<double>& y, const std::vector<double>& d) {  // standard_form.rs:52-68
    require(d.size() == sf.lim.size(), "assertion failed: d.len() == self.bounds.len()");
    double o = inner(sf.b.data(), y.data(), (int)sf.b.size());
    for (size_t i = 0; i < sf.lim.size(); ++i) {
        const Limits& L = sf.lim[i];
        if (L.kind == ELLP_LOWER) o += L.lo * d[i];
        else if (L.kind == ELLP_UPPER) o += L.hi * d[i];
        else if (L.kind == ELLP_TWOSIDED) o += (d[i] > 0.) ? L.lo * d[i] : L.hi * d[i];
        else if (L.kind == ELLP_FIXED) o += L.lo * d[i];
    }
    return o;
}

// standard_form.rs:78-191.  false => the reference returns None (Infeasible).
bool standardize(const Model& mdl, HostStdForm& out) {
    const int nv = (int)mdl.vars.size(), nr = (int)mdl.rows.size();
    int slacks = 0;
    for (auto& r : mdl.rows) slacks += (r.op != ELLP_EQ);
    const int total = nv + slacks;
    Dense A(nr, total);
    std::vector<double> c(total, 0.), b(nr, 0.);
    std::vector<Limits> lim(total, Limits{ELLP_LOWER, 0., 0.});
    std::unordered_map<int64_t, int> where;
    for (int j = 0; j < nv; ++j) { c[j] = mdl.vars[j].cost; lim[j] = mdl.vars[j].lim; where[mdl.vars[j].id] = j; }
    int slack_col = total > 0 ? total - 1 : 0;  // slack of the first inequality row is the LAST column (:115)
    for (int i = 0; i < nr; ++i) {
        const ModelRow& r = mdl.rows[i];
        b[i] = r.rhs;
        if (r.terms.empty() && b[i] != 0.) return false;  // :120-122
        for (auto& t : r.terms) {
            auto it = where.find(t.first);
            require(it != where.end(), "called `Option::unwrap()` on a `None` value (unknown variable id)");
            A.at(i, it->second) = t.second;
        }
        if (r.op != ELLP_EQ) {
            require(slack_col >= 0 && nv + slacks > 0, "attempt to subtract with overflow");
            A.at(i, slack_col) = (r.op == ELLP_LTE) ? 1. : -1.;
            --slack_col;
        }
    }
    require(!(nv == 0 && slacks > 0), "attempt to subtract with overflow (standard_form.rs:135)");
    // rank test + row order: Householder QR of A^T with column pivoting on the largest |entry| (:142)
    Dense T(total, nr);
    for (int i = 0; i < nr; ++i) for (int j = 0; j < total; ++j) T.at(j, i) = A.at(i, j);
    const int steps = std::min(T.nr, T.nc);
    std::vector<double> rdiag(steps, 0.);
    SwapList colswaps;
    for (int k = 0; k < steps; ++k) {
        int pr, pc;
        trailing_argmax(T, k, &pr, &pc);
        if (pc != k) { for (int i = 0; i < T.nr; ++i) std::swap(T.at(i, k), T.at(i, pc)); colswaps.emplace_back(k, pc); }
        double* w = T.colp(k) + k;
        const int len = T.nr - k;
        double nrm2 = 0.;
        for (int i = 0; i < len; ++i) nrm2 += w[i] * w[i];
        const double nrm = std::sqrt(nrm2);
        const double head = std::fabs(w[0]);
        const double signed_nrm = sgn1(w[0]) * nrm;
        const double scale2 = (nrm2 + head * nrm) * 2.;
        w[0] += signed_nrm;
        if (scale2 != 0.) {
            const double sc = std::sqrt(scale2);
            for (int i = 0; i < len; ++i) w[i] /= sc;
            double again = 0.;
            for (int i = 0; i < len; ++i) again += w[i] * w[i];
            again = std::sqrt(again);
            if (again != 0.) for (int i = 0; i < len; ++i) w[i] /= again;
            const double sg = sgn1(-signed_nrm);
            for (int j = k + 1; j < T.nc; ++j) {
                double* cj = T.colp(j) + k;
                const double f = inner(w, cj, len) * (-2. * sg);
                for (int i = 0; i < len; ++i) cj[i] = f * w[i] + sg * cj[i];
            }
            rdiag[k] = std::fabs(signed_nrm);
        } else {
            rdiag[k] = std::fabs(signed_nrm);
        }
    }
    undo_swaps(colswaps, b);  // :144
    for (double& r : rdiag) if (r < kTol) r = 0.;  // :148-154
    const bool trivial = steps > 0 && nr > 0 && rdiag[0] < kTol && std::fabs(b[0]) < kTol;  // :159 (R non-empty)
    if (!trivial) for (double r : rdiag) if (r == 0.) return false;  // :161-163 R^T solve hits a zero diagonal
    apply_swaps(colswaps, b);  // :165
    int keep = steps;
    for (int k = 0; k < steps; ++k) if (rdiag[k] < kTol) { keep = k; break; }  // :170-174
    std::vector<int> order(nr);
    for (int i = 0; i < nr; ++i) order[i] = i;
    apply_swaps(colswaps, order);  // :176-178
    order.resize(keep);
    out.m = keep;
    out.n = total;
    out.A = Dense(keep, total);
    out.b.assign(keep, 0.);
    for (int k = 0; k < keep; ++k) {
        for (int j = 0; j < total; ++j) out.A.at(k, j) = A.at(order[k], j);
        out.b[k] = b[order[k]];
    }
    out.c = std::move(c);
    out.lim = std::move(lim);
    out.model = mdl;
    return true;
}

std::vector<double> residual_rhs(const HostStdForm& sf, const std::vector<double>& v) {  // b - A v
    std::vector<double> r = sf.b;
    for (int j = 0; j < sf.A.nc; ++j) { const double vj = v[j]; const double* cj = sf.A.colp(j); for (int i = 0; i < sf.m; ++i) r[i] -= cj[i] * vj; }
    return r;
}

Dense pick_columns(const Dense& A, const std::vector<int32_t>& idx) {
    Dense out(A.nr, (int)idx.size());
    for (size_t k = 0; k < idx.size(); ++k) std::memcpy(out.colp((int)k), A.colp(idx[k]), sizeof(double) * A.nr);
    return out;
}

void widen(HostStdForm& sf, int new_cols) {
    Dense W(sf.A.nr, new_cols);
    std::memcpy(W.v.data(), sf.A.v.data(), sizeof(double) * sf.A.v.size());
    sf.A = std::move(W);
    sf.n = new_cols;
}

struct PrimalStage { HostStdForm sf; HostPoint pt; std::vector<int> artificial; };

// primal_problem.rs:80-261
bool build_primal_phase1(const Model& mdl, PrimalStage& ps) {
    if (!standardize(mdl, ps.sf)) return false;
    HostStdForm& sf = ps.sf;
    const int n = sf.n, m = sf.m;
    std::vector<double> v(n, 0.);
    auto& N = ps.pt.N; auto& Ns = ps.pt.Ns; auto& B = ps.pt.B;
    for (int j = 0; j < n; ++j) {  // :95-135
        const Limits& L = sf.lim[j];
        if (L.kind == ELLP_FREE) continue;
        v[j] = (L.kind == ELLP_UPPER) ? L.hi : L.lo;
        N.push_back(j);
        Ns.push_back(L.kind == ELLP_UPPER ? ELLP_NB_UPPER : ELLP_NB_LOWER);
    }
    std::fill(sf.c.begin(), sf.c.end(), 0.);  // :137-141
    sf.c.resize(n + m, 1.);
    std::vector<int32_t> freev;
    for (int j = 0; j < n; ++j) if (sf.lim[j].kind == ELLP_FREE) freev.push_back(j);
    if (!freev.empty() && m > 0 && n > 0) {  // :158-233 crash basis from the free columns
        Dense F = pick_columns(sf.A, freev);
        SwapList rs, cs;
        const int steps = std::min(F.nr, F.nc);
        for (int k = 0; k < steps; ++k) {  // full-pivot LU (:162)
            int pr, pc;
            trailing_argmax(F, k, &pr, &pc);
            const double piv = F.at(pr, pc);
            if (piv == 0.) break;
            if (pc != k) { for (int i = 0; i < F.nr; ++i) std::swap(F.at(i, k), F.at(i, pc)); cs.emplace_back(k, pc); }
            if (pr != k) { rs.emplace_back(k, pr); for (int j = 0; j < F.nc; ++j) std::swap(F.at(k, j), F.at(pr, j)); }
            eliminate_below(F, k, piv);
        }
        int rank = (int)freev.size();  // :167-175
        for (int k = 0; k < steps; ++k) if (std::fabs(F.at(k, k)) < kTol) { rank = k; break; }
        apply_swaps(cs, freev);  // :180
        require(rank <= m, "index out of bounds (free-variable crash rank > rows, primal_problem.rs:199)");
        for (int k = 0; k < rank; ++k) B.push_back(freev[k]);
        for (size_t k = rank; k < freev.size(); ++k) { sf.lim[freev[k]] = Limits{ELLP_FIXED, 0., 0.}; N.push_back(freev[k]); Ns.push_back(ELLP_NB_LOWER); }
        std::vector<double> t = residual_rhs(sf, v);  // :196-210
        apply_swaps(rs, t);
        t.resize(rank);
        for (int k = 0; k < rank; ++k) { const double s = t[k]; for (int i = k + 1; i < rank; ++i) t[i] = -s * F.at(i, k) + t[i]; }
        for (int k = rank - 1; k >= 0; --k) {
            const double dg = F.at(k, k);
            require(dg != 0., "called `Option::unwrap()` on a `None` value (solve_upper_triangular)");
            const double s = t[k] / dg;
            t[k] = s;
            for (int i = 0; i < k; ++i) t[i] = -s * F.at(i, k) + t[i];
        }
        for (int k = 0; k < rank; ++k) v[freev[k]] = t[k];  // :214-216
        std::vector<int> rows(m);
        for (int i = 0; i < m; ++i) rows[i] = i;
        apply_swaps(rs, rows);  // :218-220
        rows.erase(rows.begin(), rows.begin() + rank);
        std::vector<double> t2 = residual_rhs(sf, v);  // :222
        v.resize(n + m, 0.);
        widen(sf, n + (int)rows.size());  // :225
        int col = sf.n - 1;
        for (int i : rows) { v[col] = std::fabs(t2[i]); sf.A.at(i, col) = sgn1(t2[i]); B.push_back(col); --col; }  // :228-233
    } else {  // :234-246 one artificial per row
        std::vector<double> t = residual_rhs(sf, v);
        v.resize(n + m, 0.);
        widen(sf, n + m);
        for (int i = 0; i < m; ++i) { v[n + i] = std::fabs(t[i]); sf.A.at(i, n + i) = sgn1(t[i]); B.push_back(n + i); }
    }
    for (int k = 0; k < m; ++k) { ps.artificial.push_back((int)sf.lim.size()); sf.lim.push_back(Limits{ELLP_LOWER, 0., 0.}); }  // :248-253
    ps.pt.x = std::move(v);
    return true;
}

// primal_problem.rs:263-291
void primal_to_phase2(PrimalStage& ps) {
    HostStdForm& sf = ps.sf;
    for (int j : ps.artificial) { sf.c[j] = 0.; sf.lim[j] = Limits{ELLP_FIXED, 0., 0.}; }
    for (size_t j = 0; j < sf.model.vars.size(); ++j) { sf.c[j] = sf.model.vars[j].cost; sf.lim[j] = sf.model.vars[j].lim; }
    for (size_t k = 0; k < ps.pt.N.size(); ++k) if (sf.lim[ps.pt.N[k]].kind == ELLP_FREE) ps.pt.Ns[k] = ELLP_NB_FREE;
}

struct DualStage { HostStdForm sf; HostPoint pt; HostStdForm original; };

// dual_problem.rs:89-256
bool build_dual_phase1(const Model& mdl, DualStage& ds) {
    if (!standardize(mdl, ds.original)) return false;
    const HostStdForm& o = ds.original;
    Model boxed;  // :96-134 auxiliary problem: Free->[-1,1], Lower->[0,1], Upper->[-1,0], boxed/fixed columns dropped, b = 0
    std::vector<char> kept(o.n, 0);
    for (int j = 0; j < o.n; ++j) {
        Limits L;
        if (o.lim[j].kind == ELLP_FREE) L = Limits{ELLP_TWOSIDED, -1., 1.};
        else if (o.lim[j].kind == ELLP_LOWER) L = Limits{ELLP_TWOSIDED, 0., 1.};
        else if (o.lim[j].kind == ELLP_UPPER) L = Limits{ELLP_TWOSIDED, -1., 0.};
        else continue;
        kept[j] = 1;
        boxed.vars.push_back(ModelVar{(int64_t)j, o.c[j], L});
    }
    for (int i = 0; i < o.m; ++i) {
        ModelRow r;
        r.op = ELLP_EQ;
        r.rhs = 0.;
        for (int j = 0; j < o.n; ++j) if (kept[j]) r.terms.emplace_back((int64_t)j, o.A.at(i, j));
        if (!r.terms.empty()) boxed.rows.push_back(std::move(r));
    }
    if (!standardize(boxed, ds.sf)) return false;  // :136-139
    HostStdForm& sf = ds.sf;
    Dense T(sf.n, sf.m);  // :141 LU of A^T: its row permutation picks the starting basis
    for (int i = 0; i < sf.m; ++i) for (int j = 0; j < sf.n; ++j) T.at(j, i) = sf.A.at(i, j);
    RowPivotLU tlu(std::move(T));
    for (int k = 0; k < tlu.steps(); ++k) require(!(std::fabs(tlu.f.at(k, k)) < kTol), "should always have a basis available");  // :143-147
    std::vector<int32_t> perm(sf.n);
    for (int j = 0; j < sf.n; ++j) perm[j] = j;
    apply_swaps(tlu.swaps, perm);  // :150-152
    HostPoint& pt = ds.pt;
    pt.B.assign(perm.begin(), perm.begin() + sf.m);
    pt.N.assign(perm.begin() + sf.m, perm.end());
    pt.Ns.assign(pt.N.size(), ELLP_NB_LOWER);
    if (!pt.B.empty()) {  // :166-226
        RowPivotLU blu(pick_columns(sf.A, pt.B));
        std::vector<double> y(pt.B.size());
        for (size_t i = 0; i < pt.B.size(); ++i) y[i] = sf.c[pt.B[i]];
        require(blu.solve_transposed(y), "called `Option::unwrap()` on a `None` value (tr_solve)");
        std::vector<double> d(sf.n);
        for (int j = 0; j < sf.n; ++j) d[j] = sf.c[j] - inner(sf.A.colp(j), y.data(), sf.m);  // :172
        std::vector<double> x(sf.lim.size(), 0.);
        for (size_t k = 0; k < pt.N.size(); ++k) {  // :178-204
            const int j = pt.N[k];
            const Limits& L = sf.lim[j];
            if (L.kind == ELLP_TWOSIDED) { if (d[j] >= 0.) { x[j] = L.lo; pt.Ns[k] = ELLP_NB_LOWER; } else { x[j] = L.hi; pt.Ns[k] = ELLP_NB_UPPER; } }
            else if (L.kind == ELLP_FIXED) { x[j] = L.lo; pt.Ns[k] = (d[j] >= 0.) ? ELLP_NB_LOWER : ELLP_NB_UPPER; }
            else throw RefPanic("bounds should always be fixed or two-sided");
        }
        std::vector<double> t = residual_rhs(sf, x);  // :207-208
        require(blu.solve(t), "called `Option::unwrap()` on a `None` value (lu.solve)");
        for (size_t i = 0; i < pt.B.size(); ++i) x[pt.B[i]] = t[i];
        pt.y = std::move(y);
        pt.d = std::move(d);
        pt.x = std::move(x);
    } else {  // :227-254
        require(pt.N.size() == sf.lim.size(), "assertion `left == right` failed (N.len(), bounds.len())");
        pt.x.assign(pt.N.size(), 0.);
        for (size_t k = 0; k < pt.N.size(); ++k) {
            const Limits& L = sf.lim[pt.N[k]];
            require(L.kind == ELLP_TWOSIDED || L.kind == ELLP_FIXED, "bounds should always be fixed or two-sided");
            pt.x[pt.N[k]] = L.lo;
        }
        pt.y.clear();
        pt.d = sf.c;
    }
    return true;
}

struct DualStage2 { HostStdForm sf; HostPoint pt; };

// dual_problem.rs:258-404
void dual_to_phase2(DualStage& ds, DualStage2& out) {
    const Model& boxed = ds.sf.model;
    out.sf = std::move(ds.original);
    HostStdForm& sf = out.sf;
    HostPoint& pt = out.pt;
    std::vector<char> basic(sf.n, 0);
    for (int b : ds.pt.B) { const int j = (int)boxed.vars[b].id; basic[j] = 1; pt.B.push_back(j); }  // :264-273
    if (!pt.B.empty()) {
        // quirk Q17: phase 1 drops rows without kept variables, so B can have fewer entries than the standard form has rows; the
        // reference then panics inside nalgebra (non-square triangular solve / tr_mul dimension mismatch at :283).  Checked BEFORE
        // the solve: a non-square A_B must never reach the triangular solves.
        require((int)pt.B.size() == sf.m, "Matrix multiplication dimensions mismatch (dual_problem.rs:283)");
        RowPivotLU blu(pick_columns(sf.A, pt.B));
        std::vector<double> y(pt.B.size());
        for (size_t i = 0; i < pt.B.size(); ++i) y[i] = sf.c[pt.B[i]];
        require(blu.solve_transposed(y), "called `Option::unwrap()` on a `None` value (tr_solve)");
        std::vector<double> d(sf.n);
        for (int j = 0; j < sf.n; ++j) d[j] = sf.c[j] - inner(sf.A.colp(j), y.data(), sf.m);
        std::vector<double> xN;
        for (int j = 0; j < sf.n; ++j) {  // :285-324
            if (basic[j]) continue;
            const Limits& L = sf.lim[j];
            const double dj = d[j];
            double xv; uint8_t side;
            if (L.kind == ELLP_FREE) { require(std::fabs(dj) < kTol, "assertion failed: d_i.abs() < EPS"); xv = 0.; side = ELLP_NB_FREE; }
            else if (L.kind == ELLP_LOWER) { require(dj > -kTol, "assertion failed: d_i > -EPS"); xv = L.lo; side = ELLP_NB_LOWER; }
            else if (L.kind == ELLP_UPPER) { require(dj < kTol, "assertion failed: d_i < EPS"); xv = L.hi; side = ELLP_NB_UPPER; }
            else if (L.kind == ELLP_TWOSIDED) { if (dj >= 0.) { xv = L.lo; side = ELLP_NB_LOWER; } else { xv = L.hi; side = ELLP_NB_UPPER; } }
            else { xv = L.lo; side = ELLP_NB_LOWER; }
            xN.push_back(xv); pt.N.push_back(j); pt.Ns.push_back(side);
        }
        std::vector<double> t = sf.b;  // :326-328
        for (size_t k = 0; k < pt.N.size(); ++k) { const double* cj = sf.A.colp(pt.N[k]); for (int i = 0; i < sf.m; ++i) t[i] -= cj[i] * xN[k]; }
        require(blu.solve(t), "called `Option::unwrap()` on a `None` value (lu.solve)");
        pt.x.assign(sf.n, 0.);
        for (size_t i = 0; i < pt.B.size(); ++i) pt.x[pt.B[i]] = t[i];
        for (size_t k = 0; k < pt.N.size(); ++k) pt.x[pt.N[k]] = xN[k];
        pt.y = std::move(y);
        pt.d = std::move(d);
    } else {  // :351-402
        pt.x.assign(sf.n, 0.);
        for (int j = 0; j < sf.n; ++j) {
            if (basic[j]) continue;
            const Limits& L = sf.lim[j];
            uint8_t side = ELLP_NB_LOWER;
            if (L.kind == ELLP_FREE) { pt.x[j] = 0.; side = ELLP_NB_FREE; }
            else if (L.kind == ELLP_UPPER) { pt.x[j] = L.hi; side = ELLP_NB_UPPER; }
            else pt.x[j] = L.lo;
            pt.N.push_back(j); pt.Ns.push_back(side);
        }
        pt.y.clear();
        pt.d = sf.c;
    }
}

// ---- calling the device loop through the C ABI ---------------------------------------------------
struct AbiArrays { std::vector<uint8_t> kind; std::vector<double> lo, hi; };

int device_phase(ellp_b200_ctx* ctx, int solver, HostStdForm& sf, HostPoint& pt, const ellp_opts* base, int phase_tag,
                 ellp_solution* sol, int64_t* trace_off) {
    AbiArrays a;
    const size_t nb = sf.lim.size();
    a.kind.resize(nb); a.lo.resize(nb); a.hi.resize(nb);
    for (size_t j = 0; j < nb; ++j) { a.kind[j] = sf.lim[j].kind; a.lo[j] = sf.lim[j].lo; a.hi[j] = sf.lim[j].hi; }
    ellp_std_form f{sf.m, sf.n, sf.A.v.data(), sf.c.data(), sf.b.data(), a.kind.data(), a.lo.data(), a.hi.data()};
    // x may be longer than n (primal phase 1 keeps n+m entries, quirk Q16): the loop only touches the first n
    const int nB = (int)pt.B.size(), nN = (int)pt.N.size();
    if (sf.m == 0) { pt.N.resize(sf.n); pt.Ns.resize(sf.n); }  // solve_trivial_problem rewrites N
    if ((int)pt.x.size() < sf.n) pt.x.resize(sf.n, 0.);
    if (solver == ELLP_DUAL) { if ((int)pt.y.size() < sf.m) pt.y.resize(sf.m, 0.); if ((int)pt.d.size() < sf.n) pt.d.resize(sf.n, 0.); }
    ellp_point p{pt.x.data(), pt.B.data(), pt.N.data(), pt.Ns.data(), pt.y.data(), pt.d.data(), nB, nN};
    ellp_opts o = *base;
    o.phase_tag = phase_tag;
    if (base->trace) { o.trace = base->trace + *trace_off; o.trace_cap = std::max<int64_t>(0, base->trace_cap - *trace_off); if (o.trace_cap == 0) o.trace = nullptr; }
    ellp_result r;
    const int rc = (solver == ELLP_PRIMAL) ? ellp_b200_primal_solve_with_initial(ctx, &f, &p, &o, &r)
                                           : ellp_b200_dual_solve_with_initial(ctx, &f, &p, &o, &r);
    if (rc != ELLP_OK) throw AbiFailure{rc};
    if (sf.m == 0) { pt.N.resize(p.nN); pt.Ns.resize(p.nN); }
    sol->iters[phase_tag] += r.iters;
    sol->launches += r.launches;
    sol->ms_device += r.ms_device;
    if (base->trace) *trace_off += std::min<int64_t>(r.trace_len, o.trace_cap);
    sol->trace_len += r.trace_len;
    return r.status;
}

// primal_simplex_solver.rs:32-93
// Latency path of PrimalSimplexSolver::solve for netlib-sized problems: the whole two-phase solve (PrimalPhase1::from,
// phase 1, verdict, PrimalPhase2::from, phase 2 -- primal_simplex_solver.rs:32-93) is ONE launch of the shared-memory kernel
// K6 (batch.cuh) with a batch of one LP, instead of ~7 launches per pivot.  Taken when ellp_opts::engine is AUTO, the
// standard form has no Free variable and fits shared memory; any outcome that needs the reference's panic text or the
// MaxIter objective falls through to the general path, which reproduces it.  Pivot for pivot the same solve
// (tests/test_gpu_parity.py::test_batch_kernel_single_lp_follows_oracle_pivot_for_pivot).
bool try_small_primal(ellp_b200_ctx* ctx, const Model& mdl, const ellp_opts* o, ellp_solution* sol, int64_t* toff, bool* infeasible_std) {
    *infeasible_std = false;
    if (o->engine != ELLP_ENGINE_AUTO) return false;
    if (o->pricing != ELLP_PRICE_REFERENCE || o->ratio != ELLP_RATIO_REFERENCE) return false;  // K6 implements the reference's rules only
    HostStdForm sf;
    if (!standardize(mdl, sf)) { *infeasible_std = true; return true; }
    const int m = sf.m, n = sf.n;
    if (m < 1 || n < m || (int)sf.lim.size() != n) return false;
    if ((double)(m + 2) * (double)n * 8.0 + 64.0 * (n + m) > 200.0 * 1024.0) return false;
    std::vector<uint8_t> kind(n);
    std::vector<double> lo(n), hi(n);
    for (int j = 0; j < n; ++j) {
        if (sf.lim[j].kind == ELLP_FREE) return false;
        kind[j] = sf.lim[j].kind; lo[j] = sf.lim[j].lo; hi[j] = sf.lim[j].hi;
    }
    ellp_batch bt{1, m, n, sf.A.v.data(), sf.c.data(), sf.b.data(), kind.data(), lo.data(), hi.data()};
    std::vector<double> x((size_t)n + m, 0.);
    int32_t status = -1, iters[2] = {0, 0}, err = 0, tlen = 0;
    double obj = 0.;
    ellp_batch_result res{};
    res.status = &status; res.obj = &obj; res.x = x.data(); res.iters = iters; res.err = &err; res.trace_len = &tlen;
    const int64_t tcap = o->trace ? std::max<int64_t>(0, o->trace_cap - *toff) : 0;
    if (tcap > 0) { res.trace = o->trace + *toff; res.trace_cap = (int32_t)std::min<int64_t>(tcap, 1 << 20); }
    if (ellp_b200_primal_solve_batch(ctx, &bt, o, &res) != ELLP_OK) return false;
    if (err != 0 || status < 0 || status == ELLP_MAXITER) return false;
    sol->status = status;
    sol->iters[0] = (uint64_t)iters[0];
    sol->iters[1] = (uint64_t)iters[1];
    sol->launches += res.launches;
    sol->ms_device += res.ms_device;
    if (tcap > 0) { *toff += std::min<int64_t>(tlen, res.trace_cap); sol->trace_len += tlen; }
    else sol->trace_len += iters[0] + iters[1];
    if (status == ELLP_OPTIMAL) {
        sol->obj = obj;
        if (sol->x) std::copy(x.begin(), x.begin() + mdl.vars.size(), sol->x);
    }
    return true;
}

void run_primal(ellp_b200_ctx* ctx, const Model& mdl, const ellp_opts* o, ellp_solution* sol, int64_t* toff) {
    {
        bool infeasible_std = false;
        if (try_small_primal(ctx, mdl, o, sol, toff, &infeasible_std)) {
            if (infeasible_std) sol->status = ELLP_INFEASIBLE;
            return;
        }
    }
    PrimalStage ps;
    if (!build_primal_phase1(mdl, ps)) { sol->status = ELLP_INFEASIBLE; return; }
    int st = device_phase(ctx, ELLP_PRIMAL, ps.sf, ps.pt, o, 0, sol, toff);
    if (st == ELLP_OPTIMAL) {
        const double obj = sf_obj(ps.sf, ps.pt.x);
        require(obj > -kTol, "assertion failed: obj > -EPS");
        if (!(obj < kTol)) { sol->status = ELLP_INFEASIBLE; return; }
        primal_to_phase2(ps);
    } else if (st == ELLP_INFEASIBLE) { sol->status = ELLP_INFEASIBLE; return; }
    else if (st == ELLP_UNBOUNDED) throw RefPanic("primal phase 1 should never be unbounded");
    else { sol->status = ELLP_MAXITER; sol->obj = kInf; return; }
    st = device_phase(ctx, ELLP_PRIMAL, ps.sf, ps.pt, o, 1, sol, toff);
    if (st == ELLP_OPTIMAL) {
        sol->status = ELLP_OPTIMAL;
        sol->obj = sf_obj(ps.sf, ps.pt.x);
        if (sol->x) std::copy(ps.pt.x.begin(), ps.pt.x.begin() + ps.sf.model.vars.size(), sol->x);
    } else if (st == ELLP_INFEASIBLE) throw RefPanic("primal phase 2 should never be infeasible");
    else if (st == ELLP_UNBOUNDED) sol->status = ELLP_UNBOUNDED;
    else { sol->status = ELLP_MAXITER; sol->obj = sf_obj(ps.sf, ps.pt.x); }
}

// dual_simplex_solver.rs:33-108
void run_dual(ellp_b200_ctx* ctx, const Model& mdl, const ellp_opts* o, ellp_solution* sol, int64_t* toff) {
    DualStage ds;
    if (!build_dual_phase1(mdl, ds)) { sol->status = ELLP_INFEASIBLE; return; }
    int st = device_phase(ctx, ELLP_DUAL, ds.sf, ds.pt, o, 2, sol, toff);
    if (st == ELLP_OPTIMAL) {
        const double obj = sf_dual_obj(ds.sf, ds.pt.y, ds.pt.d);
        require(obj < kTol, "assertion failed: obj < EPS");
        if (!(obj > -kTol)) {  // :50-67 dual infeasible: let the DEFAULT primal solver decide
            ellp_opts po = *o;
            po.max_iter = 1000;
            run_primal(ctx, ds.original.model, &po, sol, toff);
            sol->used_primal_fallback = 1;
            require(sol->status != ELLP_OPTIMAL, "assertion failed: matches!(result, Infeasible | Unbounded | MaxIter)");
            return;
        }
    } else if (st == ELLP_INFEASIBLE) throw RefPanic("dual phase 1 should never be infeasible");
    else if (st == ELLP_UNBOUNDED) throw RefPanic("dual phase 1 should never be unbounded");
    else { sol->status = ELLP_MAXITER; sol->obj = kInf; return; }
    DualStage2 d2;
    dual_to_phase2(ds, d2);
    st = device_phase(ctx, ELLP_DUAL, d2.sf, d2.pt, o, 3, sol, toff);
    if (st == ELLP_OPTIMAL) {
        sol->status = ELLP_OPTIMAL;
        sol->obj = sf_obj(d2.sf, d2.pt.x);
        const size_t nv = d2.sf.model.vars.size();
        require(nv <= d2.pt.x.size(), "Matrix slicing out of bounds");
        if (sol->x) std::copy(d2.pt.x.begin(), d2.pt.x.begin() + nv, sol->x);
    } else if (st == ELLP_INFEASIBLE) sol->status = ELLP_INFEASIBLE;
    else if (st == ELLP_UNBOUNDED) throw RefPanic("dual phase 2 should never return unbounded");
    else { sol->status = ELLP_MAXITER; sol->obj = sf_dual_obj(d2.sf, d2.pt.y, d2.pt.d); }
}

void fail_msg(char* err256, const char* msg) { if (err256) std::snprintf(err256, 256, "%s", msg); }

}  // namespace

// set from engine.cu's ctx; declared here to record host-layer messages
extern "C" const char* ellp_b200_last_error(const ellp_b200_ctx* ctx);
extern "C" void ellp_b200_set_error_(ellp_b200_ctx* ctx, const char* msg);

struct ellp_b200_model { Model m; };
struct ellp_b200_stage { HostStdForm sf; HostPoint pt; };

extern "C" {

int ellp_b200_solve(ellp_b200_ctx* ctx, const ellp_problem_desc* p, int solver, const ellp_opts* o, ellp_solution* sol) {
    if (!ctx || !p || !o || !sol) return ELLP_E_ARG;
    double* xout = sol->x;
    std::memset(sol, 0, sizeof(*sol));
    sol->x = xout;
    int64_t toff = 0;
    try {
        Model mdl = model_from_desc(p);
        if (solver == ELLP_PRIMAL) run_primal(ctx, mdl, o, sol, &toff);
        else run_dual(ctx, mdl, o, sol, &toff);
        return ELLP_OK;
    } catch (const AbiFailure& f) {
        return f.code;  // message already recorded by the device layer
    } catch (const RefError& e) {
        ellp_b200_set_error_(ctx, e.what());
        return ELLP_E_ELLP;
    } catch (const RefPanic& e) {
        ellp_b200_set_error_(ctx, e.what());
        return ELLP_E_PANIC;
    } catch (const std::exception& e) {
        ellp_b200_set_error_(ctx, e.what());
        return ELLP_E_PANIC;
    }
}

int ellp_b200_stage_new(const ellp_problem_desc* p, int which, ellp_b200_stage** out, int* infeasible, char* err256) {
    if (!p || !out || !infeasible) return ELLP_E_ARG;
    *out = nullptr;
    *infeasible = 0;
    try {
        Model mdl = model_from_desc(p);
        auto st = std::make_unique<ellp_b200_stage>();
        bool ok;
        if (which == 0) ok = standardize(mdl, st->sf);
        else if (which == 1) { PrimalStage ps; ok = build_primal_phase1(mdl, ps); if (ok) { st->sf = std::move(ps.sf); st->pt = std::move(ps.pt); } }
        else { DualStage ds; ok = build_dual_phase1(mdl, ds); if (ok) { st->sf = std::move(ds.sf); st->pt = std::move(ds.pt); } }
        if (!ok) { *infeasible = 1; return ELLP_OK; }
        *out = st.release();
        return ELLP_OK;
    } catch (const std::exception& e) {
        fail_msg(err256, e.what());
        return ELLP_E_PANIC;
    }
}

void ellp_b200_stage_free(ellp_b200_stage* s) { delete s; }

void ellp_b200_stage_dims(const ellp_b200_stage* s, int32_t* m, int32_t* n, int32_t* nx, int32_t* nB, int32_t* nN,
                          int32_t* len_c, int32_t* len_bounds) {
    *m = s->sf.m; *n = s->sf.n; *nx = (int32_t)s->pt.x.size(); *nB = (int32_t)s->pt.B.size(); *nN = (int32_t)s->pt.N.size();
    *len_c = (int32_t)s->sf.c.size(); *len_bounds = (int32_t)s->sf.lim.size();
}

void ellp_b200_stage_copy(const ellp_b200_stage* s, double* A, double* c, double* b, uint8_t* kind, double* lb, double* ub,
                          double* x, int32_t* B, int32_t* N, uint8_t* N_side, double* y, double* d) {
    if (A) std::copy(s->sf.A.v.begin(), s->sf.A.v.end(), A);
    if (c) std::copy(s->sf.c.begin(), s->sf.c.end(), c);
    if (b) std::copy(s->sf.b.begin(), s->sf.b.end(), b);
    for (size_t j = 0; j < s->sf.lim.size(); ++j) {
        if (kind) kind[j] = s->sf.lim[j].kind;
        if (lb) lb[j] = s->sf.lim[j].lo;
        if (ub) ub[j] = s->sf.lim[j].hi;
    }
    if (x) std::copy(s->pt.x.begin(), s->pt.x.end(), x);
    if (B) std::copy(s->pt.B.begin(), s->pt.B.end(), B);
    if (N) std::copy(s->pt.N.begin(), s->pt.N.end(), N);
    if (N_side) std::copy(s->pt.Ns.begin(), s->pt.Ns.end(), N_side);
    if (y) std::copy(s->pt.y.begin(), s->pt.y.end(), y);
    if (d) std::copy(s->pt.d.begin(), s->pt.d.end(), d);
}

// ---- MPS reader (src/parse_mps.rs) -----------------------------------------------------------------
// Same accepted syntax and error texts as the reference: sections NAME / ROWS / COLUMNS / RHS /
// [BOUNDS] / ENDATA; exactly one (row, value) pair per COLUMNS line (:290-295); RHS lines with 2 or 3
// tokens (:370-372); bounds UP / LO / FR only (:496-506).  Unlike the reference (HashMap iteration,
// :29,:41) variables and constraints keep FILE order, which makes pivot sequences reproducible.
int ellp_b200_parse_mps(const char* text, ellp_b200_model** out, char* err256) {
    if (!text || !out) return ELLP_E_ARG;
    *out = nullptr;
    struct RowRec { bool objective; uint8_t op; std::vector<std::pair<int, double>> terms; bool has_rhs; double rhs; };
    struct ColRec { std::string name; double cost; bool has_lim; Limits lim; };
    std::vector<std::string> lines;
    {
        std::stringstream ss(text);
        std::string ln;
        while (std::getline(ss, ln, '\n')) {
            if (ln.find_first_not_of(" \t\r") == std::string::npos) continue;
            lines.push_back(ln);
        }
    }
    auto toks = [](const std::string& s) { std::vector<std::string> t; std::stringstream ss(s); std::string w; while (ss >> w) t.push_back(w); return t; };
    auto trim = [](const std::string& s) { const size_t a = s.find_first_not_of(" \t\r"); const size_t b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
    auto starts = [&](const std::string& s, const char* p) { const std::string t = s.substr(std::min(s.size(), s.find_first_not_of(" \t"))); return t.rfind(p, 0) == 0; };
    auto bad = [&](const std::string& m) { fail_msg(err256, ("MPS parsing error. " + m).c_str()); return ELLP_E_ELLP; };
    auto num = [](const std::string& s, double* v) { char* e = nullptr; *v = std::strtod(s.c_str(), &e); return e && *e == 0 && e != s.c_str(); };
    size_t at = 0;
    if (lines.empty()) return bad("could not find NAME line");
    {
        auto t = toks(lines[at]);
        if (t.size() < 2 || t[0] != "NAME") return bad("could not find name in NAME line: " + lines[at]);
        ++at;
    }
    if (at >= lines.size()) return bad("could not find ROWS line");
    if (trim(lines[at]) != "ROWS") return bad("expected 'ROWS', found '" + trim(lines[at]) + "'");
    ++at;
    std::vector<RowRec> rows;
    std::vector<std::string> row_names;
    std::unordered_map<std::string, int> row_of;
    for (; at < lines.size() && !starts(lines[at], "COLUMNS"); ++at) {
        auto t = toks(lines[at]);
        if (t.empty()) return bad("expected a row type character in this line: " + lines[at]);
        RowRec r{false, ELLP_EQ, {}, false, 0.};
        if (t[0] == "L") r.op = ELLP_LTE; else if (t[0] == "G") r.op = ELLP_GTE; else if (t[0] == "E") r.op = ELLP_EQ;
        else if (t[0] == "N") r.objective = true; else return bad("unexpected row type: " + t[0]);
        if (t.size() < 2) return bad("expected a row name in this line: " + lines[at]);
        if (t.size() > 2) return bad("unexpected input in row line: " + t[2]);
        if (row_of.count(t[1])) return bad("row name repeated: " + t[1]);
        row_of[t[1]] = (int)rows.size();
        rows.push_back(r);
        row_names.push_back(t[1]);
    }
    if (at >= lines.size()) return bad("could not find COLUMNS line");
    if (trim(lines[at]) != "COLUMNS") return bad("expected 'COLUMNS', found '" + trim(lines[at]) + "'");
    ++at;
    std::vector<ColRec> cols;
    std::unordered_map<std::string, int> col_of;
    std::map<std::pair<int, int>, bool> seen;
    for (; at < lines.size() && !starts(lines[at], "RHS"); ++at) {
        auto t = toks(lines[at]);
        if (t.size() < 1) return bad("expected a column name in this line: " + lines[at]);
        if (t.size() < 2) return bad("expected a row name in this line: " + lines[at]);
        if (t.size() < 3) return bad("expected a coefficient in this line: " + lines[at]);
        double v;
        if (!num(t[2], &v)) return bad("could not parse the coefficient " + t[2] + "\nline: " + lines[at]);
        if (t.size() > 3) return bad("unexpected input '" + t[3] + "' in column line: " + lines[at]);
        int cj;
        auto itc = col_of.find(t[0]);
        if (itc == col_of.end()) { cj = (int)cols.size(); col_of[t[0]] = cj; cols.push_back(ColRec{t[0], 0., false, Limits{ELLP_LOWER, 0., 0.}}); }
        else cj = itc->second;
        auto itr = row_of.find(t[1]);
        if (itr == row_of.end()) return bad("could not find the row " + t[1]);
        RowRec& r = rows[itr->second];
        if (r.objective) cols[cj].cost = v;
        else {
            if (seen.count({itr->second, cj}))
                return bad("specified constraint coefficient for the column " + t[0] + " and row " + t[1] + " more than once");
            seen[{itr->second, cj}] = true;
            r.terms.emplace_back(cj, v);
        }
    }
    if (at >= lines.size()) return bad("could not find RHS line");
    if (trim(lines[at]) != "RHS") return bad("expected 'RHS', found '" + trim(lines[at]) + "'");
    ++at;
    for (; at < lines.size() && !starts(lines[at], "BOUNDS") && !starts(lines[at], "ENDATA"); ++at) {
        auto t = toks(lines[at]);
        size_t k = (t.size() == 3) ? 1 : 0;
        if (t.size() <= k) return bad("expected a row name in this line: " + lines[at]);
        if (t.size() <= k + 1) return bad("expected a rhs value in this line: " + lines[at]);
        double v;
        if (!num(t[k + 1], &v)) return bad("could not parse the rhs value " + t[k + 1] + "\nline: " + lines[at]);
        auto itr = row_of.find(t[k]);
        if (itr == row_of.end()) return bad("could not find the row " + t[k]);
        RowRec& r = rows[itr->second];
        if (r.objective) return bad("should not specify rhs value for the objective");
        if (r.has_rhs) return bad("specified rhs for " + t[k] + " more than once");
        r.has_rhs = true;
        r.rhs = v;
        if (t.size() > k + 2) return bad("unexpected input in column line: " + t[k + 2]);
    }
    if (at < lines.size() && trim(lines[at]) != "ENDATA") {
        if (trim(lines[at]) != "BOUNDS") return bad("expected 'BOUNDS', found '" + trim(lines[at]) + "'");
        ++at;
        for (; at < lines.size() && !starts(lines[at], "ENDATA"); ++at) {
            auto t = toks(lines[at]);
            if (t.size() < 1) return bad("expected a bound type in this line: " + lines[at]);
            if (t.size() < 3) return bad("expected a column name in this line: " + lines[at]);
            double v = 0.;
            const bool has_v = t.size() >= 4;
            if (has_v && !num(t[3], &v)) return bad("could not parse the bound value\nline: " + lines[at]);
            Limits L;
            if (t[0] == "UP" && has_v) L = Limits{ELLP_UPPER, 0., v};
            else if (t[0] == "LO" && has_v) L = Limits{ELLP_LOWER, v, 0.};
            else if (t[0] == "FR" && !has_v) L = Limits{ELLP_FREE, 0., 0.};
            else return bad("invalid bound specification: " + lines[at]);
            auto itc = col_of.find(t[2]);
            if (itc == col_of.end()) return bad("found bound for the column " + t[2] + ", but it does not exist");
            ColRec& c = cols[itc->second];
            if (!c.has_lim) { c.has_lim = true; c.lim = L; }
            else if (c.lim.kind == ELLP_UPPER && L.kind == ELLP_LOWER) c.lim = Limits{ELLP_TWOSIDED, L.lo, c.lim.hi};
            else if (c.lim.kind == ELLP_LOWER && L.kind == ELLP_UPPER) c.lim = Limits{ELLP_TWOSIDED, c.lim.lo, L.hi};
            else return bad("invalid bounds for " + t[2]);
            if (t.size() > 4) return bad("unexpected input in column line: " + t[4]);
        }
    }
    if (at >= lines.size()) return bad("could not find ENDATA line");
    if (trim(lines[at]) != "ENDATA") return bad("expected 'ENDATA', found '" + trim(lines[at]) + "'");
    ++at;
    if (at < lines.size()) return bad("unexpected line: " + lines[at]);

    auto mdl = std::make_unique<ellp_b200_model>();
    Model& M = mdl->m;
    for (size_t j = 0; j < cols.size(); ++j) {
        Limits L = cols[j].has_lim ? cols[j].lim : Limits{ELLP_LOWER, 0., 0.};  // :31
        if (L.kind == ELLP_TWOSIDED && L.lo > L.hi) return bad("invalid variable bounds");
        M.vars.push_back(ModelVar{(int64_t)j, cols[j].cost, L});
    }
    for (auto& r : rows) {
        if (r.objective) continue;
        ModelRow mr;
        mr.op = r.op;
        mr.rhs = r.has_rhs ? r.rhs : 0.;  // :44
        for (auto& t : r.terms) mr.terms.emplace_back((int64_t)t.first, t.second);
        M.rows.push_back(std::move(mr));
    }
    M.f_ptr.push_back(0);
    for (auto& v : M.vars) { M.f_obj.push_back(v.cost); M.f_kind.push_back(v.lim.kind); M.f_lb.push_back(v.lim.lo); M.f_ub.push_back(v.lim.hi); M.f_id.push_back(v.id); }
    for (auto& r : M.rows) {
        for (auto& t : r.terms) { M.f_col.push_back(t.first); M.f_coef.push_back(t.second); }
        M.f_ptr.push_back((int32_t)M.f_col.size());
        M.f_op.push_back(r.op);
        M.f_rhs.push_back(r.rhs);
    }
    *out = mdl.release();
    return ELLP_OK;
}

void ellp_b200_model_free(ellp_b200_model* m) { delete m; }

void ellp_b200_model_desc(const ellp_b200_model* mh, ellp_problem_desc* out) {
    const Model& M = mh->m;
    out->nvars = (int32_t)M.vars.size();
    out->ncons = (int32_t)M.rows.size();
    out->obj = M.f_obj.data(); out->kind = M.f_kind.data(); out->lb = M.f_lb.data(); out->ub = M.f_ub.data();
    out->var_id = M.f_id.data(); out->row_ptr = M.f_ptr.data(); out->col_id = M.f_col.data(); out->coef = M.f_coef.data();
    out->op = M.f_op.data(); out->rhs = M.f_rhs.data();
}

}  // extern "C"
