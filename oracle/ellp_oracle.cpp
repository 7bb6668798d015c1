// ellp_oracle.cpp -- CPU ORACLE for the simplex hot path of kehlert/ellp.
//
// TEST INFRASTRUCTURE ONLY (see ellp_oracle.h).  Never linked into the product.
//
// Every function below restates the cited reference lines (paths relative to
// /root/reference).  The dense linear algebra of the reference lives in the
// un-vendored crate `nalgebra = "^0"` (Cargo.toml:16, no Cargo.lock); its
// published algorithms are restated in the "nalgebra" section from knowledge
// of that library (right-looking partial-pivot LU with reciprocal multipliers,
// Householder QR pivoting on the max-|entry| of the trailing block, full-pivot
// LU, column-oriented triangular solves).  Bit-exactness with the Rust binary
// is NOT claimed; every decision in the hot loop is EPS=1e-10 tolerant.
//
// Pinned by: tests/test_oracle_golden.py (all golden values of
// tests/problems/mod.rs:130-674).  Pivot sequences: parity unpinned.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off).

#include "ellp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

constexpr double EPS = 0.0000000001;  // util.rs:1
const double INF = std::numeric_limits<double>::infinity();

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };     // panic!/assert!
struct EllPError : std::runtime_error { using std::runtime_error::runtime_error; }; // Err(EllPError)

#define ORC_ASSERT(cond, msg) do { if (!(cond)) throw Panic(msg); } while (0)

// ---------------------------------------------------------------- dense types
using Vec = std::vector<double>;

struct Mat {  // nalgebra::DMatrix<f64>: column-major, contiguous
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)j * r + i]; }
    double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
    double* col(int j) { return a.data() + (size_t)j * r; }
    const double* col(int j) const { return a.data() + (size_t)j * r; }
    bool empty() const { return r == 0 || c == 0; }
};

// PermutationSequence: list of row swaps applied in order / in reverse.
using Perm = std::vector<std::pair<int, int>>;
template <class T> void permute_rows(const Perm& p, std::vector<T>& v) {
    for (auto& s : p) std::swap(v[s.first], v[s.second]);
}
template <class T> void inv_permute_rows(const Perm& p, std::vector<T>& v) {
    for (auto it = p.rbegin(); it != p.rend(); ++it) std::swap(v[it->first], v[it->second]);
}

double rust_signum(double x) {  // f64::signum: +0 -> 1, -0 -> -1, NaN -> NaN
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -1.0 : 1.0;
}

double dot(const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// ------------------------------------------------------------------ nalgebra
// linalg::LU::new -- partial pivoting, first max |a| in the column (icamax is
// a strict `>` scan), multipliers scaled by the reciprocal of the pivot,
// trailing update column by column (axpy, no fused multiply-add).
struct LU {
    Mat lu;
    Perm p;
    explicit LU(Mat m) : lu(std::move(m)) {
        const int nr = lu.r, nc = lu.c, mn = std::min(nr, nc);
        for (int i = 0; i < mn; ++i) {
            int piv = i;
            double best = std::fabs(lu(i, i));
            for (int k = i + 1; k < nr; ++k) {
                double v = std::fabs(lu(k, i));
                if (v > best) { best = v; piv = k; }
            }
            double diag = lu(piv, i);
            if (diag == 0.0) continue;  // no non-zero entries on this column
            if (piv != i) {
                p.emplace_back(i, piv);
                for (int j = 0; j < nc; ++j) std::swap(lu(i, j), lu(piv, j));
            }
            const double inv = 1.0 / diag;
            double* ci = lu.col(i);
            for (int k = i + 1; k < nr; ++k) ci[k] *= inv;
            for (int j = i + 1; j < nc; ++j) {
                double* cj = lu.col(j);
                const double a = -cj[i];
                for (int k = i + 1; k < nr; ++k) cj[k] = a * ci[k] + cj[k];
            }
        }
    }
    int dim() const { return std::min(lu.r, lu.c); }
    double u_diag(int i) const { return lu(i, i); }
    // LU::solve: permute, unit-lower forward (column oriented), upper backward
    // (column oriented); false when an upper diagonal is exactly zero.
    bool solve(Vec& b) const {
        const int n = lu.r;
        permute_rows(p, b);
        for (int i = 0; i < n; ++i) {
            const double coeff = b[i];
            const double* ci = lu.col(i);
            for (int k = i + 1; k < n; ++k) b[k] = -coeff * ci[k] + b[k];
        }
        for (int i = n - 1; i >= 0; --i) {
            const double diag = lu(i, i);
            if (diag == 0.0) return false;
            const double coeff = b[i] / diag;
            b[i] = coeff;
            const double* ci = lu.col(i);
            for (int k = 0; k < i; ++k) b[k] = -coeff * ci[k] + b[k];
        }
        return true;
    }
    // u().tr_solve_upper_triangular(b) then l().tr_solve_lower_triangular(..)
    // then p().inv_permute_rows(..)  == A^-T b   (primal_simplex_solver.rs:184-187)
    bool solve_transpose(Vec& b) const {
        const int n = lu.r;
        for (int i = 0; i < n; ++i) {  // U^T x = b : forward, dot form
            const double diag = lu(i, i);
            if (diag == 0.0) return false;
            b[i] = (b[i] - dot(lu.col(i), b.data(), i)) / diag;
        }
        for (int i = n - 1; i >= 0; --i) {  // L^T x = b : backward, unit diagonal
            b[i] = (b[i] - dot(lu.col(i) + i + 1, b.data() + i + 1, n - i - 1)) / 1.0;
        }
        inv_permute_rows(p, b);
        return true;
    }
};

// (row, col) of the first max |entry| in column-major scan order of the
// trailing block [i.., i..]  (Matrix::icamax_full, strict `>`).
std::pair<int, int> icamax_full(const Mat& m, int i) {
    double best = 0.0;
    std::pair<int, int> at(i, i);
    bool first = true;
    for (int j = i; j < m.c; ++j)
        for (int k = i; k < m.r; ++k) {
            double v = std::fabs(m(k, j));
            if (first || v > best) { best = v; at = {k, j}; first = false; }
        }
    return at;
}

// linalg::ColPivQR::new : Householder QR, column pivot = column holding the
// max |entry| of the trailing block.  Only |diag| and the permutation are used
// by the reference (standard_form.rs:142-181).
struct ColPivQR {
    Mat qr;
    Perm p;
    Vec diag;
    explicit ColPivQR(Mat m) : qr(std::move(m)) {
        const int nr = qr.r, nc = qr.c, mn = std::min(nr, nc);
        diag.assign(mn, 0.0);
        for (int i = 0; i < mn; ++i) {
            auto piv = icamax_full(qr, i);
            const int col_piv = piv.second;
            if (col_piv != i) {
                for (int k = 0; k < nr; ++k) std::swap(qr(k, i), qr(k, col_piv));
                p.emplace_back(i, col_piv);
            }
            // householder::clear_column_unchecked(matrix, i, 0, None)
            double* axis = qr.col(i) + i;
            const int len = nr - i;
            double sq = 0.0;
            for (int k = 0; k < len; ++k) sq += axis[k] * axis[k];
            const double norm = std::sqrt(sq);
            const double modulus = std::fabs(axis[0]);
            const double sign0 = rust_signum(axis[0]);
            const double signed_norm = sign0 * norm;
            const double factor = (sq + modulus * norm) * 2.0;
            axis[0] += signed_norm;
            if (factor != 0.0) {
                const double sf = std::sqrt(factor);
                for (int k = 0; k < len; ++k) axis[k] /= sf;
                double n2 = 0.0;  // normalize again
                for (int k = 0; k < len; ++k) n2 += axis[k] * axis[k];
                n2 = std::sqrt(n2);
                if (n2 != 0.0) for (int k = 0; k < len; ++k) axis[k] /= n2;
                const double refl_norm = -signed_norm;
                const double sign = rust_signum(refl_norm);
                for (int j = i + 1; j < nc; ++j) {  // reflect_with_sign
                    double* cj = qr.col(j) + i;
                    const double f = dot(axis, cj, len) * (-2.0 * sign);
                    for (int k = 0; k < len; ++k) cj[k] = f * axis[k] + sign * cj[k];
                }
                diag[i] = refl_norm;
            } else {
                diag[i] = signed_norm;
            }
        }
    }
    // r(): rows 0..min(nr,nc), upper triangle, diagonal replaced by |diag|
    Mat r() const {
        const int mn = std::min(qr.r, qr.c);
        Mat R(mn, qr.c);
        for (int j = 0; j < qr.c; ++j)
            for (int i = 0; i < mn && i <= j; ++i) R(i, j) = (i == j) ? std::fabs(diag[i]) : qr(i, j);
        return R;
    }
};

// linalg::FullPivLU::new (primal_problem.rs:162)
struct FullPivLU {
    Mat lu;
    Perm p, q;
    explicit FullPivLU(Mat m) : lu(std::move(m)) {
        const int nr = lu.r, nc = lu.c, mn = std::min(nr, nc);
        for (int i = 0; i < mn; ++i) {
            auto piv = icamax_full(lu, i);
            const int rp = piv.first, cp = piv.second;
            const double diag = lu(rp, cp);
            if (diag == 0.0) break;  // the remaining of the matrix is zero
            if (cp != i) {
                for (int k = 0; k < nr; ++k) std::swap(lu(k, i), lu(k, cp));
                q.emplace_back(i, cp);
            }
            if (rp != i) {
                p.emplace_back(i, rp);
                for (int j = 0; j < nc; ++j) std::swap(lu(i, j), lu(rp, j));
            }
            const double inv = 1.0 / diag;
            double* ci = lu.col(i);
            for (int k = i + 1; k < nr; ++k) ci[k] *= inv;
            for (int j = i + 1; j < nc; ++j) {
                double* cj = lu.col(j);
                const double a = -cj[i];
                for (int k = i + 1; k < nr; ++k) cj[k] = a * ci[k] + cj[k];
            }
        }
    }
};

// ------------------------------------------------------------- problem types
struct Bound {  // problem.rs:190-197
    uint8_t kind = ORC_LOWER;
    double lb = 0.0, ub = 0.0;
    static Bound Free() { return {ORC_FREE, 0, 0}; }
    static Bound Lower(double l) { return {ORC_LOWER, l, 0}; }
    static Bound Upper(double u) { return {ORC_UPPER, 0, u}; }
    static Bound TwoSided(double l, double u) { return {ORC_TWOSIDED, l, u}; }
    static Bound Fixed(double v) { return {ORC_FIXED, v, v}; }
};
struct Variable { int64_t id; double obj_coeff; Bound bound; };
struct Constraint { std::vector<std::pair<int64_t, double>> coeffs; uint8_t op; double rhs; };
struct Problem {  // problem.rs:11-17
    std::vector<Variable> variables;
    std::vector<Constraint> constraints;
};

struct Nonbasic { int index; uint8_t bound; };  // standard_form.rs:193-210
struct Point {                                  // standard_form.rs:20-25
    Vec x;
    std::vector<Nonbasic> N;
    std::vector<int> B;
};
struct StandardForm {  // standard_form.rs:27-34
    Vec c;
    Mat A;
    Vec b;
    std::vector<Bound> bounds;
    Problem prob;
    int rows() const { return A.r; }
    int cols() const { return A.c; }
    double obj(const Vec& x) const {  // :47-50 (c.dot(x); lengths are equal at every call site)
        return dot(c.data(), x.data(), (int)std::min(c.size(), x.size()));
    }
    double dual_obj(const Vec& y, const Vec& d) const {  // :52-68
        ORC_ASSERT(d.size() == bounds.size(), "assertion failed: d.len() == self.bounds.len()");
        double o = dot(b.data(), y.data(), (int)b.size());
        for (size_t i = 0; i < bounds.size(); ++i) {
            const Bound& bd = bounds[i];
            switch (bd.kind) {
                case ORC_FREE: break;
                case ORC_LOWER: o += bd.lb * d[i]; break;
                case ORC_UPPER: o += bd.ub * d[i]; break;
                case ORC_TWOSIDED: o += (d[i] > 0.) ? bd.lb * d[i] : bd.ub * d[i]; break;
                case ORC_FIXED: o += bd.lb * d[i]; break;
            }
        }
        return o;
    }
};

Problem problem_from_c(const orc_problem* p) {
    Problem prob;
    prob.variables.resize(p->nvars);
    for (int i = 0; i < p->nvars; ++i) {
        Variable& v = prob.variables[i];
        v.id = p->var_id ? p->var_id[i] : i;
        v.obj_coeff = p->obj[i];
        v.bound.kind = p->kind[i];
        v.bound.lb = p->lb ? p->lb[i] : 0.0;
        v.bound.ub = p->ub ? p->ub[i] : 0.0;
        if (v.bound.kind == ORC_FIXED) v.bound.ub = v.bound.lb;
    }
    prob.constraints.resize(p->ncons);
    for (int i = 0; i < p->ncons; ++i) {
        Constraint& c = prob.constraints[i];
        c.op = p->op[i];
        c.rhs = p->rhs[i];
        for (int k = p->row_ptr[i]; k < p->row_ptr[i + 1]; ++k) c.coeffs.emplace_back(p->col_id[k], p->coef[k]);
    }
    return prob;
}

// ---------------------------------------------- standard_form.rs:78-191
std::optional<StandardForm> to_standard_form(const Problem& prob) {
    const int n = (int)prob.variables.size();
    const int m = (int)prob.constraints.size();
    int num_slack = 0;  // :85-92
    for (auto& c : prob.constraints) num_slack += (c.op == ORC_LTE || c.op == ORC_GTE) ? 1 : 0;
    const int total = n + num_slack;  // :94

    Vec c(total, 0.0);  // :101-103
    Mat A(m, total);
    Vec b(m, 0.0);
    std::vector<Bound> bounds(total, Bound::Lower(0.));  // :106
    std::unordered_map<int64_t, int> id_to_index;
    for (int i = 0; i < n; ++i) {  // :109-113
        c[i] = prob.variables[i].obj_coeff;
        bounds[i] = prob.variables[i].bound;
        id_to_index[prob.variables[i].id] = i;
    }
    int cur_slack_col = total > 0 ? total - 1 : 0;  // :115 saturating_sub
    for (int i = 0; i < m; ++i) {                   // :117-137
        const Constraint& con = prob.constraints[i];
        b[i] = con.rhs;
        if (con.coeffs.empty() && b[i] != 0.) return std::nullopt;  // :120-122
        for (auto& kv : con.coeffs) {
            auto it = id_to_index.find(kv.first);
            ORC_ASSERT(it != id_to_index.end(), "called `Option::unwrap()` on a `None` value (unknown variable id)");
            A(i, it->second) = kv.second;
        }
        if (con.op != ORC_EQ) {
            A(i, cur_slack_col) = (con.op == ORC_LTE) ? 1. : -1.;  // :129-135
            cur_slack_col -= 1;                                    // may wrap below zero like usize only when total==0
        }
    }
    // :139 assert!(cur_slack_col == n.saturating_sub(1)) -- holds by construction, except the
    // usize underflow corner (n == 0 with slacks) where Rust would have panicked on `-= 1`.
    if (n == 0 && num_slack > 0) throw Panic("attempt to subtract with overflow (standard_form.rs:135)");

    // :142 A.transpose().col_piv_qr()
    Mat At(total, m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < total; ++j) At(j, i) = A(i, j);
    ColPivQR qr(std::move(At));
    Mat R = qr.r();              // :143
    inv_permute_rows(qr.p, b);   // :144
    for (int i = 0; i < R.r; ++i)  // :148-154
        if (std::fabs(R(i, i)) < EPS) R(i, i) = 0.;
    const bool is_trivial = !R.empty() && std::fabs(R(0, 0)) < EPS && std::fabs(b[0]) < EPS;  // :159
    if (!is_trivial) {  // :161-163  R.tr_solve_upper_triangular(&b)?  (None at the first zero diagonal)
        for (int i = 0; i < R.r; ++i)
            if (R(i, i) == 0.0) return std::nullopt;
    }
    permute_rows(qr.p, b);  // :165
    int num_indep = R.r;    // :170-174
    for (int i = 0; i < R.r; ++i)
        if (std::fabs(R(i, i)) < EPS) { num_indep = i; break; }
    std::vector<int> rows(m);  // :176-178
    for (int i = 0; i < m; ++i) rows[i] = i;
    permute_rows(qr.p, rows);
    rows.resize(num_indep);

    StandardForm sf;  // :180-189
    sf.A = Mat(num_indep, total);
    sf.b.assign(num_indep, 0.0);
    for (int k = 0; k < num_indep; ++k) {
        for (int j = 0; j < total; ++j) sf.A(k, j) = A(rows[k], j);
        sf.b[k] = b[rows[k]];
    }
    sf.c = std::move(c);
    sf.bounds = std::move(bounds);
    sf.prob = prob;
    return sf;
}

// ------------------------------------- solve_trivial_problem.rs:5-96
int solve_trivial_problem(const StandardForm& sf, Vec& x, std::vector<Nonbasic>& N, bool minimize) {
    N.clear();
    ORC_ASSERT(sf.c.size() == sf.bounds.size(), "assertion failed: c.len() == bounds.len()");
    const size_t len = std::min(x.size(), std::min(sf.c.size(), sf.bounds.size()));  // zip
    for (size_t i = 0; i < len; ++i) {
        const double c_i = sf.c[i];
        const Bound& bd = sf.bounds[i];
        switch (bd.kind) {
            case ORC_FREE:
                N.push_back({(int)i, ORC_NB_FREE});
                if (c_i != 0.) return ORC_UNBOUNDED;
                x[i] = 0.;
                break;
            case ORC_LOWER:
                N.push_back({(int)i, ORC_NB_LOWER});
                if (c_i > 0.) { if (minimize) x[i] = bd.lb; else return ORC_UNBOUNDED; }
                else if (minimize) x[i] = bd.lb;
                else { if (c_i != 0.) return ORC_UNBOUNDED; x[i] = bd.lb; }
                break;
            case ORC_UPPER:
                N.push_back({(int)i, ORC_NB_UPPER});
                if (c_i > 0.) { if (minimize) return ORC_UNBOUNDED; x[i] = bd.ub; }
                else if (minimize) { if (c_i != 0.) return ORC_UNBOUNDED; x[i] = bd.ub; }
                else x[i] = bd.ub;
                break;
            case ORC_TWOSIDED:
                if ((c_i > 0.) == minimize) { N.push_back({(int)i, ORC_NB_LOWER}); x[i] = bd.lb; }
                else { N.push_back({(int)i, ORC_NB_UPPER}); x[i] = bd.ub; }
                break;
            case ORC_FIXED:
                N.push_back({(int)i, ORC_NB_LOWER});
                x[i] = bd.lb;
                break;
        }
    }
    return ORC_OPTIMAL;
}

// ------------------------------------------------------------------- tracing
struct Trace {
    orc_trace_rec* buf = nullptr;
    int64_t cap = 0, len = 0;
    uint64_t iters[4] = {0, 0, 0, 0};
    void add(int phase, int64_t iter, int entering, int leaving, double step, double obj) {
        if (buf && len < cap) buf[len] = {phase, (int32_t)iter, entering, leaving, step, obj};
        ++len;
        ++iters[phase];
    }
};

Mat select_columns(const Mat& A, const std::vector<int>& idx) {
    Mat out(A.r, (int)idx.size());
    for (size_t k = 0; k < idx.size(); ++k) std::memcpy(out.col((int)k), A.col(idx[k]), sizeof(double) * A.r);
    return out;
}

// ------------------------- primal_simplex_solver.rs:95-236 (+ pivot :238-435)
struct PivotOut { int kind; /*0 pivot,1 optimal,2 unbounded*/ int nonbasic; int basic; uint8_t side; double lambda; };

PivotOut primal_pivot(const LU& lu, const Vec& r, const StandardForm& sf, Vec& x, std::vector<Nonbasic>& N,
                      std::vector<int>& B, int mode) {
    // :253-287 Dantzig selection
    int best = -1;
    double best_key = 0.0;
    if (mode == ORC_MODE_EXACT) {
        // Iterator::max_by keeps the LAST maximum: acc survives only if compare(acc, x) == Greater.
        for (int j = 0; j < (int)N.size(); ++j) {
            const double r_i = r[j];
            if (std::fabs(r_i) < EPS) continue;  // :258
            double key;
            if (r_i > 0. && N[j].bound == ORC_NB_UPPER) key = r_i;        // :265
            else if (!(r_i > 0.) && N[j].bound == ORC_NB_LOWER) key = -r_i;  // :266
            else if (N[j].bound == ORC_NB_FREE) key = std::fabs(r_i);     // :267
            else continue;
            if (best < 0) { best = j; best_key = key; continue; }
            bool acc_greater;
            if (std::fabs(best_key - key) >= EPS) {  // :281-282
                if (std::isnan(best_key) || std::isnan(key)) throw Panic("NaN detected");
                acc_greater = best_key > key;
            } else {
                acc_greater = N[best].index > N[j].index;  // :284
            }
            if (!acc_greater) { best = j; best_key = key; }
        }
    } else {
        // canonical (order-free) form, SURVEY appendix A.1: k* = max key; winner = largest
        // variable index among { j : k_j > k* - EPS }.
        double kmax = -INF;
        std::vector<double> keys(N.size(), -INF);
        for (int j = 0; j < (int)N.size(); ++j) {
            const double r_i = r[j];
            if (std::fabs(r_i) < EPS) continue;
            double key;
            if (r_i > 0. && N[j].bound == ORC_NB_UPPER) key = r_i;
            else if (!(r_i > 0.) && N[j].bound == ORC_NB_LOWER) key = -r_i;
            else if (N[j].bound == ORC_NB_FREE) key = std::fabs(r_i);
            else continue;
            if (std::isnan(key)) throw Panic("NaN detected");
            keys[j] = key;
            kmax = std::max(kmax, key);
        }
        for (int j = 0; j < (int)N.size(); ++j) {
            if (keys[j] == -INF) continue;
            if (kmax - keys[j] < EPS && (best < 0 || N[j].index > N[best].index)) best = j;
        }
    }
    if (best < 0) return {1, -1, -1, 0, 0.};  // :289-292

    const int q = N[best].index;
    // :295 lu.solve(A[:, q])
    Vec d(sf.A.col(q), sf.A.col(q) + sf.A.r);
    if (!lu.solve(d)) throw Panic("called `Option::unwrap()` on a `None` value (lu.solve)");
    const bool at_lower = N[best].bound == ORC_NB_LOWER;  // :296
    if (at_lower) for (double& v : d) v = -v;             // :298-300

    double lambda;  // :305-311
    switch (sf.bounds[q].kind) {
        case ORC_TWOSIDED: lambda = sf.bounds[q].ub - sf.bounds[q].lb; break;
        case ORC_FIXED: lambda = 0.; break;
        default: lambda = INF;
    }
    ORC_ASSERT(d.size() == B.size(), "assertion failed: d.len() == B.len()");

    auto ratio = [&](int i, double d_i) -> double {  // :325-367
        const Bound& bd = sf.bounds[B[i]];
        const double x_i = x[B[i]];
        switch (bd.kind) {
            case ORC_FREE: return INF;
            case ORC_LOWER:
                if (d_i > 0.) return INF;
                return (x_i > bd.lb) ? (bd.lb - x_i) / d_i : 0.;
            case ORC_UPPER:
                if (d_i > 0.) return (x_i < bd.ub) ? (bd.ub - x_i) / d_i : 0.;
                return INF;
            case ORC_TWOSIDED:
                if (d_i > 0.) return (x_i < bd.ub) ? (bd.ub - x_i) / d_i : 0.;
                return (x_i < bd.lb) ? (bd.lb - x_i) / d_i : 0.;  // :359 (sic: `<`, quirk Q3)
            default: return 0.;  // Fixed :366
        }
    };

    int new_basic = -1;
    uint8_t new_side = ORC_NB_LOWER;
    if (mode == ORC_MODE_EXACT) {
        bool have_nbi = false;
        int nbi = 0;
        for (int i = 0; i < (int)B.size(); ++i) {  // :320-400
            const double d_i = d[i];
            if (std::fabs(d_i) < EPS) continue;
            const double lambda_i = ratio(i, d_i);
            if (lambda_i < lambda - EPS) {  // :379-386
                lambda = lambda_i;
                new_basic = i;
                new_side = (d_i > 0.) ? ORC_NB_UPPER : ORC_NB_LOWER;
            } else if (std::fabs(lambda_i - lambda) < EPS) {  // :387-399
                if (!have_nbi || B[i] < nbi) {
                    have_nbi = true;
                    nbi = B[i];
                    lambda = lambda_i;
                    new_basic = i;
                    new_side = (d_i > 0.) ? ORC_NB_UPPER : ORC_NB_LOWER;
                }
            }
        }
    } else {
        // canonical form, SURVEY appendix A.2: lambda* = min; leaving = smallest variable index
        // among { i : lambda_i < lambda* + EPS }; the entering variable's own range lambda0 wins
        // (bound flip) only when every basic ratio is >= lambda0 + EPS.
        double lmin = INF;
        std::vector<double> lam(B.size(), INF);
        std::vector<char> cand(B.size(), 0);
        for (int i = 0; i < (int)B.size(); ++i) {
            const double d_i = d[i];
            if (std::fabs(d_i) < EPS) continue;
            lam[i] = ratio(i, d_i);
            cand[i] = 1;
            lmin = std::min(lmin, lam[i]);
        }
        if (lmin < lambda + EPS && lmin < INF) {
            for (int i = 0; i < (int)B.size(); ++i) {
                if (!cand[i]) continue;
                if (lam[i] - lmin < EPS && (new_basic < 0 || B[i] < B[new_basic])) new_basic = i;
            }
            lambda = lam[new_basic];
            new_side = (d[new_basic] > 0.) ? ORC_NB_UPPER : ORC_NB_LOWER;
        }
    }

    ORC_ASSERT(lambda >= 0., "assertion failed: lambda >= 0.");  // :402
    if (std::isinf(lambda)) return {2, best, -1, 0, lambda};      // :404-406
    if (lambda > 0.) {                                            // :408-417
        for (size_t i = 0; i < B.size(); ++i) x[B[i]] += lambda * d[i];
        if (at_lower) x[q] += lambda; else x[q] -= lambda;
        return {0, best, new_basic, new_side, lambda};
    } else if (lambda == 0. || lambda > -1E-6) {  // :425-430
        return {0, best, new_basic, new_side, lambda};
    }
    return {2, best, -1, 0, lambda};  // :431-434
}

int primal_solve_with_initial(const StandardForm& sf, Point& pt, uint64_t max_iter, int mode, Trace& tr, int phase) {
    Vec& x = pt.x;
    auto& N = pt.N;
    auto& B = pt.B;
    if (sf.rows() == 0) {  // :118-122
        ORC_ASSERT(B.empty(), "assertion failed: B.is_empty()");
        return solve_trivial_problem(sf, x, N, true);
    }
    if ((int)B.size() != sf.rows()) {  // :124-130
        char buf[128];
        snprintf(buf, sizeof buf, "invalid B, has %zu elements but %d expected", B.size(), sf.rows());
        throw EllPError(buf);
    }
    if (sf.cols() < sf.rows()) throw Panic("called `Option::unwrap()` on a `None` value (cols - rows)");  // :132
    const int expected_N = sf.cols() - sf.rows();
    if ((int)N.size() != expected_N) {  // :134-140
        char buf[128];
        snprintf(buf, sizeof buf, "invalid N, has %zu elements but %d expected", N.size(), expected_N);
        throw EllPError(buf);
    }
    Mat A_B = select_columns(sf.A, B);  // :142-145
    Vec c_B(B.size());
    for (size_t i = 0; i < B.size(); ++i) c_B[i] = sf.c[B[i]];
    if (N.empty()) return ORC_OPTIMAL;  // :149-151
    std::vector<int> Nidx(N.size());
    for (size_t j = 0; j < N.size(); ++j) Nidx[j] = N[j].index;
    Mat A_N = select_columns(sf.A, Nidx);  // :153-155
    Vec c_N(N.size());
    for (size_t j = 0; j < N.size(); ++j) c_N[j] = sf.c[N[j].index];

    uint64_t iter = 1;  // :157
    Vec r(N.size());
    for (;;) {
        if (iter > max_iter) return ORC_MAXITER;  // :163-166
        const double obj_before = sf.obj(x);
        iter += 1;
        LU lu(A_B);  // :173 (clone + lu)
        for (int i = 0; i < lu.dim(); ++i)  // :175-179
            if (std::fabs(lu.u_diag(i)) < EPS) throw EllPError("invalid B, A_B is not invertible");
        Vec u = c_B;  // :184-187
        if (!lu.solve_transpose(u)) throw Panic("called `Option::unwrap()` on a `None` value (tr_solve)");
        for (size_t j = 0; j < N.size(); ++j) r[j] = c_N[j] - dot(A_N.col((int)j), u.data(), A_N.r);  // :189

        PivotOut pv = primal_pivot(lu, r, sf, x, N, B, mode);  // :191
        if (pv.kind == 1) return ORC_OPTIMAL;                   // :198-200
        if (pv.kind == 2) return ORC_UNBOUNDED;                 // :202
        Nonbasic& nb = N[pv.nonbasic];
        const int entering = nb.index;
        if (pv.basic >= 0) {  // :208-221
            const int leaving = B[pv.basic];
            tr.add(phase, (int64_t)iter - 2, entering, leaving, pv.lambda, obj_before);
            std::swap(B[pv.basic], nb.index);
            double* a = A_N.col(pv.nonbasic);
            double* bcol = A_B.col(pv.basic);
            for (int k = 0; k < A_N.r; ++k) std::swap(a[k], bcol[k]);
            std::swap(c_N[pv.nonbasic], c_B[pv.basic]);
            nb.bound = pv.side;
        } else {  // :223-231
            tr.add(phase, (int64_t)iter - 2, entering, -1, pv.lambda, obj_before);
            if (nb.bound == ORC_NB_LOWER) nb.bound = ORC_NB_UPPER;
            else if (nb.bound == ORC_NB_UPPER) nb.bound = ORC_NB_LOWER;
            else throw Panic("pivot should have been unbounded");
        }
    }
}

// ------------------------------------------ primal_problem.rs:80-291
struct PrimalPhase {
    StandardForm sf;
    Point point;
    std::vector<int> phase_1_vars;
};

std::optional<PrimalPhase> make_primal_phase1(const Problem& prob) {
    auto sfo = to_standard_form(prob);  // :82-85
    if (!sfo) return std::nullopt;
    PrimalPhase ph;
    StandardForm& sf = ph.sf;
    sf = std::move(*sfo);
    const int n = sf.cols(), m = sf.rows();
    auto& N = ph.point.N;
    auto& B = ph.point.B;
    Vec v(n, 0.0);  // :93
    for (int i = 0; i < n; ++i) {  // :95-135
        const Bound& bd = sf.bounds[i];
        switch (bd.kind) {
            case ORC_FREE: break;
            case ORC_LOWER: v[i] = bd.lb; N.push_back({i, ORC_NB_LOWER}); break;
            case ORC_UPPER: v[i] = bd.ub; N.push_back({i, ORC_NB_UPPER}); break;
            case ORC_TWOSIDED: v[i] = bd.lb; N.push_back({i, ORC_NB_LOWER}); break;
            case ORC_FIXED: v[i] = bd.lb; N.push_back({i, ORC_NB_LOWER}); break;
        }
    }
    for (double& ci : sf.c) ci = 0.;  // :137-139
    sf.c.resize(n + m, 1.);           // :141
    std::vector<int> free_vars;       // :143-156
    for (int i = 0; i < n; ++i)
        if (sf.bounds[i].kind == ORC_FREE) free_vars.push_back(i);

    auto b_minus_Av = [&](const Vec& vv) {
        Vec bt = sf.b;
        for (int j = 0; j < sf.A.c; ++j) {
            const double vj = vv[j];
            const double* cj = sf.A.col(j);
            for (int i = 0; i < m; ++i) bt[i] -= cj[i] * vj;
        }
        return bt;
    };

    if (!free_vars.empty() && !sf.A.empty()) {  // :158-233
        Mat A_F = select_columns(sf.A, free_vars);
        FullPivLU lu(A_F);  // :162
        const int U_rows = std::min(A_F.r, A_F.c), U_cols = A_F.c;
        const int max_rank = std::min(U_rows, U_cols);
        int rank = (int)free_vars.size();  // :167-175 (falls back to free_vars.len(): quirk Q16)
        for (int i = 0; i < max_rank; ++i)
            if (std::fabs(lu.lu(i, i)) < EPS) { rank = i; break; }
        permute_rows(lu.q, free_vars);  // :180
        ORC_ASSERT(rank <= m, "index out of bounds (free-variable crash rank > rows, primal_problem.rs:199)");
        for (int k = 0; k < rank; ++k) B.push_back(free_vars[k]);  // :182-184
        for (size_t k = rank; k < free_vars.size(); ++k) {         // :186-193
            sf.bounds[free_vars[k]] = Bound::Fixed(0.);
            N.push_back({free_vars[k], ORC_NB_LOWER});
        }
        Vec bt = b_minus_Av(v);  // :196-199
        permute_rows(lu.p, bt);
        bt.resize(rank);
        // :201-210  U^-1 (L^-1 b~) on the leading rank x rank blocks (L unit lower)
        for (int i = 0; i < rank; ++i) {  // solve_lower_triangular, column oriented
            const double coeff = bt[i] / 1.0;
            bt[i] = coeff;
            for (int k = i + 1; k < rank; ++k) bt[k] = -coeff * lu.lu(k, i) + bt[k];
        }
        for (int i = rank - 1; i >= 0; --i) {  // solve_upper_triangular
            const double diag = lu.lu(i, i);
            if (diag == 0.0) throw Panic("called `Option::unwrap()` on a `None` value (solve_upper_triangular)");
            const double coeff = bt[i] / diag;
            bt[i] = coeff;
            for (int k = 0; k < i; ++k) bt[k] = -coeff * lu.lu(k, i) + bt[k];
        }
        for (int k = 0; k < rank; ++k) v[free_vars[k]] = bt[k];  // :214-216
        std::vector<int> rows(m);                                // :218-220
        for (int i = 0; i < m; ++i) rows[i] = i;
        permute_rows(lu.p, rows);
        rows.erase(rows.begin(), rows.begin() + rank);
        Vec bt2 = b_minus_Av(v);  // :222
        v.resize(n + m, 0.);      // :223
        Mat A2(m, n + (int)rows.size());  // :225
        std::memcpy(A2.a.data(), sf.A.a.data(), sizeof(double) * sf.A.a.size());
        sf.A = std::move(A2);
        int cur_col = sf.A.c - 1;  // :226
        for (int i : rows) {       // :228-233
            v[cur_col] = std::fabs(bt2[i]);
            sf.A(i, cur_col) = rust_signum(bt2[i]);
            B.push_back(cur_col);
            cur_col -= 1;
        }
    } else {  // :234-246
        Vec bt = b_minus_Av(v);
        v.resize(n + m, 0.);
        Mat A2(m, n + m);
        std::memcpy(A2.a.data(), sf.A.a.data(), sizeof(double) * sf.A.a.size());
        sf.A = std::move(A2);
        for (int i = 0; i < m; ++i) {
            const int index = n + i;
            v[index] = std::fabs(bt[i]);
            sf.A(i, index) = rust_signum(bt[i]);
            B.push_back(index);
        }
    }
    for (int k = 0; k < m; ++k) {  // :248-253
        ph.phase_1_vars.push_back((int)sf.bounds.size());
        sf.bounds.push_back(Bound::Lower(0.));
    }
    ph.point.x = std::move(v);
    return ph;
}

void primal_phase1_to_phase2(PrimalPhase& ph) {  // :263-291
    StandardForm& sf = ph.sf;
    for (int i : ph.phase_1_vars) {
        sf.c[i] = 0.;
        sf.bounds[i] = Bound::Fixed(0.);
    }
    for (size_t i = 0; i < sf.prob.variables.size(); ++i) {
        sf.c[i] = sf.prob.variables[i].obj_coeff;
        sf.bounds[i] = sf.prob.variables[i].bound;
    }
    for (auto& nb : ph.point.N)
        if (sf.bounds[nb.index].kind == ORC_FREE) nb.bound = ORC_NB_FREE;
}

struct SolveOut {
    int status = ORC_OPTIMAL;
    double obj = 0.;
    Vec x;  // Solution::x()
    bool fallback = false;
};

// primal_simplex_solver.rs:32-93
SolveOut primal_solve(const Problem& prob, uint64_t max_iter, int mode, Trace& tr) {
    SolveOut out;
    auto p1 = make_primal_phase1(prob);
    if (!p1) { out.status = ORC_INFEASIBLE; return out; }  // :33-36
    PrimalPhase& ph = *p1;
    int st = primal_solve_with_initial(ph.sf, ph.point, max_iter, mode, tr, 0);  // :40
    if (st == ORC_OPTIMAL) {
        const double obj = ph.sf.obj(ph.point.x);  // :42
        ORC_ASSERT(obj > -EPS, "assertion failed: obj > -EPS");
        if (!(obj < EPS)) { out.status = ORC_INFEASIBLE; return out; }  // :45-51
        primal_phase1_to_phase2(ph);
    } else if (st == ORC_INFEASIBLE) {
        out.status = ORC_INFEASIBLE; return out;  // :54-57
    } else if (st == ORC_UNBOUNDED) {
        throw Panic("primal phase 1 should never be unbounded");  // :59
    } else {
        out.status = ORC_MAXITER; out.obj = INF; return out;  // :61-64
    }
    st = primal_solve_with_initial(ph.sf, ph.point, max_iter, mode, tr, 1);  // :69
    if (st == ORC_OPTIMAL) {  // :70-79, solver.rs:47-53
        out.status = ORC_OPTIMAL;
        out.obj = ph.sf.obj(ph.point.x);
        out.x.assign(ph.point.x.begin(), ph.point.x.begin() + ph.sf.prob.variables.size());
    } else if (st == ORC_INFEASIBLE) {
        throw Panic("primal phase 2 should never be infeasible");  // :81
    } else if (st == ORC_UNBOUNDED) {
        out.status = ORC_UNBOUNDED;
    } else {
        out.status = ORC_MAXITER; out.obj = ph.sf.obj(ph.point.x);  // :88-91
    }
    return out;
}

// ------------------------------ dual_simplex_solver.rs:110-335
struct DualPoint { Vec y, d; Point point; };

int dual_solve_with_initial(const StandardForm& sf, DualPoint& dp, uint64_t max_iter, Trace& tr, int phase) {
    Vec& y = dp.y;
    Vec& d = dp.d;
    Vec& x = dp.point.x;
    auto& N = dp.point.N;
    auto& B = dp.point.B;
    if (sf.rows() == 0) {  // :132-136
        ORC_ASSERT(B.empty(), "assertion failed: B.is_empty()");
        return solve_trivial_problem(sf, x, N, true);
    }
    for (auto& nb : N) {  // :139-151
        const double d_i = d[nb.index];
        bool infeasible;
        switch (nb.bound) {
            case ORC_NB_LOWER: infeasible = d_i < -EPS; break;
            case ORC_NB_UPPER: infeasible = d_i > EPS; break;
            default: infeasible = std::fabs(d_i) > EPS;
        }
        if (infeasible) throw Panic("initial point of dual phase 2 is dual infeasible");
    }
    if ((int)B.size() != sf.rows()) {  // :153-159
        char buf[128];
        snprintf(buf, sizeof buf, "invalid B, has %zu elements but %d expected", B.size(), sf.rows());
        throw EllPError(buf);
    }
    if (sf.cols() < sf.rows()) throw Panic("called `Option::unwrap()` on a `None` value (cols - rows)");
    const int expected_N = sf.cols() - sf.rows();
    if ((int)N.size() != expected_N) {  // :163-169
        char buf[128];
        snprintf(buf, sizeof buf, "invalid N, has %zu elements but %d expected", N.size(), expected_N);
        throw EllPError(buf);
    }
    Mat A_B = select_columns(sf.A, B);  // :171-173
    Vec c_B(B.size());
    for (size_t i = 0; i < B.size(); ++i) c_B[i] = sf.c[B[i]];
    if (N.empty()) return ORC_OPTIMAL;  // :175-177
    std::vector<int> Nidx(N.size());
    for (size_t j = 0; j < N.size(); ++j) Nidx[j] = N[j].index;
    Mat A_N = select_columns(sf.A, Nidx);  // :179-181
    Vec c_N(N.size());
    for (size_t j = 0; j < N.size(); ++j) c_N[j] = sf.c[N[j].index];

    uint64_t iter = 0;                // :183
    double obj = sf.dual_obj(y, d);   // :184
    const int m = sf.rows();
    Vec alpha(N.size());
    for (;;) {
        if (iter >= max_iter) return ORC_MAXITER;  // :191-194
        iter += 1;
        // :200-236 first infeasible basic in position order
        int leaving = -1;
        double delta = 0.;
        uint8_t side = ORC_NB_LOWER;
        for (int i = 0; i < m && leaving < 0; ++i) {
            const double x_i = x[B[i]];
            const Bound& bd = sf.bounds[B[i]];
            switch (bd.kind) {
                case ORC_LOWER:
                    if (x_i < bd.lb - EPS) { leaving = i; delta = x_i - bd.lb; side = ORC_NB_LOWER; }
                    break;
                case ORC_UPPER:
                    if (x_i > bd.ub + EPS) { leaving = i; delta = x_i - bd.ub; side = ORC_NB_UPPER; }
                    break;
                case ORC_TWOSIDED:
                    if (x_i > bd.ub + EPS) { leaving = i; delta = x_i - bd.ub; side = ORC_NB_UPPER; }
                    else if (x_i < bd.lb - EPS) { leaving = i; delta = x_i - bd.lb; side = ORC_NB_LOWER; }
                    break;
                default: break;  // Free, Fixed never leave
            }
        }
        if (leaving < 0) return ORC_OPTIMAL;  // :243-246 (the reference computes the LU first: same result)
        LU lu(A_B);                           // :241
        Vec rho(m, 0.0);                      // :248-253
        rho[leaving] = 1.;
        if (!lu.solve_transpose(rho)) throw Panic("called `Option::unwrap()` on a `None` value (tr_solve)");
        for (size_t j = 0; j < N.size(); ++j) alpha[j] = dot(A_N.col((int)j), rho.data(), m);  // :255
        if (delta < 0.) for (double& a : alpha) a = -a;                                        // :257-259
        // :263-279 min_by keeps the FIRST minimum
        int entering = -1;
        double theta_dual = 0.;
        for (int j = 0; j < (int)N.size(); ++j) {
            bool keep;
            switch (N[j].bound) {
                case ORC_NB_LOWER: keep = alpha[j] > EPS; break;
                case ORC_NB_UPPER: keep = alpha[j] < -EPS; break;
                default: keep = true;
            }
            if (!keep) continue;
            const double ratio = d[N[j].index] / alpha[j];
            if (std::isnan(ratio) || (entering >= 0 && std::isnan(theta_dual)))
                throw Panic("called `Option::unwrap()` on a `None` value (partial_cmp)");
            if (entering < 0 || ratio < theta_dual) { entering = j; theta_dual = ratio; }
        }
        if (entering < 0) return ORC_INFEASIBLE;  // :281-284
        if (delta < 0.) {                         // :286-289
            for (double& a : alpha) a = -a;
            theta_dual = -theta_dual;
        }
        const int leave_var = B[leaving], enter_var = N[entering].index;
        Vec alpha_q(sf.A.col(enter_var), sf.A.col(enter_var) + m);  // :294
        if (!lu.solve(alpha_q)) throw Panic("called `Option::unwrap()` on a `None` value (lu.solve)");
        d[leave_var] = -theta_dual;                                                  // :296
        for (size_t j = 0; j < N.size(); ++j) d[N[j].index] -= theta_dual * alpha[j];  // :298-300
        d[enter_var] = 0.;                                                           // :302
        for (int i = 0; i < m; ++i) y[i] += theta_dual * rho[i];                     // :304
        const double theta_primal = delta / alpha_q[leaving];                        // :306
        for (int i = 0; i < m; ++i) x[B[i]] -= theta_primal * alpha_q[i];            // :310-312
        x[enter_var] += theta_primal;                                                // :314
        tr.add(phase, (int64_t)iter - 1, enter_var, leave_var, theta_primal, obj);
        obj += theta_dual * delta;                                                   // :316
        std::swap(B[leaving], N[entering].index);                                    // :322
        N[entering].bound = side;                                                    // :323
        double* a = A_B.col(leaving);
        double* bcol = A_N.col(entering);
        for (int k = 0; k < m; ++k) std::swap(a[k], bcol[k]);  // :325-331
        std::swap(c_B[leaving], c_N[entering]);                // :333
    }
}

// ------------------------------------------ dual_problem.rs:89-404
struct DualPhase1 {
    StandardForm sf;       // auxiliary boxed problem
    DualPoint point;
    StandardForm orig_sf;
};

std::optional<DualPhase1> make_dual_phase1(const Problem& prob) {
    auto osf = to_standard_form(prob);  // :91-94
    if (!osf) return std::nullopt;
    DualPhase1 ph;
    ph.orig_sf = std::move(*osf);
    const StandardForm& o = ph.orig_sf;
    Problem p1;  // :96-134
    std::vector<char> kept(o.cols(), 0);
    for (int i = 0; i < o.cols(); ++i) {
        Bound nb;
        switch (o.bounds[i].kind) {
            case ORC_FREE: nb = Bound::TwoSided(-1., 1.); break;
            case ORC_LOWER: nb = Bound::TwoSided(0., 1.); break;
            case ORC_UPPER: nb = Bound::TwoSided(-1., 0.); break;
            default: continue;
        }
        kept[i] = 1;
        p1.variables.push_back({(int64_t)i, o.c[i], nb});
    }
    for (int i = 0; i < o.rows(); ++i) {
        Constraint con;
        con.op = ORC_EQ;
        con.rhs = 0.;
        for (int j = 0; j < o.cols(); ++j)
            if (kept[j]) con.coeffs.emplace_back((int64_t)j, o.A(i, j));
        if (!con.coeffs.empty()) p1.constraints.push_back(std::move(con));
    }
    auto sfo = to_standard_form(p1);  // :136-139
    if (!sfo) return std::nullopt;
    ph.sf = std::move(*sfo);
    StandardForm& sf = ph.sf;

    Mat At(sf.A.c, sf.A.r);  // :141 std_form.A.transpose().lu()
    for (int i = 0; i < sf.A.r; ++i)
        for (int j = 0; j < sf.A.c; ++j) At(j, i) = sf.A(i, j);
    LU lut(std::move(At));
    for (int i = 0; i < lut.dim(); ++i)  // :143-147
        if (std::fabs(lut.u_diag(i)) < EPS) throw Panic("should always have a basis available");
    const int n = sf.A.c;
    std::vector<int> perm_cols(n);  // :150-152
    for (int i = 0; i < n; ++i) perm_cols[i] = i;
    permute_rows(lut.p, perm_cols);
    auto& B = ph.point.point.B;
    auto& N = ph.point.point.N;
    for (int i = 0; i < sf.A.r; ++i) B.push_back(perm_cols[i]);                    // :154-156
    for (int i = sf.A.r; i < n; ++i) N.push_back({perm_cols[i], ORC_NB_LOWER});    // :158-160
    Vec c_B(B.size());
    for (size_t i = 0; i < B.size(); ++i) c_B[i] = sf.c[B[i]];  // :162
    LU A_B_lu(select_columns(sf.A, B));                          // :163-164

    if (!B.empty()) {  // :166-226
        Vec y = c_B;
        if (!A_B_lu.solve_transpose(y)) throw Panic("called `Option::unwrap()` on a `None` value (tr_solve)");
        Vec d(sf.c);  // :172
        for (int j = 0; j < sf.A.c; ++j) d[j] = sf.c[j] - dot(sf.A.col(j), y.data(), sf.A.r);
        Vec x(sf.bounds.size(), 0.0);
        ORC_ASSERT(d.size() == sf.bounds.size(), "assertion failed: d.len() == std_form.bounds.len()");
        for (auto& nb : N) {  // :178-204
            const int i = nb.index;
            const Bound& bd = sf.bounds[i];
            if (bd.kind == ORC_TWOSIDED) {
                if (d[i] >= 0.) { x[i] = bd.lb; nb.bound = ORC_NB_LOWER; }
                else { x[i] = bd.ub; nb.bound = ORC_NB_UPPER; }
            } else if (bd.kind == ORC_FIXED) {
                x[i] = bd.lb;
                nb.bound = (d[i] >= 0.) ? ORC_NB_LOWER : ORC_NB_UPPER;
            } else {
                throw Panic("bounds should always be fixed or two-sided");
            }
        }
        Vec bt = sf.b;  // :207-208
        for (int j = 0; j < sf.A.c; ++j) {
            const double* cj = sf.A.col(j);
            for (int i = 0; i < sf.A.r; ++i) bt[i] -= cj[i] * x[j];
        }
        if (!A_B_lu.solve(bt)) throw Panic("called `Option::unwrap()` on a `None` value (lu.solve)");
        for (size_t i = 0; i < B.size(); ++i) x[B[i]] = bt[i];  // :212-214
        ph.point.y = std::move(y);
        ph.point.d = std::move(d);
        ph.point.point.x = std::move(x);
    } else {  // :227-254
        Vec x(N.size(), 0.0);
        ORC_ASSERT(N.size() == sf.bounds.size(), "assertion `left == right` failed (N.len(), bounds.len())");
        for (auto& nb : N) {
            nb.bound = ORC_NB_LOWER;
            const Bound& bd = sf.bounds[nb.index];
            if (bd.kind == ORC_TWOSIDED || bd.kind == ORC_FIXED) x[nb.index] = bd.lb;
            else throw Panic("bounds should always be fixed or two-sided");
        }
        ph.point.y.clear();
        ph.point.d = sf.c;
        ph.point.point.x = std::move(x);
    }
    return ph;
}

struct DualPhase2 { StandardForm sf; DualPoint point; };

DualPhase2 dual_phase1_to_phase2(DualPhase1& p1) {  // :258-404
    DualPhase2 p2;
    const Problem& p1prob = p1.sf.prob;
    p2.sf = std::move(p1.orig_sf);
    StandardForm& sf = p2.sf;
    std::vector<char> is_basic(sf.cols(), 0);
    auto& B = p2.point.point.B;
    auto& N = p2.point.point.N;
    for (int bidx : p1.point.point.B) {  // :264-273
        const int index = (int)p1prob.variables[bidx].id;
        is_basic[index] = 1;
        B.push_back(index);
    }
    if (!B.empty()) {  // :275-350
        Vec y(B.size());
        for (size_t i = 0; i < B.size(); ++i) y[i] = sf.c[B[i]];
        // NB: quirk Q17 -- when B.len() != rows the reference panics inside nalgebra (non-square triangular solve, or the
        // A.tr_mul(&y) dimension mismatch at :283); checked before the solve so that a non-square A_B never reaches it
        ORC_ASSERT((int)B.size() == sf.A.r, "Matrix multiplication dimensions mismatch (dual_problem.rs:283)");
        LU A_B_lu(select_columns(sf.A, B));
        if (!A_B_lu.solve_transpose(y)) throw Panic("called `Option::unwrap()` on a `None` value (tr_solve)");
        Vec d(sf.c);
        for (int j = 0; j < sf.A.c; ++j) d[j] = sf.c[j] - dot(sf.A.col(j), y.data(), sf.A.r);
        Vec x_N;
        for (int i = 0; i < sf.cols(); ++i) {  // :285-324
            if (is_basic[i]) continue;
            const double d_i = d[i];
            const Bound& bd = sf.bounds[i];
            switch (bd.kind) {
                case ORC_FREE:
                    ORC_ASSERT(std::fabs(d_i) < EPS, "assertion failed: d_i.abs() < EPS");
                    x_N.push_back(0.); N.push_back({i, ORC_NB_FREE}); break;
                case ORC_LOWER:
                    ORC_ASSERT(d_i > -EPS, "assertion failed: d_i > -EPS");
                    x_N.push_back(bd.lb); N.push_back({i, ORC_NB_LOWER}); break;
                case ORC_UPPER:
                    ORC_ASSERT(d_i < EPS, "assertion failed: d_i < EPS");
                    x_N.push_back(bd.ub); N.push_back({i, ORC_NB_UPPER}); break;
                case ORC_TWOSIDED:
                    if (d_i >= 0.) { x_N.push_back(bd.lb); N.push_back({i, ORC_NB_LOWER}); }
                    else { x_N.push_back(bd.ub); N.push_back({i, ORC_NB_UPPER}); }
                    break;
                default:
                    x_N.push_back(bd.lb); N.push_back({i, ORC_NB_LOWER});
            }
        }
        Vec rhs = sf.b;  // :326-328
        for (size_t k = 0; k < N.size(); ++k) {
            const double* cj = sf.A.col(N[k].index);
            for (int i = 0; i < sf.A.r; ++i) rhs[i] -= cj[i] * x_N[k];
        }
        if (!A_B_lu.solve(rhs)) throw Panic("called `Option::unwrap()` on a `None` value (lu.solve)");
        Vec x(sf.A.c, 0.0);  // :330-342
        ORC_ASSERT(B.size() == rhs.size(), "assertion failed: B.len() == x_B.len()");
        for (size_t i = 0; i < B.size(); ++i) x[B[i]] = rhs[i];
        for (size_t k = 0; k < N.size(); ++k) x[N[k].index] = x_N[k];
        p2.point.y = std::move(y);
        p2.point.d = std::move(d);
        p2.point.point.x = std::move(x);
    } else {  // :351-402
        Vec x_N(sf.A.c, 0.0);
        for (int i = 0; i < sf.cols(); ++i) {
            if (is_basic[i]) continue;
            const Bound& bd = sf.bounds[i];
            switch (bd.kind) {
                case ORC_FREE: x_N[i] = 0.; N.push_back({i, ORC_NB_FREE}); break;
                case ORC_LOWER: x_N[i] = bd.lb; N.push_back({i, ORC_NB_LOWER}); break;
                case ORC_UPPER: x_N[i] = bd.ub; N.push_back({i, ORC_NB_UPPER}); break;
                case ORC_TWOSIDED: x_N[i] = bd.lb; N.push_back({i, ORC_NB_LOWER}); break;
                default: x_N[i] = bd.lb; N.push_back({i, ORC_NB_LOWER});
            }
        }
        p2.point.y.clear();
        p2.point.d = sf.c;
        p2.point.point.x = std::move(x_N);
    }
    return p2;
}

// dual_simplex_solver.rs:33-108
SolveOut dual_solve(const Problem& prob, uint64_t max_iter, int mode, Trace& tr) {
    SolveOut out;
    auto p1o = make_dual_phase1(prob);
    if (!p1o) { out.status = ORC_INFEASIBLE; return out; }  // :34-37
    DualPhase1& p1 = *p1o;
    int st = dual_solve_with_initial(p1.sf, p1.point, max_iter, tr, 2);  // :41
    if (st == ORC_OPTIMAL) {
        const double obj = p1.sf.dual_obj(p1.point.y, p1.point.d);  // :43
        ORC_ASSERT(obj < EPS, "assertion failed: obj < EPS");       // :45
        if (!(obj > -EPS)) {  // :50-67 primal fallback with the DEFAULT primal solver (max_iter 1000)
            out = primal_solve(p1.orig_sf.prob, 1000, mode, tr);
            out.fallback = true;
            ORC_ASSERT(out.status != ORC_OPTIMAL, "assertion failed: matches!(result, Infeasible | Unbounded | MaxIter)");
            return out;
        }
    } else if (st == ORC_INFEASIBLE) {
        throw Panic("dual phase 1 should never be infeasible");  // :70-72
    } else if (st == ORC_UNBOUNDED) {
        throw Panic("dual phase 1 should never be unbounded");  // :74
    } else {
        out.status = ORC_MAXITER; out.obj = INF; return out;  // :76-79
    }
    DualPhase2 p2 = dual_phase1_to_phase2(p1);  // :49
    st = dual_solve_with_initial(p2.sf, p2.point, max_iter, tr, 3);  // :84
    if (st == ORC_OPTIMAL) {  // :85-94
        out.status = ORC_OPTIMAL;
        out.obj = p2.sf.obj(p2.point.point.x);
        const size_t nv = p2.sf.prob.variables.size();
        ORC_ASSERT(nv <= p2.point.point.x.size(), "Matrix slicing out of bounds");
        out.x.assign(p2.point.point.x.begin(), p2.point.point.x.begin() + nv);
    } else if (st == ORC_INFEASIBLE) {
        out.status = ORC_INFEASIBLE;  // :96-99
    } else if (st == ORC_UNBOUNDED) {
        // reachable only through solve_trivial_problem (rows()==0); the reference panics here (:101)
        throw Panic("dual phase 2 should never return unbounded");
    } else {
        out.status = ORC_MAXITER;
        out.obj = p2.sf.dual_obj(p2.point.y, p2.point.d);  // :103-106
    }
    return out;
}

void set_err(orc_result* out, const char* msg) {
    if (!out) return;
    std::snprintf(out->err, sizeof out->err, "%s", msg);
}

StandardForm sf_from_c(const orc_std_form* s) {
    StandardForm sf;
    sf.A = Mat(s->m, s->n);
    if (s->m && s->n) std::memcpy(sf.A.a.data(), s->A, sizeof(double) * (size_t)s->m * s->n);
    sf.c.assign(s->c, s->c + s->n);
    sf.b.assign(s->b, s->b + s->m);
    sf.bounds.resize(s->n);
    for (int i = 0; i < s->n; ++i) {
        sf.bounds[i].kind = s->kind[i];
        sf.bounds[i].lb = s->lb[i];
        sf.bounds[i].ub = s->ub[i];
        if (s->kind[i] == ORC_FIXED) sf.bounds[i].ub = s->lb[i];
    }
    return sf;
}

void point_from_c(const orc_std_form* s, const orc_point* p, Point& pt) {
    pt.x.assign(p->x, p->x + s->n);
    pt.B.assign(p->B, p->B + p->nB);
    pt.N.resize(p->nN);
    for (int j = 0; j < p->nN; ++j) pt.N[j] = {p->N[j], p->N_side[j]};
}

void point_to_c(const Point& pt, orc_point* p) {
    std::copy(pt.x.begin(), pt.x.end(), p->x);
    std::copy(pt.B.begin(), pt.B.end(), p->B);
    for (size_t j = 0; j < pt.N.size() && (int)j < p->nN; ++j) { p->N[j] = pt.N[j].index; p->N_side[j] = pt.N[j].bound; }
    p->nN = (int)pt.N.size();
}

template <class F> int guarded(orc_result* out, F&& f) {
    try {
        f();
        return ORC_OK;
    } catch (const EllPError& e) {
        set_err(out, e.what());
        return ORC_ERR_ELLP;
    } catch (const Panic& e) {
        set_err(out, e.what());
        return ORC_ERR_PANIC;
    } catch (const std::exception& e) {
        set_err(out, e.what());
        return ORC_ERR_PANIC;
    }
}

}  // namespace

struct orc_stage {
    StandardForm sf;
    Point pt;
    Vec y, d;
};

extern "C" {

const char* ellp_oracle_version(void) { return "ellp-oracle 1 (restates kehlert/ellp 0.2.0, Cargo.toml:3)"; }

int ellp_oracle_solve(const orc_problem* p, int solver, uint64_t max_iter, int mode, orc_result* out) {
    out->err[0] = 0;
    out->trace_len = 0;
    out->used_primal_fallback = 0;
    for (auto& v : out->iters) v = 0;
    Trace tr;
    tr.buf = out->trace;
    tr.cap = out->trace_cap;
    int rc = guarded(out, [&] {
        Problem prob = problem_from_c(p);
        SolveOut so = (solver == ORC_PRIMAL) ? primal_solve(prob, max_iter, mode, tr) : dual_solve(prob, max_iter, mode, tr);
        out->status = so.status;
        out->obj = so.obj;
        out->used_primal_fallback = so.fallback ? 1 : 0;
        if (out->x && so.status == ORC_OPTIMAL) std::copy(so.x.begin(), so.x.end(), out->x);
    });
    out->trace_len = tr.len;
    for (int k = 0; k < 4; ++k) out->iters[k] = tr.iters[k];
    return rc;
}

int ellp_oracle_primal_solve_with_initial(const orc_std_form* s, orc_point* p, uint64_t max_iter, int mode, orc_result* out) {
    out->err[0] = 0;
    for (auto& v : out->iters) v = 0;
    Trace tr;
    tr.buf = out->trace;
    tr.cap = out->trace_cap;
    int rc = guarded(out, [&] {
        StandardForm sf = sf_from_c(s);
        Point pt;
        point_from_c(s, p, pt);
        out->status = primal_solve_with_initial(sf, pt, max_iter, mode, tr, 1);
        out->obj = sf.obj(pt.x);
        point_to_c(pt, p);
    });
    out->trace_len = tr.len;
    for (int k = 0; k < 4; ++k) out->iters[k] = tr.iters[k];
    return rc;
}

int ellp_oracle_dual_solve_with_initial(const orc_std_form* s, orc_point* p, uint64_t max_iter, int mode, orc_result* out) {
    (void)mode;
    out->err[0] = 0;
    for (auto& v : out->iters) v = 0;
    Trace tr;
    tr.buf = out->trace;
    tr.cap = out->trace_cap;
    int rc = guarded(out, [&] {
        StandardForm sf = sf_from_c(s);
        DualPoint dp;
        point_from_c(s, p, dp.point);
        dp.y.assign(p->y, p->y + s->m);
        dp.d.assign(p->d, p->d + s->n);
        out->status = dual_solve_with_initial(sf, dp, max_iter, tr, 3);
        out->obj = sf.dual_obj(dp.y, dp.d);
        point_to_c(dp.point, p);
        std::copy(dp.y.begin(), dp.y.end(), p->y);
        std::copy(dp.d.begin(), dp.d.end(), p->d);
    });
    out->trace_len = tr.len;
    for (int k = 0; k < 4; ++k) out->iters[k] = tr.iters[k];
    return rc;
}

orc_stage* ellp_oracle_stage_new(const orc_problem* p, int which, int* infeasible, char* err256) {
    *infeasible = 0;
    if (err256) err256[0] = 0;
    try {
        Problem prob = problem_from_c(p);
        auto* st = new orc_stage();
        if (which == 0) {
            auto sf = to_standard_form(prob);
            if (!sf) { delete st; *infeasible = 1; return nullptr; }
            st->sf = std::move(*sf);
        } else if (which == 1) {
            auto ph = make_primal_phase1(prob);
            if (!ph) { delete st; *infeasible = 1; return nullptr; }
            st->sf = std::move(ph->sf);
            st->pt = std::move(ph->point);
        } else {
            auto ph = make_dual_phase1(prob);
            if (!ph) { delete st; *infeasible = 1; return nullptr; }
            st->sf = std::move(ph->sf);
            st->pt = std::move(ph->point.point);
            st->y = std::move(ph->point.y);
            st->d = std::move(ph->point.d);
        }
        return st;
    } catch (const std::exception& e) {
        if (err256) std::snprintf(err256, 256, "%s", e.what());
        return nullptr;
    }
}

void ellp_oracle_stage_free(orc_stage* s) { delete s; }

void ellp_oracle_stage_dims(const orc_stage* s, int32_t* m, int32_t* n, int32_t* nx, int32_t* nB, int32_t* nN) {
    *m = s->sf.rows();
    *n = s->sf.cols();
    *nx = (int32_t)s->pt.x.size();
    *nB = (int32_t)s->pt.B.size();
    *nN = (int32_t)s->pt.N.size();
}

void ellp_oracle_stage_copy(const orc_stage* s, double* A, double* c, int32_t* len_c, double* b, uint8_t* kind,
                            double* lb, double* ub, int32_t* len_bounds, double* x, int32_t* B, int32_t* N,
                            uint8_t* N_side, double* y, double* d) {
    if (A) std::copy(s->sf.A.a.begin(), s->sf.A.a.end(), A);
    if (c) std::copy(s->sf.c.begin(), s->sf.c.end(), c);
    if (len_c) *len_c = (int32_t)s->sf.c.size();
    if (b) std::copy(s->sf.b.begin(), s->sf.b.end(), b);
    if (len_bounds) *len_bounds = (int32_t)s->sf.bounds.size();
    for (size_t i = 0; i < s->sf.bounds.size(); ++i) {
        if (kind) kind[i] = s->sf.bounds[i].kind;
        if (lb) lb[i] = s->sf.bounds[i].lb;
        if (ub) ub[i] = s->sf.bounds[i].ub;
    }
    if (x) std::copy(s->pt.x.begin(), s->pt.x.end(), x);
    if (B) std::copy(s->pt.B.begin(), s->pt.B.end(), B);
    for (size_t j = 0; j < s->pt.N.size(); ++j) {
        if (N) N[j] = s->pt.N[j].index;
        if (N_side) N_side[j] = s->pt.N[j].bound;
    }
    if (y) std::copy(s->y.begin(), s->y.end(), y);
    if (d) std::copy(s->d.begin(), s->d.end(), d);
}

void ellp_oracle_rank1_update(double* E, int64_t R, int64_t C, int64_t ld, const double* alpha, const double* rho,
                              int64_t r) {
    const double ar = alpha[r];
    for (int64_t j = 0; j < C; ++j) {
        const double p = rho[j] / ar;
        double* col = E + j * ld;
        for (int64_t i = 0; i < R; ++i) col[i] = (i == r) ? p : std::fma(-alpha[i], p, col[i]);
    }
}

void ellp_oracle_gemv_t(const double* M, int64_t R, int64_t C, int64_t ld, const double* v, double* y) {
    // Mirrors the summation order of the CUDA warp-per-column kernel: lane l of 32 owns the
    // double2 chunks l, l+32, ... (i.e. rows 2l, 2l+1, 2l+64, ...) accumulated with fma into two
    // partial sums (even row, odd row), added together, then a 5-step xor butterfly (16,8,4,2,1).
    for (int64_t j = 0; j < C; ++j) {
        const double* col = M + j * ld;
        double lane[32];
        for (int l = 0; l < 32; ++l) {
            double s0 = 0.0, s1 = 0.0;
            for (int64_t i = 2 * l; i < R; i += 64) {
                s0 = std::fma(col[i], v[i], s0);
                if (i + 1 < R) s1 = std::fma(col[i + 1], v[i + 1], s1);
            }
            lane[l] = s0 + s1;
        }
        for (int off = 16; off >= 1; off >>= 1) {
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
            for (int l = 0; l < 32; ++l) lane[l] = nxt[l];
        }
        y[j] = lane[0];
    }
}

}  // extern "C"
