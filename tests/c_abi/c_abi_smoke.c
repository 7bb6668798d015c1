/* c_abi_smoke.c -- plain C99 consumer of include/ellp_b200.h (no C++, no CUDA headers, no torch).
 *
 * Proves that the header is valid C, that the struct layouts equal the ctypes mirrors of ellp_b200/_native.py
 * (`layout` mode prints sizeof / offsetof of every field as JSON; tests/test_abi_cpu.py compares), and that a C program
 * can drive the drop-in boundary: `solve` mode calls ellp_b200_primal_solve_with_initial and
 * ellp_b200_dual_solve_with_initial -- the replacements of PrimalSimplexSolver::solve_with_initial
 * (primal_simplex_solver.rs:95-236) and DualSimplexSolver::solve_with_initial (dual_simplex_solver.rs:110-335) -- on a
 * 3 x 5 LP and checks status, objective and point.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude tests/c_abi/c_abi_smoke.c -o c_abi_smoke -Lellp_b200 -lellp_b200
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "ellp_b200.h"

#define FIELD(T, f) printf("%s\"%s.%s\": [%zu, %zu]", first ? "" : ", ", #T, #f, offsetof(T, f), sizeof(((T*)0)->f)), first = 0
#define SIZE(T) printf("%s\"sizeof %s\": %zu", first ? "" : ", ", #T, sizeof(T)), first = 0

static int layout(void) {
    int first = 1;
    printf("{");
    SIZE(ellp_std_form); FIELD(ellp_std_form, m); FIELD(ellp_std_form, n); FIELD(ellp_std_form, A); FIELD(ellp_std_form, c);
    FIELD(ellp_std_form, b); FIELD(ellp_std_form, kind); FIELD(ellp_std_form, lb); FIELD(ellp_std_form, ub);
    SIZE(ellp_point); FIELD(ellp_point, x); FIELD(ellp_point, B); FIELD(ellp_point, N); FIELD(ellp_point, N_side);
    FIELD(ellp_point, y); FIELD(ellp_point, d); FIELD(ellp_point, nB); FIELD(ellp_point, nN);
    SIZE(ellp_trace_rec); FIELD(ellp_trace_rec, phase); FIELD(ellp_trace_rec, iter); FIELD(ellp_trace_rec, entering);
    FIELD(ellp_trace_rec, leaving); FIELD(ellp_trace_rec, step); FIELD(ellp_trace_rec, obj);
    SIZE(ellp_opts); FIELD(ellp_opts, max_iter); FIELD(ellp_opts, tie_rule); FIELD(ellp_opts, engine); FIELD(ellp_opts, refactor_every);
    FIELD(ellp_opts, check_every); FIELD(ellp_opts, phase_tag); FIELD(ellp_opts, profile); FIELD(ellp_opts, trace);
    FIELD(ellp_opts, trace_cap); FIELD(ellp_opts, pricing); FIELD(ellp_opts, ratio); FIELD(ellp_opts, block_k);
    SIZE(ellp_result); FIELD(ellp_result, status); FIELD(ellp_result, iters); FIELD(ellp_result, obj); FIELD(ellp_result, trace_len);
    FIELD(ellp_result, launches); FIELD(ellp_result, ms_device); FIELD(ellp_result, ms_rank1); FIELD(ellp_result, n_rank1);
    FIELD(ellp_result, refactors);
    SIZE(ellp_batch); FIELD(ellp_batch, nlp); FIELD(ellp_batch, m); FIELD(ellp_batch, n); FIELD(ellp_batch, A); FIELD(ellp_batch, c);
    FIELD(ellp_batch, b); FIELD(ellp_batch, kind); FIELD(ellp_batch, lb); FIELD(ellp_batch, ub);
    SIZE(ellp_batch_result); FIELD(ellp_batch_result, status); FIELD(ellp_batch_result, obj); FIELD(ellp_batch_result, x);
    FIELD(ellp_batch_result, iters); FIELD(ellp_batch_result, err); FIELD(ellp_batch_result, trace); FIELD(ellp_batch_result, trace_cap);
    FIELD(ellp_batch_result, trace_len); FIELD(ellp_batch_result, ms_device); FIELD(ellp_batch_result, launches);
    FIELD(ellp_batch_result, pivots);
    SIZE(ellp_problem_desc); FIELD(ellp_problem_desc, nvars); FIELD(ellp_problem_desc, ncons); FIELD(ellp_problem_desc, obj);
    FIELD(ellp_problem_desc, kind); FIELD(ellp_problem_desc, lb); FIELD(ellp_problem_desc, ub); FIELD(ellp_problem_desc, var_id);
    FIELD(ellp_problem_desc, row_ptr); FIELD(ellp_problem_desc, col_id); FIELD(ellp_problem_desc, coef); FIELD(ellp_problem_desc, op);
    FIELD(ellp_problem_desc, rhs);
    SIZE(ellp_solution); FIELD(ellp_solution, status); FIELD(ellp_solution, obj); FIELD(ellp_solution, x); FIELD(ellp_solution, iters);
    FIELD(ellp_solution, used_primal_fallback); FIELD(ellp_solution, trace_len); FIELD(ellp_solution, launches);
    FIELD(ellp_solution, ms_device);
    printf("}\n");
    return 0;
}

/* min -3 x0 - 5 x1   s.t.  x0 + s0 = 4,  2 x1 + s1 = 12,  3 x0 + 2 x1 + s2 = 18,  all >= 0   (optimum -36 at x = (2, 6)) */
static int solve(void) {
    const double A[15] = {1, 0, 3, 0, 2, 2, 1, 0, 0, 0, 1, 0, 0, 0, 1}; /* column-major 3 x 5 */
    const double c[5] = {-3, -5, 0, 0, 0}, b[3] = {4, 12, 18};
    const uint8_t kind[5] = {ELLP_LOWER, ELLP_LOWER, ELLP_LOWER, ELLP_LOWER, ELLP_LOWER};
    const double lb[5] = {0, 0, 0, 0, 0}, ub[5] = {0, 0, 0, 0, 0};
    ellp_b200_ctx* ctx = NULL;
    ellp_std_form sf;
    ellp_point pt;
    ellp_opts o;
    ellp_result res;
    double x[5] = {0, 0, 4, 12, 18};
    int32_t B[3] = {2, 3, 4}, N[2] = {0, 1};
    uint8_t Ns[2] = {ELLP_NB_LOWER, ELLP_NB_LOWER};
    int rc = ellp_b200_create(0, &ctx);
    if (rc != ELLP_OK) { fprintf(stderr, "ellp_b200_create failed (%d): no CUDA device?\n", rc); return 2; }
    printf("%s\n", ellp_b200_version());
    sf.m = 3; sf.n = 5; sf.A = A; sf.c = c; sf.b = b; sf.kind = kind; sf.lb = lb; sf.ub = ub;
    pt.x = x; pt.B = B; pt.N = N; pt.N_side = Ns; pt.y = NULL; pt.d = NULL; pt.nB = 3; pt.nN = 2;
    ellp_b200_default_opts(&o);
    if (o.max_iter != 1000) { fprintf(stderr, "default max_iter %llu\n", (unsigned long long)o.max_iter); return 1; }
    rc = ellp_b200_primal_solve_with_initial(ctx, &sf, &pt, &o, &res);
    if (rc != ELLP_OK) { fprintf(stderr, "primal rc=%d: %s\n", rc, ellp_b200_last_error(ctx)); return 1; }
    printf("primal: status %d obj %.12g x = (%g, %g) pivots %llu launches %llu\n", res.status, res.obj, x[0], x[1],
           (unsigned long long)res.iters, (unsigned long long)res.launches);
    if (res.status != ELLP_OPTIMAL || fabs(res.obj + 36.0) > 1e-9 || fabs(x[0] - 2.0) > 1e-9 || fabs(x[1] - 6.0) > 1e-9) return 1;
    /* Err(EllPError) text is preserved: wrong basis length (primal :124-130) */
    pt.nB = 2;
    rc = ellp_b200_primal_solve_with_initial(ctx, &sf, &pt, &o, &res);
    if (rc != ELLP_E_ELLP || strcmp(ellp_b200_last_error(ctx), "invalid B, has 2 elements but 3 expected") != 0) {
        fprintf(stderr, "expected the reference's Err, got rc=%d '%s'\n", rc, ellp_b200_last_error(ctx));
        return 1;
    }
    /* dual: min 3 x0 + 5 x1  s.t.  x0 - s0 = 4 ... as Gte rows: the slack basis (-I) is dual feasible, x_B = -b */
    {
        const double Ad[15] = {1, 0, 3, 0, 2, 2, -1, 0, 0, 0, -1, 0, 0, 0, -1};
        const double cd[5] = {3, 5, 0, 0, 0};
        double xd[5] = {0, 0, -4, -12, -18}, y[3] = {0, 0, 0}, d[5] = {3, 5, 0, 0, 0};
        int32_t Bd[3] = {2, 3, 4}, Nd[2] = {0, 1};
        uint8_t Nsd[2] = {ELLP_NB_LOWER, ELLP_NB_LOWER};
        int engines[2] = {ELLP_ENGINE_REVISED, ELLP_ENGINE_TABLEAU};
        int e;
        for (e = 0; e < 2; ++e) {
            double xe[5], ye[3], de[5];
            int32_t Be[3], Ne[2];
            uint8_t Nse[2];
            memcpy(xe, xd, sizeof xd); memcpy(ye, y, sizeof y); memcpy(de, d, sizeof d);
            memcpy(Be, Bd, sizeof Bd); memcpy(Ne, Nd, sizeof Nd); memcpy(Nse, Nsd, sizeof Nsd);
            sf.A = Ad; sf.c = cd;
            pt.x = xe; pt.B = Be; pt.N = Ne; pt.N_side = Nse; pt.y = ye; pt.d = de; pt.nB = 3; pt.nN = 2;
            o.engine = engines[e];
            o.block_k = 4;
            rc = ellp_b200_dual_solve_with_initial(ctx, &sf, &pt, &o, &res);
            if (rc != ELLP_OK) { fprintf(stderr, "dual rc=%d: %s\n", rc, ellp_b200_last_error(ctx)); return 1; }
            printf("dual (engine %d): status %d obj %.12g x = (%g, %g) y = (%g, %g, %g) pivots %llu\n", engines[e], res.status, res.obj,
                   xe[0], xe[1], ye[0], ye[1], ye[2], (unsigned long long)res.iters);
            /* optimum of min 3 x0 + 5 x1, x0 >= 4, 2 x1 >= 12, 3 x0 + 2 x1 >= 18: x = (4, 6), obj 42 */
            if (res.status != ELLP_OPTIMAL || fabs(res.obj - 42.0) > 1e-9 || fabs(xe[0] - 4.0) > 1e-9 || fabs(xe[1] - 6.0) > 1e-9) return 1;
        }
    }
    ellp_b200_destroy(ctx);
    printf("C_ABI_SMOKE_OK\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "solve") == 0) return solve();
    return layout();
}
