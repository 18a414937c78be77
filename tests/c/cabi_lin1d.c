/*
 * Plain-C client of libcorleone_b200.so: no Python, no torch -- the boundary as a Julia `ccall` (or any FFI) sees it.
 *
 * Problem: lib/CorleoneOED/test/core/1d_oed.jl:11-13  x' = p x, x(0) = 1, p tunable (default -2), tspan (0, 1),
 * single shooting without controls, Tsit5 with 64 fixed steps.  Closed form: x(1) = exp(p), dx(1)/dp = exp(p).
 * The descriptor is filled by hand exactly as LuxCore.initialstates would produce it (Int64, 1-based).
 *
 *   gcc -O2 -I include tests/c/cabi_lin1d.c -o /tmp/cabi_lin1d -L corleone.jl_b200 -lcorleone_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "corleone_b200.h"

int main(void)
{
    cb200_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = (int32_t)sizeof d;
    d.model = CB200_MODEL_LIN1D;
    d.method = CB200_TSIT5;
    d.steps_per_piece = 64;
    d.nx = 1;
    d.np_all = 1;
    d.n_controls = 0;
    d.n_tunable_p = 1;
    d.n_intervals = 1;
    d.n_quadrature = 0;
    d.device = 0;
    const double u0[1] = {1.0};
    const int64_t tunable_p[1] = {1}, none64[1] = {0};
    const int32_t n_pieces[1] = {1}, n_u0[1] = {0}, n_c[1] = {0}, is_shooting[1] = {0}, ncon[1] = {1}, nsi[1] = {0};
    const double tspans[2] = {0.0, 1.0};
    const int64_t sorting[1] = {1}, constants[1] = {1};
    d.u0 = u0;
    d.tunable_p = tunable_p;
    d.control_indices = none64;
    d.quadrature_indices = none64;
    d.n_pieces = n_pieces;
    d.tspans = tspans;
    d.index_grid = none64;
    d.n_u0 = n_u0;
    d.n_local_controls = n_c;
    d.is_shooting = is_shooting;
    d.ic_sorting = sorting;
    d.ic_n_constants = ncon;
    d.ic_constants = constants;
    d.n_shooting_indices = nsi;
    d.shooting_indices = none64;

    cb200_handle *h = NULL;
    int rc = cb200_create(&d, &h);
    if (rc != CB200_OK) {
        fprintf(stderr, "cb200_create failed (%d): %s\n", rc, cb200_last_error());
        return 2;
    }
    cb200_sizes sz;
    cb200_get_sizes(h, &sz);
    if (sz.nvars != 1 || sz.n_defects != 0 || sz.n_sens != 1) {
        fprintf(stderr, "unexpected sizes: nvars %lld, defects %lld, sens %lld\n", (long long)sz.nvars,
                (long long)sz.n_defects, (long long)sz.n_sens);
        return 3;
    }
    enum { B = 5 };
    const double ps[B] = {-2.0, -1.0, 0.0, 0.5, 1.0};   /* nvars x B matrix, one candidate per column */
    double xend[B], sens[B], obj[B], grad[B];
    int32_t status[B];
    cb200_out out;
    memset(&out, 0, sizeof out);
    out.xend = xend;
    out.sens = sens;
    out.objective = obj;
    out.grad = grad;
    out.status = status;
    out.objective_row = 1;
    rc = cb200_eval(h, ps, B, 1, CB200_WANT_XEND | CB200_WANT_SENS | CB200_WANT_OBJ | CB200_WANT_GRAD, 0, &out);
    if (rc != CB200_OK) {
        fprintf(stderr, "cb200_eval failed (%d): %s\n", rc, cb200_last_error());
        return 4;
    }
    double worst = 0.0;
    for (int b = 0; b < B; b++) {
        const double ex = exp(ps[b]);
        const double e1 = fabs(xend[b] - ex) / ex, e2 = fabs(sens[b] - ex) / ex, e3 = fabs(obj[b] - xend[b]),
                     e4 = fabs(grad[b] - sens[b]);
        if (e1 > worst) worst = e1;
        if (e2 > worst) worst = e2;
        if (e3 > worst) worst = e3;
        if (e4 > worst) worst = e4;
        if (status[b] != 0) {
            fprintf(stderr, "candidate %d flagged\n", b);
            return 5;
        }
    }
    /* an out-of-range objective row must be refused with a message, not crash */
    out.objective_row = 7;
    if (cb200_eval(h, ps, B, 1, CB200_WANT_OBJ, 0, &out) != CB200_ERR_ARG || strlen(cb200_last_error()) == 0) return 6;
    cb200_destroy(h);
    printf("cabi_lin1d ok: version %d, max relative error vs exp(p) %.3e (Tsit5, 64 steps)\n", cb200_version(), worst);
    return worst < 1e-9 ? 0 : 1;
}
