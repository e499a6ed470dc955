/* examples/robertson.c -- the reference's example program (test/example.f90:30-77) as a C program on
 * the batched solver: the same arguments, the same call loop, the same table -- for a whole batch.
 *
 *   gcc -O2 -I include examples/robertson.c -L dvode_b200 -lbvode -Wl,-rpath,$PWD/dvode_b200 -o robertson
 *   ./robertson [nsys] [strict]
 *
 * System 0 has the reference's rate constants (0.04, 1e4, 3e7); system i > 0 has them scaled by
 * (1 + 0.01 i).  Prints, per output point, t and y of system 0 and of the last system with %.17g (the
 * test compares them bit for bit with the CPU oracle), then the statistics of system 0 in the
 * reference's words. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bvode.h"

int main(int argc, char **argv) {
    const int nsys = argc > 1 ? atoi(argv[1]) : 4;
    const unsigned flags = (argc > 2 && strcmp(argv[2], "strict") == 0) ? BVODE_FLAG_STRICT : 0;
    const int neq = 3, lrw = 20, liw = 30; /* the headers only: the work arrays proper live on the GPU */
    double *y = malloc(sizeof(double) * nsys * neq), *t = malloc(sizeof(double) * nsys);
    double *k = malloc(sizeof(double) * nsys * 3), *rwork = calloc((size_t)nsys * lrw, sizeof(double));
    int *istate = malloc(sizeof(int) * nsys), *iwork = calloc((size_t)nsys * liw, sizeof(int));
    double atol[3] = {1.0e-8, 1.0e-14, 1.0e-6}, rtol[1] = {1.0e-4}, tout = 0.4;
    const int itol = 2, itask = 1, iopt = 0, mf = 21;
    bvode_solver *solver;
    int rc, i, iout;

    for (i = 0; i < nsys; i++) {
        const double sc = 1.0 + 0.01 * i;
        k[3 * i] = 0.04 * sc; k[3 * i + 1] = 1.0e4 * sc; k[3 * i + 2] = 3.0e7 * sc;
        y[3 * i] = 1.0; y[3 * i + 1] = 0.0; y[3 * i + 2] = 0.0;
        t[i] = 0.0;
        istate[i] = 1;
    }
    /* call solver%initialize(f=fex, jac=jex) */
    rc = bvode_create(&solver, bvode_problem_lookup("robertson"), nsys, 0, flags);
    if (rc == BVODE_OK) rc = bvode_set_params(solver, k);
    if (rc != BVODE_OK) { fprintf(stderr, "bvode: %s\n", bvode_last_error()); return 1; }
    for (iout = 1; iout <= 12; iout++) {
        /* call solver%solve(neq,y,t,tout,itol,rtol,atol,itask,istate,iopt,rwork,lrw,iwork,liw,mf) */
        rc = bvode_solve(solver, neq, y, t, &tout, 0, itol, rtol, atol, itask, istate, iopt, rwork, lrw,
                         iwork, liw, mf);
        if (rc != BVODE_OK) { fprintf(stderr, "bvode: %s\n", bvode_last_error()); return 1; }
        printf(" at t = %.17g   y = %.17g %.17g %.17g   | %.17g %.17g %.17g\n", t[0], y[0], y[1], y[2],
               y[3 * (nsys - 1)], y[3 * (nsys - 1) + 1], y[3 * (nsys - 1) + 2]);
        for (i = 0; i < nsys; i++)
            if (istate[i] < 0) {
                char msg[512];
                bvode_format_message(solver, i, msg, sizeof msg);
                printf("\n\n\n error halt: system %d istate = %d\n%s\n", i, istate[i], msg);
                return 2;
            }
        tout = tout * 10.0;
    }
    printf("\n no. steps = %d   no. f-s = %d   no. j-s = %d   no. lu-s = %d\n"
           "  no. nonlinear iterations = %d\n  no. nonlinear convergence failures = %d\n"
           "  no. error test failures = %d\n",
           iwork[10], iwork[11], iwork[12], iwork[18], iwork[19], iwork[20], iwork[21]);
    bvode_destroy(solver);
    return 0;
}
