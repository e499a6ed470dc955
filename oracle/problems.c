/*
 * problems.c -- CPU restatement of the reference's test problems, plus the batch
 * driver that runs one independent oracle solver instance per system.
 *
 * TEST INFRASTRUCTURE ONLY (see dvode_oracle.h).
 *
 *   problem 0  Robertson kinetics         /root/reference/test/example.f90:79-109
 *   problem 1  Van der Pol                /root/reference/test/dvdemo.f90:207-228
 *   problem 2  banded 2-D advection       /root/reference/test/dvdemo.f90:230-274
 *   problem 3  synthetic 32-species mass-action network (builder-defined, not in
 *              the reference; tables shared with the device build through
 *              include/bvode_mech32.h)
 *   problem 7  Van der Pol with a user dewset whose first weight turns non-positive once
 *              |y_1| < 1/2 (builder-defined): drives the `ewt(i) <= 0` exit, istate = -6
 *              (M:1516-1521), which the default weights only reach when a component is exactly 0
 */
#include "dvode_oracle.h"
#include "oracle_batch.h"
#include "../include/bvode_mech32.h"
#include <tgmath.h> /* type-generic fabs / sqrt / copysign: float functions in the REAL32 build, as in C++ */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------- Robertson ---- */
/* par = (k1, k2, k3); reference values 0.04, 1e4, 3e7 (example.f90:87-89) */
static void rob_f(void *ctx, int neq, real t, const real *y, real *ydot)
{
    const real *k = (const real *)ctx;
    (void)neq; (void)t;
    ydot[0] = -k[0] * y[0] + k[1] * y[1] * y[2];
    ydot[2] = k[2] * y[1] * y[1];
    ydot[1] = -ydot[0] - ydot[2];
}
static void rob_jac(void *ctx, int neq, real t, real *y, int ml, int mu, real *pd,
                    int nrowpd)
{
    const real *k = (const real *)ctx;
    (void)neq; (void)t; (void)ml; (void)mu;
#define PD(i, j) pd[((i)-1) + ((j)-1) * nrowpd]
    PD(1, 1) = -k[0];
    PD(1, 2) = k[1] * y[2];
    PD(1, 3) = k[1] * y[1];
    PD(2, 1) = k[0];
    PD(2, 3) = -PD(1, 3);
    PD(3, 2) = (DVO_R(2.0) * k[2]) * y[1];
    PD(2, 2) = -PD(1, 2) - PD(3, 2);
}

/* -------------------------------------------------------- Van der Pol ---- */
/* par = (eta), reference value 3 (dvdemo.f90:214,227) */
static void vdp_f(void *ctx, int neq, real t, const real *y, real *ydot)
{
    const real *p = (const real *)ctx;
    (void)neq; (void)t;
    ydot[0] = y[1];
    ydot[1] = p[0] * (DVO_R(1.0) - y[0] * y[0]) * y[1] - y[0];
}
static void vdp_jac(void *ctx, int neq, real t, real *y, int ml, int mu, real *pd,
                    int nrowpd)
{
    const real *p = (const real *)ctx;
    (void)neq; (void)t; (void)ml; (void)mu;
    PD(1, 1) = DVO_R(0.0);
    PD(1, 2) = DVO_R(1.0);
    PD(2, 1) = -(DVO_R(2.0) * p[0]) * y[0] * y[1] - DVO_R(1.0);
    PD(2, 2) = p[0] * (DVO_R(1.0) - y[0] * y[0]);
}

/* Van der Pol with USER error weights and norm (initialize's optional dewset / dvnorm,
 * M:431-503): ewt_i = rtol_i*max(|y_i|, 1) + atol_i, max norm. */
static void vdpm_dewset(void *ctx, int n, int itol, const real *rtol, const real *atol,
                        const real *ycur, real *ewt)
{
    (void)ctx;
    for (int i = 0; i < n; i++) {
        real r = (itol >= 3) ? rtol[i] : rtol[0];
        real a = (itol == 2 || itol == 4) ? atol[i] : atol[0];
        real ay = fabs(ycur[i]);
        if (ay < DVO_R(1.0)) ay = DVO_R(1.0);
        ewt[i] = r * ay + a;
    }
}
/* problem 7: the default weights, except that ewt_1 drops by one once |y_1| < 1/2 */
static void vdpd_dewset(void *ctx, int n, int itol, const real *rtol, const real *atol,
                        const real *ycur, real *ewt)
{
    (void)ctx;
    for (int i = 0; i < n; i++) {
        real r = (itol >= 3) ? rtol[i] : rtol[0];
        real a = (itol == 2 || itol == 4) ? atol[i] : atol[0];
        ewt[i] = r * fabs(ycur[i]) + a;
        if (i == 0 && fabs(ycur[i]) < DVO_R(0.5)) ewt[i] = ewt[i] - DVO_R(1.0);
    }
}
static real vdpm_dvnorm(void *ctx, int n, const real *v, const real *w)
{
    real vm = DVO_R(0.0);
    (void)ctx;
    for (int i = 0; i < n; i++) {
        real p = fabs(v[i] * w[i]);
        if (p > vm) vm = p;
    }
    return vm;
}

/* ---------------------------------------------------- banded advection --- */
/* par = (alph1, alph2), reference 1, 1; ng = 5 (dvdemo.f90:236-238) */
#define NG 5
static void adv_f(void *ctx, int neq, real t, const real *y, real *ydot)
{
    const real *p = (const real *)ctx;
    (void)t;
    for (int j = 1; j <= NG; j++)
        for (int i = 1; i <= NG; i++) {
            int k = i + (j - 1) * NG;
            if (k > neq) continue;  /* (a call with a reduced neq, M:1127-1135: ydot has neq entries) */
            real d = -DVO_R(2.0) * y[k - 1];
            if (i != 1) d = d + y[k - 2] * p[0];
            if (j != 1) d = d + y[k - NG - 1] * p[1];
            ydot[k - 1] = d;
        }
}
static void adv_jac(void *ctx, int neq, real t, real *y, int ml, int mu, real *pd,
                    int nrowpd)
{
    const real *p = (const real *)ctx;
    (void)t; (void)y;
    int mband = ml + mu + 1, mu1 = mu + 1, mu2 = mu + 2;
    for (int j = 1; j <= neq; j++) {
        PD(mu1, j) = -DVO_R(2.0);
        PD(mu2, j) = p[0];
        PD(mband, j) = p[1];
    }
    for (int j = NG; j <= neq; j += NG) PD(mu2, j) = DVO_R(0.0);
}

/* ------------------------------------------- synthetic 32-species network - */
/* par = 64 rate constants.  ydot_i = sum over the species' entry list, in
 * reaction order, of sign * rate_r (same order as the device functor). */
static void mech_f(void *ctx, int neq, real t, const real *y, real *ydot)
{
    const real *k = (const real *)ctx;
    real rate[BVODE_MECH_NR];
    (void)t;
    for (int r = 0; r < BVODE_MECH_NR; r++) {
        real q = k[r] * y[bvode_mech_ra[r]];
        if (bvode_mech_rb[r] >= 0) q = q * y[bvode_mech_rb[r]];
        rate[r] = q;
    }
    for (int i = 0; i < BVODE_MECH_NS && i < neq; i++) {  /* (neq may have been reduced: ydot has neq entries) */
        real d = DVO_R(0.0);
        for (int e = 0; e < BVODE_MECH_MAXE; e++) {
            int r = bvode_mech_ent_r[i][e];
            if (r < 0) break;
            d = d + bvode_mech_ent_s[i][e] * rate[r];
        }
        ydot[i] = d;
    }
}
#undef PD

int dvo_problem_info(int problem, int *neq, int *npar)
{
    switch (problem) {
    case 0: *neq = 3; *npar = 3; return 0;
    case 1: *neq = 2; *npar = 1; return 0;
    case 2: *neq = NG * NG; *npar = 2; return 0;
    case 3: *neq = BVODE_MECH_NS; *npar = BVODE_MECH_NR; return 0;
    case 4: *neq = 2; *npar = 1; return 0;
    case 5: *neq = NG * NG; *npar = 2; return 0;  /* advection with the user norms of problem 4 */
    case 6: *neq = 48; *npar = 2; return 0;
    case 7: *neq = 2; *npar = 1; return 0;
    default: return -1;
    }
}

/* analytic Jacobian of the mechanism, entry by entry in the order of mech_f (pd pre-zeroed, M:2947) */
static void mech_jac(void *ctx, int neq, real t, real *y, int ml, int mu, real *pd, int nrowpd)
{
    const real *k = (const real *)ctx;
    (void)t; (void)ml; (void)mu;
    for (int i = 0; i < BVODE_MECH_NS && i < neq; i++) {  /* (neq may have been reduced: pd is neq x neq) */
        for (int e = 0; e < BVODE_MECH_MAXE; e++) {
            int r = bvode_mech_ent_r[i][e];
            if (r < 0) break;
            real s = bvode_mech_ent_s[i][e];
            int a = bvode_mech_ra[r], b = bvode_mech_rb[r];
            if (b < 0) {
                if (a < neq) pd[i + a * nrowpd] = pd[i + a * nrowpd] + s * k[r];
            } else {
                if (a < neq) pd[i + a * nrowpd] = pd[i + a * nrowpd] + s * (k[r] * y[b]);
                if (b < neq) pd[i + b * nrowpd] = pd[i + b * nrowpd] + s * (k[r] * y[a]);
            }
        }
    }
}

/* ------------------------------------------- reaction-diffusion chain, n = 48 --- */
/* builder-defined (dvode_b200/csrc/problems.cuh, Diffusion48): par = (D, k) */
static void dif_f(void *ctx, int neq, real t, const real *y, real *ydot)
{
    const real *p = (const real *)ctx;
    (void)t;
    /* ydot has neq entries (fewer than 48 after an istate = 3 call with a smaller neq, M:1127-1135: savf is followed
       by acor in rwork); y is the caller's array and keeps all 48 */
    for (int i = 0; i < neq; i++) {
        real l = (i > 0) ? y[i - 1] : DVO_R(0.0), r = (i + 1 < 48) ? y[i + 1] : DVO_R(0.0);
        real s = (i < 4) ? DVO_R(1.0) : DVO_R(0.0);
        ydot[i] = p[0] * ((l - DVO_R(2.0) * y[i]) + r) - p[1] * y[i] * y[i] + s;
    }
}

static int problem_fns(int problem, dvo_f_fn *f, dvo_jac_fn *jac)
{
    switch (problem) {
    case 0: *f = rob_f; *jac = rob_jac; return 0;
    case 1: *f = vdp_f; *jac = vdp_jac; return 0;
    case 2: *f = adv_f; *jac = adv_jac; return 0;
    case 3: *f = mech_f; *jac = mech_jac; return 0;
    case 4: *f = vdp_f; *jac = vdp_jac; return 0;
    case 5: *f = adv_f; *jac = adv_jac; return 0;
    case 6: *f = dif_f; *jac = NULL; return 0;
    case 7: *f = vdp_f; *jac = vdp_jac; return 0;
    default: return -1;
    }
}

void dvo_problem_f(int problem, const real *par, real t, const real *y, real *ydot)
{
    dvo_f_fn f; dvo_jac_fn jac; int neq, npar;
    if (problem_fns(problem, &f, &jac) || dvo_problem_info(problem, &neq, &npar)) return;
    f((void *)par, neq, t, y, ydot);
}

/* work array sizes, M:1246-1273 */
void dvo_work_sizes(int neq, int mf, int ml, int mu, int maxord_in, int *lrw, int *liw)
{
    int jsv = mf >= 0 ? 1 : -1, mfa = mf < 0 ? -mf : mf, meth = mfa / 10,
        miter = mfa - 10 * meth;
    int maxord = (meth == 1) ? 12 : 5;
    if (maxord_in > 0 && maxord_in < maxord) maxord = maxord_in;
    int jco = jsv > 0 ? 1 : 0, lenwm = 0;
    switch (miter) {
    case 0: lenwm = 0; break;
    case 1: case 2: lenwm = 2 + (1 + jco) * neq * neq; break;
    case 3: lenwm = 2 + neq; break;
    default: lenwm = 2 + (2 * ml + mu + 1) * neq + jco * (ml + mu + 1) * neq; break;
    }
    *lrw = 20 + neq * (maxord + 1) + lenwm + 3 * neq;
    *liw = (miter == 0 || miter == 3) ? 30 : 30 + neq;
}

/* threads the last batch call actually ran on (bench.py reports it as `cores`) */
static int g_last_threads = 1;
int dvo_last_threads(void) { return g_last_threads; }

/* One independent solver instance per system, replaying the reference's call
 * sequence (example.f90:56-66): for each tout, one dvode call with itask as
 * configured; istate carried between calls; stop at the first negative istate. */
int dvo_batch_solve(const dvo_batch_cfg *cfg, int nsys, const real *params, const real *y0,
                    real t0, int nout, const real *touts, const real *rtol,
                    const real *atol, real *y_out, real *t_out, int *istate_out,
                    real *rstats, int *istats, int nthreads)
{
    dvo_f_fn f; dvo_jac_fn jac; int neq, npar;
    if (problem_fns(cfg->problem, &f, &jac) || dvo_problem_info(cfg->problem, &neq, &npar))
        return -1;
    int lrw, liw;
    dvo_work_sizes(neq, cfg->mf, cfg->ml, cfg->mu, cfg->iopt ? cfg->maxord : 0, &lrw, &liw);
    if (liw < 30) liw = 30;
    int fail = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
#ifdef _OPENMP
#pragma omp single nowait
        g_last_threads = omp_get_num_threads();
#endif
        real *rwork = (real *)malloc(sizeof(real) * (size_t)lrw);
        int *iwork = (int *)malloc(sizeof(int) * (size_t)liw);
        real *y = (real *)malloc(sizeof(real) * (size_t)neq);
        real *par = (real *)malloc(sizeof(real) * (size_t)(npar > 0 ? npar : 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int s = 0; s < nsys; s++) {
            dvo_solver sol;
            memcpy(par, params + (size_t)s * npar, sizeof(real) * (size_t)npar);
            dvo_initialize(&sol, f, jac, par);
            if (cfg->problem == 4 || cfg->problem == 5) dvo_set_user_norms(&sol, vdpm_dewset, vdpm_dvnorm);
            if (cfg->problem == 7) dvo_set_user_norms(&sol, vdpd_dewset, NULL);
            sol.pow_mode = cfg->pow_mode;
            memset(rwork, 0, sizeof(real) * (size_t)lrw);
            memset(iwork, 0, sizeof(int) * (size_t)liw);
            memcpy(y, y0 + (size_t)s * neq, sizeof(real) * (size_t)neq);
            iwork[0] = cfg->ml;
            iwork[1] = cfg->mu;
            if (cfg->iopt) {
                rwork[4] = cfg->h0; rwork[5] = cfg->hmax; rwork[6] = cfg->hmin;
                iwork[4] = cfg->maxord; iwork[5] = cfg->mxstep; iwork[6] = cfg->mxhnil;
            }
            rwork[0] = cfg->tcrit;
            real t = t0;
            int istate = 1;
            for (int io = 0; io < nout; io++) {
                if (istate >= 0)
                    dvo_solve(&sol, neq, y, &t, touts[io], cfg->itol, rtol, atol, cfg->itask,
                              &istate, cfg->iopt, rwork, lrw, iwork, liw, cfg->mf);
                size_t o = (size_t)s * nout + io;
                memcpy(y_out + o * neq, y, sizeof(real) * (size_t)neq);
                if (t_out) t_out[o] = t;
                if (istate_out) istate_out[o] = istate;
                if (rstats) memcpy(rstats + o * 4, rwork + 10, sizeof(real) * 4);
                if (istats) memcpy(istats + o * 12, iwork + 10, sizeof(int) * 12);
            }
            if (istate < 0) {
#ifdef _OPENMP
#pragma omp atomic
#endif
                fail++;
            }
        }
        free(rwork); free(iwork); free(y); free(par);
    }
    return fail;
}

/* neq of the calls from the istate = 3 call on in dvo_batch_solve2 (0 = unchanged): the reference allows a
 * smaller neq there (M:1127-1135); the dropped components keep their last values in the output rows. */
static int g_neq2 = 0;
void dvo_set_neq2(int neq2) { g_neq2 = neq2; }

/* Two-phase run: outputs [0, nsplit) with cfg / rtol / atol, then ONE call with istate = 3 and
 * cfg2 / rtol2 / atol2 (same mf; the reference's "parameter change" restart, M:1382-1390) and
 * ordinary continuation calls for the remaining outputs. */
int dvo_batch_solve2(const dvo_batch_cfg *cfg, const dvo_batch_cfg *cfg2, int nsplit, int nsys,
                     const real *params, const real *y0, real t0, int nout,
                     const real *touts, const real *rtol, const real *atol,
                     const real *rtol2, const real *atol2, real *y_out, real *t_out,
                     int *istate_out, real *rstats, int *istats, int nthreads)
{
    dvo_f_fn f; dvo_jac_fn jac; int neq, npar;
    if (problem_fns(cfg->problem, &f, &jac) || dvo_problem_info(cfg->problem, &neq, &npar))
        return -1;
    int lrw, liw;
    dvo_work_sizes(neq, cfg->mf, cfg->ml, cfg->mu, cfg->iopt ? cfg->maxord : 0, &lrw, &liw);
    {   /* the istate = 3 call may carry another method flag (M:999-1046): the larger work arrays */
        int lrw2, liw2;
        dvo_work_sizes(neq, cfg2->mf, cfg2->ml, cfg2->mu, cfg->iopt ? cfg->maxord : 0, &lrw2, &liw2);
        if (lrw2 > lrw) lrw = lrw2;
        if (liw2 > liw) liw = liw2;
    }
    if (liw < 30) liw = 30;
    int fail = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
#ifdef _OPENMP
#pragma omp single nowait
        g_last_threads = omp_get_num_threads();
#endif
        real *rwork = (real *)malloc(sizeof(real) * (size_t)lrw);
        int *iwork = (int *)malloc(sizeof(int) * (size_t)liw);
        real *y = (real *)malloc(sizeof(real) * (size_t)neq);
        real *par = (real *)malloc(sizeof(real) * (size_t)(npar > 0 ? npar : 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int s = 0; s < nsys; s++) {
            dvo_solver sol;
            memcpy(par, params + (size_t)s * npar, sizeof(real) * (size_t)npar);
            dvo_initialize(&sol, f, jac, par);
            if (cfg->problem == 4 || cfg->problem == 5) dvo_set_user_norms(&sol, vdpm_dewset, vdpm_dvnorm);
            if (cfg->problem == 7) dvo_set_user_norms(&sol, vdpd_dewset, NULL);
            sol.pow_mode = cfg->pow_mode;
            memset(rwork, 0, sizeof(real) * (size_t)lrw);
            memset(iwork, 0, sizeof(int) * (size_t)liw);
            memcpy(y, y0 + (size_t)s * neq, sizeof(real) * (size_t)neq);
            real t = t0;
            int istate = 1;
            for (int io = 0; io < nout; io++) {
                const dvo_batch_cfg *c = (io < nsplit) ? cfg : cfg2;
                const real *rt = (io < nsplit) ? rtol : rtol2, *at = (io < nsplit) ? atol : atol2;
                if (io == nsplit && istate == 2) istate = 3;
                iwork[0] = c->ml;
                iwork[1] = c->mu;
                rwork[4] = c->h0; rwork[5] = c->hmax; rwork[6] = c->hmin;
                iwork[4] = c->maxord; iwork[5] = c->mxstep; iwork[6] = c->mxhnil;
                rwork[0] = c->tcrit;
                if (istate >= 0)
                    dvo_solve(&sol, (io >= nsplit && g_neq2 > 0) ? g_neq2 : neq, y, &t, touts[io], c->itol, rt, at, c->itask, &istate,
                              c->iopt, rwork, lrw, iwork, liw, c->mf);
                size_t o = (size_t)s * nout + io;
                memcpy(y_out + o * neq, y, sizeof(real) * (size_t)neq);
                if (t_out) t_out[o] = t;
                if (istate_out) istate_out[o] = istate;
                if (rstats) memcpy(rstats + o * 4, rwork + 10, sizeof(real) * 4);
                if (istats) memcpy(istats + o * 12, iwork + 10, sizeof(int) * 12);
            }
            if (istate < 0) {
#ifdef _OPENMP
#pragma omp atomic
#endif
                fail++;
            }
        }
        free(rwork); free(iwork); free(y); free(par);
    }
    return fail;
}

/* dvindy at user points after the last solve of a single system: convenience
 * used by the dense-output parity tests.  Integrates system 0 of the batch to
 * tout with one call, then evaluates derivative order k at each tq. */
int dvo_single_dense(const dvo_batch_cfg *cfg, const real *par_in, const real *y0, real t0,
                     real tout, const real *rtol, const real *atol, int k, int nq,
                     const real *tqs, real *dky_out, int *iflag_out, real *tcur,
                     real *hu)
{
    dvo_f_fn f; dvo_jac_fn jac; int neq, npar;
    if (problem_fns(cfg->problem, &f, &jac) || dvo_problem_info(cfg->problem, &neq, &npar))
        return -1;
    int lrw, liw;
    dvo_work_sizes(neq, cfg->mf, cfg->ml, cfg->mu, cfg->iopt ? cfg->maxord : 0, &lrw, &liw);
    real *rwork = (real *)calloc((size_t)lrw, sizeof(real));
    int *iwork = (int *)calloc((size_t)liw, sizeof(int));
    real *y = (real *)malloc(sizeof(real) * (size_t)neq);
    real *par = (real *)malloc(sizeof(real) * (size_t)(npar > 0 ? npar : 1));
    memcpy(par, par_in, sizeof(real) * (size_t)npar);
    memcpy(y, y0, sizeof(real) * (size_t)neq);
    dvo_solver sol;
    dvo_initialize(&sol, f, jac, par);
    if (cfg->problem == 4 || cfg->problem == 5) dvo_set_user_norms(&sol, vdpm_dewset, vdpm_dvnorm);
            if (cfg->problem == 7) dvo_set_user_norms(&sol, vdpd_dewset, NULL);
    sol.pow_mode = cfg->pow_mode;
    iwork[0] = cfg->ml; iwork[1] = cfg->mu;
    if (cfg->iopt) {
        rwork[4] = cfg->h0; rwork[5] = cfg->hmax; rwork[6] = cfg->hmin;
        iwork[4] = cfg->maxord; iwork[5] = cfg->mxstep; iwork[6] = cfg->mxhnil;
    }
    real t = t0;
    int istate = 1;
    dvo_solve(&sol, neq, y, &t, tout, cfg->itol, rtol, atol, cfg->itask, &istate, cfg->iopt,
              rwork, lrw, iwork, liw, cfg->mf);
    if (tcur) *tcur = rwork[12];
    if (hu) *hu = rwork[10];
    for (int q = 0; q < nq; q++)
        dvo_dvindy(&sol, tqs[q], k, rwork + 20, neq, dky_out + (size_t)q * neq, iflag_out + q);
    free(rwork); free(iwork); free(y); free(par);
    return istate;
}
