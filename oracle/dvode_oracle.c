/*
 * dvode_oracle.c -- CPU restatement of jacobwilliams/dvode (REAL64) in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see dvode_oracle.h).  Same operation order as the
 * reference; build with -ffp-contract=off so that no FMA is formed (the
 * reference's gfortran release build on baseline x86-64 has none either).
 *
 * File:line citations are into /root/reference:
 *   M  = src/dvode_module.F90, LP = src/dvode_linpack_module.F90,
 *   BL = src/dvode_blas_module.F90.
 * Fortran labels (50/100/200/300/400/420/450/500) are kept as C labels so the
 * control flow can be compared side by side with the reference.
 */
#include "dvode_oracle.h"
#include <tgmath.h> /* type-generic fabs / sqrt / copysign: float functions in the REAL32 build, as in C++ */
#include <string.h>
#include <stdint.h>

#define ZERO ((real)0.0)
#define ONE ((real)1.0)
static const int mord[3] = {0, 12, 5}; /* M:38 */

static inline real dmax2(real a, real b) { return a > b ? a : b; }
static inline real dmin2(real a, real b) { return a < b ? a : b; }
static inline int imax2(int a, int b) { return a > b ? a : b; }
static inline int imin2(int a, int b) { return a < b ? a : b; }
static inline real dsign(real a, real b) { return copysign(fabs(a), b); }

/* ---------------------------------------------------------------- pow ---- */

/* gfortran lowers real**integer to __builtin_powi -> libgcc __powidf2 */
real dvo_powi(real x, int m)
{
    unsigned int n = m < 0 ? -(unsigned int)m : (unsigned int)m;
    real y = (n % 2) ? x : DVO_R(1.0);
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? DVO_R(1.0) / y : y;
}

/* REALIFY-OFF (double inside whatever `real` is) */
/* Portable x**y for x > 0: exp(y*log(x)) from + - * / and bit moves only, with
 * every operation individually rounded (no FMA), so that the device strict
 * build reproduces it bit for bit.  Accuracy: a few ulp; it is NOT the
 * reference's libm pow and is used only when pow_mode = 1. */
double dvo_ppow(double x, double y)
{
    if (x == 0.0) return 0.0;
    if (x == 1.0 || y == 0.0) return 1.0;
    union { double d; uint64_t u; } c;
    c.d = x;
    int e = (int)((c.u >> 52) & 0x7ff);
    if (e == 0) { /* subnormal: scale up */
        c.d = x * 18014398509481984.0; /* 2^54 */
        e = (int)((c.u >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    c.u = (c.u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL; /* m in [1,2) */
    double m = c.d;
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }        /* m in (0.707,1.414] */
    /* log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716 */
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 25.0;
    p = p * s2 + 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0;
    p = p * s2 + 1.0 / 19.0;
    p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    double lm = 2.0 * s + (2.0 * s) * (s2 * p);
    const double ln2_hi = 6.93147180369123816490e-01; /* 33 bits */
    const double ln2_lo = 1.90821492927058770002e-10;
    double ed = (double)e;
    double lx = (ed * ln2_hi + lm) + ed * ln2_lo;
    double z = y * lx;
    if (z > 709.0) return HUGE_VAL;
    if (z < -745.0) return 0.0;
    /* exp(z): z = k ln2 + r, |r| <= ln2/2 */
    double kd = z * 1.44269504088896338700e+00;
    kd = (kd >= 0.0) ? (double)(long long)(kd + 0.5) : (double)(long long)(kd - 0.5);
    double r = (z - kd * ln2_hi) - kd * ln2_lo;
    /* halve r four times -> |r| <= 0.0217, Taylor degree 10, square four times */
    r = r * 0.0625;
    double q = 1.0 / 3628800.0;
    q = q * r + 1.0 / 362880.0;
    q = q * r + 1.0 / 40320.0;
    q = q * r + 1.0 / 5040.0;
    q = q * r + 1.0 / 720.0;
    q = q * r + 1.0 / 120.0;
    q = q * r + 1.0 / 24.0;
    q = q * r + 1.0 / 6.0;
    q = q * r + 0.5;
    q = q * r + 1.0;
    q = q * r; /* expm1(r) */
    /* (1+q)^2 - 1 = 2q + q^2, repeated to keep the small part accurate */
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    double er = 1.0 + q;
    int k = (int)kd;
    /* scale by 2^k in two steps to stay in range */
    int k1 = k / 2, k2 = k - k1;
    union { double d; uint64_t u; } t1, t2;
    t1.u = (uint64_t)(k1 + 1023) << 52;
    t2.u = (uint64_t)(k2 + 1023) << 52;
    return (er * t1.d) * t2.d;
}

/* REALIFY-ON */
static inline real rpow(const dvo_solver *me, real x, real y)
{
    return me->pow_mode ? (real)dvo_ppow((real)x, (real)y) : (real)pow((real)x, (real)y);
}

/* ------------------------------------------------------------- BLAS-1 ---- */
/* BL:37 daxpy (da == 0 early-out, BL:50), BL:100 dcopy, BL:233 dscal,
 * BL:311 idamax (first index of max |x|, strict '>', BL:334).  Unit stride
 * only -- every call site on the path uses incx = incy = 1. */
static void daxpy(int n, real da, const real *dx, real *dy)
{
    if (n <= 0) return;
    if (da == ZERO) return;
    for (int i = 0; i < n; i++) dy[i] = dy[i] + da * dx[i];
}
static void dcopy(int n, const real *dx, real *dy)
{
    for (int i = 0; i < n; i++) dy[i] = dx[i];
}
static void dscal(int n, real da, real *dx)
{
    for (int i = 0; i < n; i++) dx[i] = da * dx[i];
}
static int idamax(int n, const real *dx) /* returns 1-based index */
{
    if (n <= 0) return 0;
    int imax = 1;
    if (n == 1) return imax;
    real dmax = fabs(dx[0]);
    for (int i = 2; i <= n; i++) {
        real xmag = fabs(dx[i - 1]);
        if (xmag > dmax) { imax = i; dmax = xmag; }
    }
    return imax;
}

/* ------------------------------------------------------------ LINPACK ---- */
#define A(i, j) a[((i)-1) + (size_t)((j)-1) * lda]

/* LP:46-120 */
/* Test instrumentation (not part of the restated algorithm): how many times dvjac reported a singular
 * matrix since the last reset, so that a test can show it really drove the integrator down that path. */
static int g_singular_events = 0;
void dvo_singular_events_add(void)
{
#ifdef _OPENMP
#pragma omp atomic
#endif
    g_singular_events++;
}
int dvo_singular_events(int reset)
{
    int v = g_singular_events;
    if (reset) g_singular_events = 0;
    return v;
}

void dvo_dgefa(real *a, int lda, int n, int *ipvt, int *info)
{
    real t;
    int j, k, kp1, l, nm1;
    *info = 0;
    nm1 = n - 1;
    if (nm1 >= 1) {
        for (k = 1; k <= nm1; k++) {
            kp1 = k + 1;
            l = idamax(n - k + 1, &A(k, k)) + k - 1;
            ipvt[k - 1] = l;
            if (A(l, k) == DVO_R(0.0)) {
                *info = k;
            } else {
                if (l != k) { t = A(l, k); A(l, k) = A(k, k); A(k, k) = t; }
                t = -DVO_R(1.0) / A(k, k);
                dscal(n - k, t, &A(k + 1, k));
                for (j = kp1; j <= n; j++) {
                    t = A(l, j);
                    if (l != k) { A(l, j) = A(k, j); A(k, j) = t; }
                    daxpy(n - k, t, &A(k + 1, k), &A(k + 1, j));
                }
            }
        }
    }
    ipvt[n - 1] = n;
    if (A(n, n) == DVO_R(0.0)) *info = n;
}

/* LP:157-228, job = 0 branch (LP:206-225) */
void dvo_dgesl(const real *a, int lda, int n, const int *ipvt, real *b)
{
    real t;
    int k, kb, l, nm1;
    nm1 = n - 1;
    if (nm1 >= 1) {
        for (k = 1; k <= nm1; k++) {
            l = ipvt[k - 1];
            t = b[l - 1];
            if (l != k) { b[l - 1] = b[k - 1]; b[k - 1] = t; }
            daxpy(n - k, t, &A(k + 1, k), &b[k]);
        }
    }
    for (kb = 1; kb <= n; kb++) {
        k = n + 1 - kb;
        b[k - 1] = b[k - 1] / A(k, k);
        t = -b[k - 1];
        daxpy(k - 1, t, &A(1, k), &b[0]);
    }
}
#undef A

#define ABD(i, j) abd[((i)-1) + (size_t)((j)-1) * lda]
/* LP:275-396 */
void dvo_dgbfa(real *abd, int lda, int n, int ml, int mu, int *ipvt, int *info)
{
    real t;
    int i, i0, j, ju, jz, j0, j1, k, kp1, l, lm, m, mm, nm1;
    m = ml + mu + 1;
    *info = 0;
    j0 = mu + 2;
    j1 = imin2(n, m) - 1;
    if (j1 >= j0) {
        for (jz = j0; jz <= j1; jz++) {
            i0 = m + 1 - jz;
            for (i = i0; i <= ml; i++) ABD(i, jz) = DVO_R(0.0);
        }
    }
    jz = j1;
    ju = 0;
    nm1 = n - 1;
    if (nm1 >= 1) {
        for (k = 1; k <= nm1; k++) {
            kp1 = k + 1;
            jz = jz + 1;
            if (jz <= n) {
                if (ml >= 1)
                    for (i = 1; i <= ml; i++) ABD(i, jz) = DVO_R(0.0);
            }
            lm = imin2(ml, n - k);
            l = idamax(lm + 1, &ABD(m, k)) + m - 1;
            ipvt[k - 1] = l + k - m;
            if (ABD(l, k) == DVO_R(0.0)) {
                *info = k;
            } else {
                if (l != m) { t = ABD(l, k); ABD(l, k) = ABD(m, k); ABD(m, k) = t; }
                t = -DVO_R(1.0) / ABD(m, k);
                dscal(lm, t, &ABD(m + 1, k));
                ju = imin2(imax2(ju, mu + ipvt[k - 1]), n);
                mm = m;
                if (ju >= kp1) {
                    for (j = kp1; j <= ju; j++) {
                        l = l - 1;
                        mm = mm - 1;
                        t = ABD(l, j);
                        if (l != mm) { ABD(l, j) = ABD(mm, j); ABD(mm, j) = t; }
                        daxpy(lm, t, &ABD(m + 1, k), &ABD(mm + 1, j));
                    }
                }
            }
        }
    }
    ipvt[n - 1] = n;
    if (ABD(m, n) == DVO_R(0.0)) *info = n;
}

/* LP:435-522, job = 0 branch (LP:494-519) */
void dvo_dgbsl(const real *abd, int lda, int n, int ml, int mu, const int *ipvt, real *b)
{
    real t;
    int k, kb, l, la, lb, lm, m, nm1;
    m = mu + ml + 1;
    nm1 = n - 1;
    if (ml != 0) {
        if (nm1 >= 1) {
            for (k = 1; k <= nm1; k++) {
                lm = imin2(ml, n - k);
                l = ipvt[k - 1];
                t = b[l - 1];
                if (l != k) { b[l - 1] = b[k - 1]; b[k - 1] = t; }
                daxpy(lm, t, &ABD(m + 1, k), &b[k]);
            }
        }
    }
    for (kb = 1; kb <= n; kb++) {
        k = n + 1 - kb;
        b[k - 1] = b[k - 1] / ABD(m, k);
        lm = imin2(k, m) - 1;
        la = m - lm;
        lb = k - lm;
        t = -b[k - 1];
        daxpy(lm, t, &ABD(la, k), &b[lb - 1]);
    }
}
#undef ABD

/* ------------------------------------------------ dewset / dvnorm -------- */
/* M:3238-3273 */
static void dewset_default(int n, int itol, const real *rtol, const real *atol,
                           const real *ycur, real *ewt)
{
    int i;
    switch (itol) {
    case 2: for (i = 0; i < n; i++) ewt[i] = rtol[0] * fabs(ycur[i]) + atol[i]; return;
    case 3: for (i = 0; i < n; i++) ewt[i] = rtol[i] * fabs(ycur[i]) + atol[0]; return;
    case 4: for (i = 0; i < n; i++) ewt[i] = rtol[i] * fabs(ycur[i]) + atol[i]; return;
    default: break;
    }
    for (i = 0; i < n; i++) ewt[i] = rtol[0] * fabs(ycur[i]) + atol[0];
}

/* M:3295-3311 */
static real dvnorm_default(int n, const real *v, const real *w)
{
    real sum = ZERO;
    for (int i = 0; i < n; i++) {
        real p = v[i] * w[i];
        sum = sum + p * p;
    }
    return sqrt(sum / (real)n);
}

/* the procedure pointers of dvode_t (M:305-306): user routine or default */
#define dewset(n, itol, rtol, atol, ycur, ewt)                                        \
    do {                                                                              \
        if (me->dewset) me->dewset(me->ctx, (n), (itol), (rtol), (atol), (ycur), (ewt)); \
        else dewset_default((n), (itol), (rtol), (atol), (ycur), (ewt));              \
    } while (0)
#define dvnorm(n, v, w) (me->dvnorm ? me->dvnorm(me->ctx, (n), (v), (w)) : dvnorm_default((n), (v), (w)))

/* ------------------------------------------------------------- init ------ */
void dvo_initialize(dvo_solver *s, dvo_f_fn f, dvo_jac_fn jac, void *ctx)
{
    memset(s, 0, sizeof(*s));
    s->f = f;
    s->jac = jac;
    s->ctx = ctx;
}

void dvo_set_user_norms(dvo_solver *s, dvo_ewt_fn ew, dvo_norm_fn nm)
{
    s->dewset = ew;
    s->dvnorm = nm;
}

void dvo_dvsrco(dvo_solver *s, dvo_solver *sav, int job)
{
    if (job == 1) *sav = *s;
    else if (job == 2) {
        dvo_f_fn f = s->f; dvo_jac_fn jac = s->jac; void *ctx = s->ctx;
        *s = *sav;
        s->f = f; s->jac = jac; s->ctx = ctx;
    }
}

/* ------------------------------------------------------------- dvhin ----- */
/* M:1738-1851.  All arrays 0-based here. */
static void dvhin(dvo_solver *me, int n, real t0, const real *y0, const real *ydot,
                  real tout, real uround, const real *ewt, int itol,
                  const real *atol, real *y, real *temp, real *h0, int *niter,
                  int *ier)
{
    const real half = DVO_R(0.5), hun = DVO_R(100.0), pt1 = DVO_R(0.1), two = DVO_R(2.0);
    real afi, atoli, delyi, h, hg, hlb, hnew = DVO_R(0.0), hrat, hub, t1, tdist, tround, yddnrm;
    int i, iter;

    *niter = 0;
    tdist = fabs(tout - t0);
    tround = uround * dmax2(fabs(t0), fabs(tout));
    if (tdist < two * tround) { *ier = -1; return; }
    hlb = hun * tround;
    hub = pt1 * tdist;
    atoli = atol[0];
    for (i = 0; i < n; i++) {
        if (itol == 2 || itol == 4) atoli = atol[i];
        delyi = pt1 * fabs(y0[i]) + atoli;
        afi = fabs(ydot[i]);
        if (afi * hub > delyi) hub = delyi / afi;
    }
    iter = 0;
    hg = sqrt(hlb * hub);
    if (hub < hlb) {
        *h0 = hg;
    } else {
        for (;;) {
            h = dsign(hg, tout - t0);
            t1 = t0 + h;
            for (i = 0; i < n; i++) y[i] = y0[i] + h * ydot[i];
            me->f(me->ctx, n, t1, y, temp);
            for (i = 0; i < n; i++) temp[i] = (temp[i] - ydot[i]) / h;
            yddnrm = dvnorm(n, temp, ewt);
            if (yddnrm * hub * hub > two) hnew = sqrt(two / yddnrm);
            else hnew = sqrt(hg * hub);
            iter = iter + 1;
            if (iter < 4) {
                hrat = hnew / hg;
                if ((hrat <= half) || (hrat >= two)) {
                    if ((iter >= 2) && (hnew > two * hg)) { hnew = hg; break; }
                    hg = hnew;
                    continue;
                }
            }
            break;
        }
        *h0 = hnew * half;
        if (*h0 < hlb) *h0 = hlb;
        if (*h0 > hub) *h0 = hub;
    }
    *h0 = dsign(*h0, tout - t0);
    *niter = iter;
    *ier = 0;
}

/* ------------------------------------------------------------ dvindy ----- */
/* M:1878-1959 */
void dvo_dvindy(dvo_solver *me, real t, int k, const real *yh, int ldyh, real *dky,
                int *iflag)
{
#define YH(i, j) yh[((i)-1) + (size_t)((j)-1) * ldyh]
    const real hun = DVO_R(100.0);
    real c, r, s, tfuzz, tn1, tp;
    int i, ic, j, jb, jb2, jj, jj1, jp1;

    *iflag = 0;
    if (k < 0 || k > me->nq) { me->last_msg = 51; *iflag = -1; return; }
    tfuzz = hun * me->uround * dsign(fabs(me->tn) + fabs(me->hu), me->hu);
    tp = me->tn - me->hu - tfuzz;
    tn1 = me->tn + tfuzz;
    if ((t - tp) * (t - tn1) > ZERO) { me->last_msg = 52; *iflag = -2; return; }
    s = (t - me->tn) / me->h;
    ic = 1;
    if (k != 0) {
        jj1 = me->l - k;
        for (jj = jj1; jj <= me->nq; jj++) ic = ic * jj;
    }
    c = (real)ic;
    for (i = 1; i <= me->n; i++) dky[i - 1] = c * YH(i, me->l);
    if (k != me->nq) {
        jb2 = me->nq - k;
        for (jb = 1; jb <= jb2; jb++) {
            j = me->nq - jb;
            jp1 = j + 1;
            ic = 1;
            if (k != 0) {
                jj1 = jp1 - k;
                for (jj = jj1; jj <= j; jj++) ic = ic * jj;
            }
            c = (real)ic;
            for (i = 1; i <= me->n; i++) dky[i - 1] = c * YH(i, jp1) + s * dky[i - 1];
        }
        if (k == 0) return;
    }
    r = dvo_powi(me->h, -k);
    dscal(me->n, r, dky);
#undef YH
}

/* ------------------------------------------------------------- dvset ----- */
/* M:2422-2561.  el, tau, tq are used 1-based (element 0 unused). */
static void dvset(dvo_solver *me)
{
    const real cortes = DVO_R(0.1), one = DVO_R(1.0), six = DVO_R(6.0), two = DVO_R(2.0);
    real ahatn0, alph0, cnqm1, csum, elp, em[14], em0, floti, flotl, flotnq, hsum, rxi,
        rxis, s, t1, t2, t3, t4, t5, t6, xi;
    int i, iback, j, jp1, nqm1, nqm2;
    real *el = me->el, *tau = me->tau, *tq = me->tq;

    flotl = (real)me->l;
    nqm1 = me->nq - 1;
    nqm2 = me->nq - 2;
    if (me->meth == 2) {
        for (i = 3; i <= me->l; i++) el[i] = ZERO;
        el[1] = one;
        el[2] = one;
        alph0 = -one;
        ahatn0 = -one;
        hsum = me->h;
        rxi = one;
        rxis = one;
        if (me->nq != 1) {
            for (j = 1; j <= nqm2; j++) {
                hsum = hsum + tau[j];
                rxi = me->h / hsum;
                jp1 = j + 1;
                alph0 = alph0 - one / (real)jp1;
                for (iback = 1; iback <= jp1; iback++) {
                    i = (j + 3) - iback;
                    el[i] = el[i] + el[i - 1] * rxi;
                }
            }
            alph0 = alph0 - one / (real)me->nq;
            rxis = -el[2] - alph0;
            hsum = hsum + tau[nqm1];
            rxi = me->h / hsum;
            ahatn0 = -el[2] - rxi;
            for (iback = 1; iback <= me->nq; iback++) {
                i = (me->nq + 2) - iback;
                el[i] = el[i] + el[i - 1] * rxis;
            }
        }
        t1 = one - ahatn0 + alph0;
        t2 = one + (real)me->nq * t1;
        tq[2] = fabs(alph0 * t2 / t1);
        tq[5] = fabs(t2 / (el[me->l] * rxi / rxis));
        if (me->nqwait == 1) {
            cnqm1 = rxis / el[me->l];
            t3 = alph0 + one / (real)me->nq;
            t4 = ahatn0 + rxi;
            elp = t3 / (one - t4 + t3);
            tq[1] = fabs(elp / cnqm1);
            hsum = hsum + tau[me->nq];
            rxi = me->h / hsum;
            t5 = alph0 - one / (real)(me->nq + 1);
            t6 = ahatn0 - rxi;
            elp = t2 / (one - t6 + t5);
            tq[3] = fabs(elp * rxi * (flotl + one) * t5);
        }
    } else if (me->nq != 1) {
        /* Adams, M:2492-2550 */
        hsum = me->h;
        em[1] = one;
        flotnq = flotl - one;
        for (i = 2; i <= me->l; i++) em[i] = ZERO;
        for (j = 1; j <= nqm1; j++) {
            if ((j == nqm1) && (me->nqwait == 1)) {
                s = one;
                csum = ZERO;
                for (i = 1; i <= nqm1; i++) {
                    csum = csum + s * em[i] / (real)(i + 1);
                    s = -s;
                }
                tq[1] = em[nqm1] / (flotnq * csum);
            }
            rxi = me->h / hsum;
            for (iback = 1; iback <= j; iback++) {
                i = (j + 2) - iback;
                em[i] = em[i] + em[i - 1] * rxi;
            }
            hsum = hsum + tau[j];
        }
        s = one;
        em0 = ZERO;
        csum = ZERO;
        for (i = 1; i <= me->nq; i++) {
            floti = (real)i;
            em0 = em0 + s * em[i] / floti;
            csum = csum + s * em[i] / (floti + one);
            s = -s;
        }
        s = one / em0;
        el[1] = one;
        for (i = 1; i <= me->nq; i++) el[i + 1] = s * em[i] / (real)i;
        xi = hsum / me->h;
        tq[2] = xi * em0 / csum;
        tq[5] = xi / el[me->l];
        if (me->nqwait == 1) {
            rxi = one / xi;
            for (iback = 1; iback <= me->nq; iback++) {
                i = (me->l + 1) - iback;
                em[i] = em[i] + em[i - 1] * rxi;
            }
            s = one;
            csum = ZERO;
            for (i = 1; i <= me->l; i++) {
                csum = csum + s * em[i] / (real)(i + 1);
                s = -s;
            }
            tq[3] = flotl * em0 / csum;
        }
    } else {
        el[1] = one;
        el[2] = one;
        tq[1] = one;
        tq[2] = two;
        tq[3] = six * tq[2];
        tq[5] = one;
    }
    tq[4] = cortes * tq[2];
}

/* ------------------------------------------------------------ dvjust ----- */
/* M:2568-2694 */
static void dvjust(dvo_solver *me, real *yh, int ldyh, int iord)
{
#define YH(i, j) yh[((i)-1) + (size_t)((j)-1) * ldyh]
    const real one = DVO_R(1.0);
    real alph0, alph1, hsum, prod, t1, xi, xiold;
    int i, iback, j, jp1, lp1, nqm1, nqm2, nqp1;
    real *el = me->el, *tau = me->tau;

    if ((me->nq == 2) && (iord != 1)) return;
    nqm1 = me->nq - 1;
    nqm2 = me->nq - 2;
    if (me->meth == 2) {
        if (iord == 1) {
            for (j = 1; j <= me->lmax; j++) el[j] = ZERO;
            el[3] = one;
            alph0 = -one;
            alph1 = one;
            prod = one;
            xiold = one;
            hsum = me->hscal;
            if (me->nq != 1) {
                for (j = 1; j <= nqm1; j++) {
                    jp1 = j + 1;
                    hsum = hsum + tau[jp1];
                    xi = hsum / me->hscal;
                    prod = prod * xi;
                    alph0 = alph0 - one / (real)jp1;
                    alph1 = alph1 + one / xi;
                    for (iback = 1; iback <= jp1; iback++) {
                        i = (j + 4) - iback;
                        el[i] = el[i] * xiold + el[i - 1];
                    }
                    xiold = xi;
                }
            }
            t1 = (-alph0 - alph1) / prod;
            lp1 = me->l + 1;
            for (i = 1; i <= me->n; i++) YH(i, lp1) = t1 * YH(i, me->lmax);
            nqp1 = me->nq + 1;
            for (j = 3; j <= nqp1; j++) daxpy(me->n, el[j], &YH(1, lp1), &YH(1, j));
        } else {
            for (j = 1; j <= me->lmax; j++) el[j] = ZERO;
            el[3] = one;
            hsum = ZERO;
            for (j = 1; j <= nqm2; j++) {
                hsum = hsum + tau[j];
                xi = hsum / me->hscal;
                jp1 = j + 1;
                for (iback = 1; iback <= jp1; iback++) {
                    i = (j + 4) - iback;
                    el[i] = el[i] * xi + el[i - 1];
                }
            }
            for (j = 3; j <= me->nq; j++)
                for (i = 1; i <= me->n; i++) YH(i, j) = YH(i, j) - YH(i, me->l) * el[j];
            return;
        }
    } else if (iord == 1) {
        lp1 = me->l + 1;
        for (i = 1; i <= me->n; i++) YH(i, lp1) = ZERO;
        return;
    } else {
        for (j = 1; j <= me->lmax; j++) el[j] = ZERO;
        el[2] = one;
        hsum = ZERO;
        for (j = 1; j <= nqm2; j++) {
            hsum = hsum + tau[j];
            xi = hsum / me->hscal;
            jp1 = j + 1;
            for (iback = 1; iback <= jp1; iback++) {
                i = (j + 3) - iback;
                el[i] = el[i] * xi + el[i - 1];
            }
        }
        for (j = 2; j <= nqm1; j++) el[j + 1] = (real)me->nq * el[j] / (real)j;
        for (j = 3; j <= me->nq; j++)
            for (i = 1; i <= me->n; i++) YH(i, j) = YH(i, j) - YH(i, me->l) * el[j];
        return;
    }
#undef YH
}

/* ------------------------------------------------------------- dacopy ---- */
/* M:3112-3127 */
static void dacopy(int nrow, int ncol, const real *a, int nrowa, real *b, int nrowb)
{
    for (int ic = 0; ic < ncol; ic++)
        dcopy(nrow, a + (size_t)ic * nrowa, b + (size_t)ic * nrowb);
}

/* -------------------------------------------------------------- dvjac ---- */
/* M:2897-3104.  wm and iwm 1-based via macros; y, ewt, ftem, savf 1-based too. */
static void dvjac(dvo_solver *me, real *y0, const real *yh, int ldyh, const real *ewt0,
                  real *ftem0, const real *savf0, real *wm0, int *iwm0, int *ierpj)
{
#define YH(i, j) yh[((i)-1) + (size_t)((j)-1) * ldyh]
#define WM(i) wm0[(i)-1]
#define IWM(i) iwm0[(i)-1]
#define Y(i) y0[(i)-1]
#define EWT(i) ewt0[(i)-1]
#define FTEM(i) ftem0[(i)-1]
#define SAVF(i) savf0[(i)-1]
    const real one = DVO_R(1.0), thou = DVO_R(1000.0), pt1 = DVO_R(0.1);
    real con, di, fac, hrl1, r, r0, srur, yi, yj, yjj;
    int i, i1, i2, ier, ii, j, j1, jj, jok, lenp = 0, mba, mband, meb1, meband, ml, ml3, mu,
        np1;
    const int n = me->n;

    *ierpj = 0;
    hrl1 = me->h * me->rl1;
    jok = me->jsv;
    if (me->jsv == 1) {
        if (me->nst == 0 || me->nst > me->nslj + me->msbj) jok = -1;
        if (me->icf == 1 && me->drc < me->ccmxj) jok = -1;
        if (me->icf == 2) jok = -1;
    }

    if (jok == -1 && me->miter == 1) {
        me->nje = me->nje + 1;
        me->nslj = me->nst;
        me->jcur = 1;
        lenp = n * n;
        for (i = 1; i <= lenp; i++) WM(i + 2) = ZERO;
        me->jac(me->ctx, n, me->tn, y0, 0, 0, &WM(3), n);
        if (me->jsv == 1) dcopy(lenp, &WM(3), &WM(me->locjs));
    }

    if (jok == -1 && me->miter == 2) {
        me->nje = me->nje + 1;
        me->nslj = me->nst;
        me->jcur = 1;
        fac = dvnorm(n, savf0, ewt0);
        r0 = thou * fabs(me->h) * me->uround * (real)n * fac;
        if (r0 == ZERO) r0 = one;
        srur = WM(1);
        j1 = 2;
        for (j = 1; j <= n; j++) {
            yj = Y(j);
            r = dmax2(srur * fabs(yj), r0 / EWT(j));
            Y(j) = Y(j) + r;
            fac = one / r;
            me->f(me->ctx, n, me->tn, y0, ftem0);
            for (i = 1; i <= n; i++) WM(i + j1) = (FTEM(i) - SAVF(i)) * fac;
            Y(j) = yj;
            j1 = j1 + n;
        }
        me->nfe = me->nfe + n;
        lenp = n * n;
        if (me->jsv == 1) dcopy(lenp, &WM(3), &WM(me->locjs));
    }

    if (jok == 1 && (me->miter == 1 || me->miter == 2)) {
        me->jcur = 0;
        lenp = n * n;
        dcopy(lenp, &WM(me->locjs), &WM(3));
    }

    if (me->miter == 1 || me->miter == 2) {
        con = -hrl1;
        dscal(lenp, con, &WM(3));
        j = 3;
        np1 = n + 1;
        for (i = 1; i <= n; i++) {
            WM(j) = WM(j) + one;
            j = j + np1;
        }
        me->nlu = me->nlu + 1;
        dvo_dgefa(&WM(3), n, n, &IWM(31), &ier);
        if (ier != 0) *ierpj = 1;
        return;
    }

    if (me->miter == 3) {
        me->nje = me->nje + 1;
        me->jcur = 1;
        WM(2) = hrl1;
        r = me->rl1 * pt1;
        for (i = 1; i <= n; i++) Y(i) = Y(i) + r * (me->h * SAVF(i) - YH(i, 2));
        me->f(me->ctx, n, me->tn, y0, &WM(3));
        me->nfe = me->nfe + 1;
        for (i = 1; i <= n; i++) {
            r0 = me->h * SAVF(i) - YH(i, 2);
            di = pt1 * r0 - me->h * (WM(i + 2) - SAVF(i));
            WM(i + 2) = one;
            if (fabs(r0) >= me->uround / EWT(i)) {
                if (fabs(di) == ZERO) { *ierpj = 1; return; }
                WM(i + 2) = pt1 * r0 / di;
            }
        }
        return;
    }

    ml = IWM(1);
    mu = IWM(2);
    ml3 = ml + 3;
    mband = ml + mu + 1;
    meband = mband + ml;
    lenp = meband * n;

    if (jok == -1 && me->miter == 4) {
        me->nje = me->nje + 1;
        me->nslj = me->nst;
        me->jcur = 1;
        for (i = 1; i <= lenp; i++) WM(i + 2) = ZERO;
        me->jac(me->ctx, n, me->tn, y0, ml, mu, &WM(ml3), meband);
        if (me->jsv == 1) dacopy(mband, n, &WM(ml3), meband, &WM(me->locjs), mband);
    }

    if (jok == -1 && me->miter == 5) {
        me->nje = me->nje + 1;
        me->nslj = me->nst;
        me->jcur = 1;
        mba = imin2(mband, n);
        meb1 = meband - 1;
        srur = WM(1);
        fac = dvnorm(n, savf0, ewt0);
        r0 = thou * fabs(me->h) * me->uround * (real)n * fac;
        if (r0 == ZERO) r0 = one;
        for (j = 1; j <= mba; j++) {
            for (i = j; i <= n; i += mband) {
                yi = Y(i);
                r = dmax2(srur * fabs(yi), r0 / EWT(i));
                Y(i) = Y(i) + r;
            }
            me->f(me->ctx, n, me->tn, y0, ftem0);
            for (jj = j; jj <= n; jj += mband) {
                Y(jj) = YH(jj, 1);
                yjj = Y(jj);
                r = dmax2(srur * fabs(yjj), r0 / EWT(jj));
                fac = one / r;
                i1 = imax2(jj - mu, 1);
                i2 = imin2(jj + ml, n);
                ii = jj * meb1 - ml + 2;
                for (i = i1; i <= i2; i++) WM(ii + i) = (FTEM(i) - SAVF(i)) * fac;
            }
        }
        me->nfe = me->nfe + mba;
        if (me->jsv == 1) dacopy(mband, n, &WM(ml3), meband, &WM(me->locjs), mband);
    }

    if (jok == 1) {
        me->jcur = 0;
        dacopy(mband, n, &WM(me->locjs), mband, &WM(ml3), meband);
    }

    con = -hrl1;
    dscal(lenp, con, &WM(3));
    ii = mband + 2;
    for (i = 1; i <= n; i++) {
        WM(ii) = WM(ii) + one;
        ii = ii + meband;
    }
    me->nlu = me->nlu + 1;
    dvo_dgbfa(&WM(3), meband, n, ml, mu, &IWM(31), &ier);
    if (ier != 0) *ierpj = 1;
#undef YH
#undef Y
#undef EWT
#undef FTEM
#undef SAVF
}

/* -------------------------------------------------------------- dvsol ---- */
/* M:3139-3191 */
static void dvsol(dvo_solver *me, real *wm0, int *iwm0, real *x, int *iersl)
{
    const real one = DVO_R(1.0);
    int i, meband, ml, mu;
    real di, hrl1, phrl1, r;
    const int n = me->n;

    *iersl = 0;
    switch (me->miter) {
    case 3:
        phrl1 = WM(2);
        hrl1 = me->h * me->rl1;
        WM(2) = hrl1;
        if (hrl1 != phrl1) {
            r = hrl1 / phrl1;
            for (i = 1; i <= n; i++) {
                di = one - r * (one - one / WM(i + 2));
                if (fabs(di) == ZERO) { *iersl = 1; return; }
                WM(i + 2) = one / di;
            }
        }
        for (i = 1; i <= n; i++) x[i - 1] = WM(i + 2) * x[i - 1];
        break;
    case 4:
    case 5:
        ml = IWM(1);
        mu = IWM(2);
        meband = 2 * ml + mu + 1;
        dvo_dgbsl(&WM(3), meband, n, ml, mu, &IWM(31), x);
        break;
    default:
        dvo_dgesl(&WM(3), n, n, &IWM(31), x);
        break;
    }
}
#undef WM
#undef IWM

/* ------------------------------------------------------------- dvnlsd ---- */
/* M:2703-2879.  y, savf, ewt, acor 0-based; yh 1-based via macro. */
static void dvnlsd(dvo_solver *me, real *y, real *yh, int ldyh, real *vsav, real *savf,
                   const real *ewt, real *acor, int *iwm, real *wm, int *nflag)
{
#define YH(i, j) yh[((i)-1) + (size_t)((j)-1) * ldyh]
    const int maxcor = 3, msbp = 20;
    const real ccmax = DVO_R(0.3), crdown = DVO_R(0.3), rdiv = DVO_R(2.0), one = DVO_R(1.0), two = DVO_R(2.0);
    real cscale, dcon, del = DVO_R(0.0), delp;
    int i, ierpj, iersl, m;
    const int n = me->n;
    (void)vsav;

    if (me->jstart == 0) me->nslp = 0;
    if (*nflag == 0) me->icf = 0;
    if (*nflag == -2) me->ipup = me->miter;
    if ((me->jstart == 0) || (me->jstart == -1)) me->ipup = me->miter;
    if (me->miter == 0) {
        me->crate = one;
    } else {
        me->drc = fabs(me->rc - one);
        if (me->drc > ccmax || me->nst >= me->nslp + msbp) me->ipup = me->miter;
    }

    for (;;) { /* corrector */
        m = 0;
        delp = ZERO;
        dcopy(n, &YH(1, 1), y);
        me->f(me->ctx, n, me->tn, y, savf);
        me->nfe = me->nfe + 1;
        if (me->ipup > 0) {
            dvjac(me, y, yh, ldyh, ewt, acor, savf, wm, iwm, &ierpj);
            me->ipup = 0;
            me->rc = one;
            me->drc = ZERO;
            me->crate = one;
            me->nslp = me->nst;
            if (ierpj != 0) {
                dvo_singular_events_add();  /* test instrumentation: a singular Newton matrix (M:2999) */
                break;
            }
        }
        for (i = 0; i < n; i++) acor[i] = ZERO;

        for (;;) {
            if (me->miter != 0) {
                for (i = 1; i <= n; i++)
                    y[i - 1] = (me->rl1 * me->h) * savf[i - 1] -
                               (me->rl1 * YH(i, 2) + acor[i - 1]);
                dvsol(me, wm, iwm, y, &iersl);
                me->nni = me->nni + 1;
                if (iersl > 0) break;
                if (me->meth == 2 && me->rc != one) {
                    cscale = two / (one + me->rc);
                    dscal(n, cscale, y);
                }
                del = dvnorm(n, y, ewt);
                daxpy(n, one, y, acor);
                for (i = 1; i <= n; i++) y[i - 1] = YH(i, 1) + acor[i - 1];
            } else {
                for (i = 1; i <= n; i++)
                    savf[i - 1] = me->rl1 * (me->h * savf[i - 1] - YH(i, 2));
                for (i = 1; i <= n; i++) y[i - 1] = savf[i - 1] - acor[i - 1];
                del = dvnorm(n, y, ewt);
                for (i = 1; i <= n; i++) y[i - 1] = YH(i, 1) + savf[i - 1];
                dcopy(n, savf, acor);
            }
            if (m != 0) me->crate = dmax2(crdown * me->crate, del / delp);
            dcon = del * dmin2(one, me->crate) / me->tq[4];
            if (dcon <= one) {
                *nflag = 0;
                me->jcur = 0;
                me->icf = 0;
                if (m == 0) me->acnrm = del;
                if (m > 0) me->acnrm = dvnorm(n, acor, ewt);
                return;
            }
            m = m + 1;
            if (m != maxcor) {
                if (m < 2 || del <= rdiv * delp) {
                    delp = del;
                    me->f(me->ctx, n, me->tn, y, savf);
                    me->nfe = me->nfe + 1;
                    continue;
                }
            }
            break;
        }

        if (me->miter != 0 && me->jcur != 1) {
            me->icf = 1;
            me->ipup = me->miter;
        } else {
            break;
        }
    }

    *nflag = -1;
    me->icf = 2;
    me->ipup = me->miter;
#undef YH
}

/* ------------------------------------------------------------- dvstep ---- */
/* M:1975-2377 */
static void dvstep(dvo_solver *me, real *y, real *yh, int ldyh, real *yh1, real *ewt,
                   real *savf, real *vsav, real *acor, real *wm, int *iwm)
{
#define YH(i, j) yh[((i)-1) + (size_t)((j)-1) * ldyh]
#define YH1(i) yh1[(i)-1]
    const int kfc = -3, kfh = -7, mxncf = 10;
    const real one = DVO_R(1.0), addon = DVO_R(1.0e-6), bias1 = DVO_R(6.0), bias2 = DVO_R(6.0), bias3 = DVO_R(10.0),
                 etacf = DVO_R(0.25), etamin = DVO_R(0.1), etamxf = DVO_R(0.2), etamx1 = DVO_R(1.0e4), etamx2 = DVO_R(10.0),
                 etamx3 = DVO_R(10.0), onepsm = DVO_R(1.00001), thresh = DVO_R(1.5);
    real cnquot, ddn, dsm, dup, etaqp1, flotl, r, told;
    int i, i1, i2, iback, j, jb, ncf, nflag;
    const int n = me->n;

    me->kflag = 0;
    told = me->tn;
    ncf = 0;
    me->jcur = 0;
    nflag = 0;

    if (me->jstart <= 0) {
        if (me->jstart == -1) {
            /* M:2044-2079 */
            me->lmax = me->maxord + 1;
            if (n != ldyh) {
                i1 = 1 + (me->newq + 1) * ldyh;
                i2 = (me->maxord + 1) * ldyh;
                if (i1 <= i2)
                    for (i = i1; i <= i2; i++) YH1(i) = ZERO;
            }
            if (me->newq > me->maxord) {
                flotl = (real)me->lmax;
                if (me->maxord < me->nq - 1) {
                    ddn = dvnorm(n, savf, ewt) / me->tq[1];
                    me->eta = one / (rpow(me, bias1 * ddn, one / flotl) + addon);
                }
                if (me->maxord == me->nq && me->newq == me->nq + 1) me->eta = me->etaq;
                if (me->maxord == me->nq - 1 && me->newq == me->nq + 1) {
                    me->eta = me->etaqm1;
                    dvjust(me, yh, ldyh, -1);
                }
                if (me->maxord == me->nq - 1 && me->newq == me->nq) {
                    ddn = dvnorm(n, savf, ewt) / me->tq[1];
                    me->eta = one / (rpow(me, bias1 * ddn, one / flotl) + addon);
                    dvjust(me, yh, ldyh, -1);
                }
                me->eta = dmin2(me->eta, one);
                me->nq = me->maxord;
                me->l = me->lmax;
            }
            if (me->kuth == 1) me->eta = dmin2(me->eta, fabs(me->h / me->hscal));
            if (me->kuth == 0) me->eta = dmax2(me->eta, me->hmin / fabs(me->hscal));
            me->eta = me->eta / dmax2(one, fabs(me->hscal) * me->hmxi * me->eta);
            me->newh = 1;
            me->nqwait = me->l;
            if (me->newq > me->maxord) goto L300;
        } else {
            /* M:2089-2099 */
            me->lmax = me->maxord + 1;
            me->nq = 1;
            me->l = 2;
            me->nqnyh = me->nq * ldyh;
            me->tau[1] = me->h;
            me->prl1 = one;
            me->rc = ZERO;
            me->etamax = etamx1;
            me->nqwait = 2;
            me->hscal = me->h;
            goto L400;
        }
    } else if (me->kuth == 1) {
        me->eta = dmin2(me->eta, me->h / me->hscal);
        me->newh = 1;
    }

    if (me->newh == 0) goto L400;

    if (me->newq < me->nq) {
        dvjust(me, yh, ldyh, -1);
        me->nq = me->newq;
        me->l = me->nq + 1;
        me->nqwait = me->l;
    } else if (me->newq > me->nq) {
        dvjust(me, yh, ldyh, 1);
        me->nq = me->newq;
        me->l = me->nq + 1;
        me->nqwait = me->l;
    }

L300: /* rescale the history array for a change in h by a factor of eta, M:2131-2139 */
    r = one;
    for (j = 2; j <= me->l; j++) {
        r = r * me->eta;
        dscal(n, r, &YH(1, j));
    }
    me->h = me->hscal * me->eta;
    me->hscal = me->h;
    me->rc = me->rc * me->eta;
    me->nqnyh = me->nq * ldyh;

L400: /* predict, M:2147-2158 */
    me->tn = me->tn + me->h;
    i1 = me->nqnyh + 1;
    for (jb = 1; jb <= me->nq; jb++) {
        i1 = i1 - ldyh;
        for (i = i1; i <= me->nqnyh; i++) YH1(i) = YH1(i) + YH1(i + ldyh);
    }
    dvset(me);
    me->rl1 = one / me->el[2];
    me->rc = me->rc * (me->rl1 / me->prl1);
    me->prl1 = me->rl1;

    dvnlsd(me, y, yh, ldyh, vsav, savf, ewt, acor, iwm, wm, &nflag);

    if (nflag == 0) {
        dsm = me->acnrm / me->tq[2];
        if (dsm > one) {
            /* error test failed, M:2177-2237 */
            me->kflag = me->kflag - 1;
            me->netf = me->netf + 1;
            nflag = -2;
            me->tn = told;
            i1 = me->nqnyh + 1;
            for (jb = 1; jb <= me->nq; jb++) {
                i1 = i1 - ldyh;
                for (i = i1; i <= me->nqnyh; i++) YH1(i) = YH1(i) - YH1(i + ldyh);
            }
            if (fabs(me->h) <= me->hmin * onepsm) {
                me->kflag = -1;
                me->jstart = 1;
                return;
            } else {
                me->etamax = one;
                if (me->kflag > kfc) {
                    flotl = (real)me->l;
                    me->eta = one / (rpow(me, bias2 * dsm, one / flotl) + addon);
                    me->eta = dmax2(dmax2(me->eta, me->hmin / fabs(me->h)), etamin);
                    if ((me->kflag <= -2) && (me->eta > etamxf)) me->eta = etamxf;
                    goto L300;
                } else if (me->kflag == kfh) {
                    me->kflag = -1;
                    me->jstart = 1;
                    return;
                } else if (me->nq == 1) {
                    me->eta = dmax2(etamin, me->hmin / fabs(me->h));
                    me->h = me->h * me->eta;
                    me->hscal = me->h;
                    me->tau[1] = me->h;
                    me->f(me->ctx, n, me->tn, y, savf);
                    me->nfe = me->nfe + 1;
                    for (i = 1; i <= n; i++) YH(i, 2) = me->h * savf[i - 1];
                    me->nqwait = 10;
                    goto L400;
                } else {
                    me->eta = dmax2(etamin, me->hmin / fabs(me->h));
                    dvjust(me, yh, ldyh, -1);
                    me->l = me->nq;
                    me->nq = me->nq - 1;
                    me->nqwait = me->l;
                    goto L300;
                }
            }
        } else {
            /* successful step, M:2245-2317 */
            me->kflag = 0;
            me->nst = me->nst + 1;
            me->hu = me->h;
            me->nqu = me->nq;
            for (iback = 1; iback <= me->nq; iback++) {
                i = me->l - iback;
                me->tau[i + 1] = me->tau[i];
            }
            me->tau[1] = me->h;
            for (j = 1; j <= me->l; j++) daxpy(n, me->el[j], acor, &YH(1, j));
            me->nqwait = me->nqwait - 1;
            if ((me->l != me->lmax) && (me->nqwait == 1)) {
                dcopy(n, acor, &YH(1, me->lmax));
                me->conp = me->tq[5];
            }
            if (me->etamax != one) {
                flotl = (real)me->l;
                me->etaq = one / (rpow(me, bias2 * dsm, one / flotl) + addon);
                if (me->nqwait == 0) {
                    me->nqwait = 2;
                    me->etaqm1 = ZERO;
                    if (me->nq != 1) {
                        ddn = dvnorm(n, &YH(1, me->l), ewt) / me->tq[1];
                        me->etaqm1 = one / (rpow(me, bias1 * ddn, one / (flotl - one)) + addon);
                    }
                    etaqp1 = ZERO;
                    if (me->l != me->lmax) {
                        cnquot = (me->tq[5] / me->conp) * dvo_powi(me->h / me->tau[2], me->l);
                        for (i = 1; i <= n; i++)
                            savf[i - 1] = acor[i - 1] - cnquot * YH(i, me->lmax);
                        dup = dvnorm(n, savf, ewt) / me->tq[3];
                        etaqp1 = one / (rpow(me, bias3 * dup, one / (flotl + one)) + addon);
                    }
                    if (me->etaq < etaqp1) {
                        if (etaqp1 <= me->etaqm1) goto L420;
                        me->eta = etaqp1;
                        me->newq = me->nq + 1;
                        dcopy(n, acor, &YH(1, me->lmax));
                        goto L450;
                    } else if (me->etaq < me->etaqm1) {
                        goto L420;
                    }
                }
                me->eta = me->etaq;
                me->newq = me->nq;
                goto L450;
            } else {
                if (me->nqwait < 2) me->nqwait = 2;
                me->newq = me->nq;
                me->newh = 0;
                me->eta = one;
                me->hnew = me->h;
                goto L500;
            }
        L420:
            me->eta = me->etaqm1;
            me->newq = me->nq - 1;
        }

    L450: /* M:2320-2330 */
        if (me->eta < thresh || me->etamax == one) {
            me->newq = me->nq;
            me->newh = 0;
            me->eta = one;
            me->hnew = me->h;
        } else {
            me->eta = dmin2(me->eta, me->etamax);
            me->eta = me->eta / dmax2(one, fabs(me->h) * me->hmxi * me->eta);
            me->newh = 1;
            me->hnew = me->h * me->eta;
        }
    } else {
        /* corrector failed, M:2338-2367 */
        ncf = ncf + 1;
        me->ncfn = me->ncfn + 1;
        me->etamax = one;
        me->tn = told;
        i1 = me->nqnyh + 1;
        for (jb = 1; jb <= me->nq; jb++) {
            i1 = i1 - ldyh;
            for (i = i1; i <= me->nqnyh; i++) YH1(i) = YH1(i) - YH1(i + ldyh);
        }
        if (nflag < -1) {
            if (nflag == -2) me->kflag = -3;
            if (nflag == -3) me->kflag = -4;
            me->jstart = 1;
            return;
        } else if (fabs(me->h) <= me->hmin * onepsm) {
            me->kflag = -2;
            me->jstart = 1;
            return;
        } else if (ncf == mxncf) {
            me->kflag = -2;
            me->jstart = 1;
            return;
        } else {
            me->eta = etacf;
            me->eta = dmax2(me->eta, me->hmin / fabs(me->h));
            nflag = -1;
            goto L300;
        }
    }

L500: /* M:2370-2375 */
    me->etamax = etamx3;
    if (me->nst <= 10) me->etamax = etamx2;
    r = one / me->tq[2];
    dscal(n, r, acor);
    me->jstart = 1;
#undef YH
#undef YH1
}

/* -------------------------------------------------------------- dvode ---- */
/* M:604-1725 */
void dvo_solve(dvo_solver *me, int neq, real *y, real *t, real tout, int itol,
               const real *rtol, const real *atol, int itask, int *istate, int iopt,
               real *rwork, int lrw, int *iwork, int liw, int mf)
{
#define RW(i) rwork[(i)-1]
#define IW(i) iwork[(i)-1]
    const int mxhnl0 = 10, mxstp0 = 500;
    const real one = DVO_R(1.0), two = DVO_R(2.0), four = DVO_R(4.0), pt2 = DVO_R(0.2), hun = DVO_R(100.0);
    int ihit = 0;
    real atoli, big, h0 = DVO_R(0.0), hmax, hmx, rh, rtoli, size, tcrit = DVO_R(0.0), tnext, tolsf, tp;
    int i, ier, iflag, imxer, jco, leniw, lenj, lenp, lenrw, lenwm = 0, lf0, mband, mfa,
        ml = 0, mu = 0, niter, nslast = 0;

#define SET_Y_AND_T()                                \
    do {                                             \
        dcopy(me->n, &RW(me->lyh), y);               \
        *t = me->tn;                                 \
    } while (0)
#define CONTINUE_RETURN()                            \
    do {                                             \
        *istate = 2;                                 \
        RW(11) = me->hu;                             \
        RW(12) = me->hnew;                           \
        RW(13) = me->tn;                             \
        IW(11) = me->nst;                            \
        IW(12) = me->nfe;                            \
        IW(13) = me->nje;                            \
        IW(14) = me->nqu;                            \
        IW(15) = me->newq;                           \
        IW(19) = me->nlu;                            \
        IW(20) = me->nni;                            \
        IW(21) = me->ncfn;                           \
        IW(22) = me->netf;                           \
        return;                                      \
    } while (0)
#define FINISH_RETURN()                              \
    do {                                             \
        SET_Y_AND_T();                               \
        RW(11) = me->hu;                             \
        RW(12) = me->h;                              \
        RW(13) = me->tn;                             \
        IW(11) = me->nst;                            \
        IW(12) = me->nfe;                            \
        IW(13) = me->nje;                            \
        IW(14) = me->nqu;                            \
        IW(15) = me->nq;                             \
        IW(19) = me->nlu;                            \
        IW(20) = me->nni;                            \
        IW(21) = me->ncfn;                           \
        IW(22) = me->netf;                           \
        return;                                      \
    } while (0)
#define COMPUTE_IMXER()                                                    \
    do {                                                                   \
        big = ZERO;                                                        \
        imxer = 1;                                                         \
        for (i = 1; i <= me->n; i++) {                                     \
            size = fabs(RW(i + me->lacor - 1) * RW(i + me->lewt - 1));     \
            if (big < size) { big = size; imxer = i; }                     \
        }                                                                  \
        IW(16) = imxer;                                                    \
    } while (0)
#define ILLEGAL(msgno)                               \
    do {                                             \
        me->last_msg = (msgno);                      \
        *istate = -3;                                \
        return;                                      \
    } while (0)

    me->last_msg = 0;

    /* block a, M:1073-1109 */
    if (*istate < 1 || *istate > 3) {
        me->last_msg = 1;
        if (*istate >= 0) { *istate = -3; return; }
        me->last_msg = 303; /* run aborted: apparent infinite loop (level 2 -> stop) */
        return;
    }
    if (itask < 1 || itask > 5) ILLEGAL(2);
    if (*istate == 1) {
        me->init = 0;
        if (tout == *t) return;
    } else if (me->init != 1) {
        ILLEGAL(3);
    } else if (*istate == 2) {
        goto L50;
    }

    /* block b, M:1120-1303 */
    if (neq <= 0) ILLEGAL(4);
    if (*istate != 1) {
        if (neq > me->n) ILLEGAL(5);
    }
    me->n = neq;
    if (itol < 1 || itol > 4) ILLEGAL(6);
    if (iopt < 0 || iopt > 1) ILLEGAL(7);
    me->jsv = (mf >= 0) ? 1 : -1;
    mfa = mf < 0 ? -mf : mf;
    me->meth = mfa / 10;
    me->miter = mfa - 10 * me->meth;
    if ((me->meth < 1 || me->meth > 2) || (me->miter < 0 || me->miter > 5)) ILLEGAL(8);
    if (me->miter > 3) {
        ml = IW(1);
        mu = IW(2);
        if (ml < 0 || ml >= me->n) ILLEGAL(9);
        if (mu < 0 || mu >= me->n) ILLEGAL(10);
    }
    if (iopt == 1) {
        me->maxord = IW(5);
        if (me->maxord < 0) ILLEGAL(11);
        if (me->maxord == 0) me->maxord = 100;
        me->maxord = imin2(me->maxord, mord[me->meth]);
        me->mxstep = IW(6);
        if (me->mxstep < 0) ILLEGAL(12);
        if (me->mxstep == 0) me->mxstep = mxstp0;
        me->mxhnil = IW(7);
        if (me->mxhnil < 0) ILLEGAL(13);
        if (me->mxhnil == 0) me->mxhnil = mxhnl0;
        if (*istate == 1) {
            h0 = RW(5);
            if ((tout - *t) * h0 < ZERO) ILLEGAL(14);
        }
        hmax = RW(6);
        if (hmax < ZERO) ILLEGAL(15);
        me->hmxi = ZERO;
        if (hmax > ZERO) me->hmxi = one / hmax;
        me->hmin = RW(7);
        if (me->hmin < ZERO) ILLEGAL(16);
    } else {
        me->maxord = mord[me->meth];
        me->mxstep = mxstp0;
        me->mxhnil = mxhnl0;
        if (*istate == 1) h0 = ZERO;
        me->hmxi = ZERO;
        me->hmin = ZERO;
    }
    /* work array pointers, M:1246-1284 */
    me->lyh = 21;
    if (*istate == 1) me->nyh = me->n;
    me->lwm = me->lyh + (me->maxord + 1) * me->nyh;
    jco = imax2(0, me->jsv);
    switch (me->miter) {
    case 0: lenwm = 0; break;
    case 1:
    case 2:
        lenwm = 2 + (1 + jco) * me->n * me->n;
        me->locjs = me->n * me->n + 3;
        break;
    case 3: lenwm = 2 + me->n; break;
    case 4:
    case 5:
        mband = ml + mu + 1;
        lenp = (mband + ml) * me->n;
        lenj = mband * me->n;
        lenwm = 2 + lenp + jco * lenj;
        me->locjs = lenp + 3;
        break;
    }
    me->lewt = me->lwm + lenwm;
    me->lsavf = me->lewt + me->n;
    me->lacor = me->lsavf + me->n;
    lenrw = me->lacor + me->n - 1;
    IW(17) = lenrw;
    me->liwm = 1;
    leniw = 30 + me->n;
    if (me->miter == 0 || me->miter == 3) leniw = 30;
    IW(18) = leniw;
    if (lenrw > lrw) ILLEGAL(17);
    if (leniw > liw) ILLEGAL(18);
    rtoli = rtol[0];
    atoli = atol[0];
    for (i = 1; i <= me->n; i++) {
        if (itol >= 3) rtoli = rtol[i - 1];
        if (itol == 2 || itol == 4) atoli = atol[i - 1];
        if (rtoli < ZERO) ILLEGAL(19);
        if (atoli < ZERO) ILLEGAL(20);
    }
    if (*istate == 1) {
        /* block c, M:1312-1381 */
#ifdef DVO_REAL32
        me->uround = 1.1920929e-07f; /* epsilon(1.0_real32), M:36 with wp = real32 */
#else
        me->uround = DVO_R(2.220446049250313e-16); /* epsilon(1.0_real64), M:36 */
#endif
        me->tn = *t;
        if (itask == 4 || itask == 5) {
            tcrit = RW(1);
            if ((tcrit - tout) * (tout - *t) < ZERO) ILLEGAL(25);
            if (h0 != ZERO && (*t + h0 - tcrit) * h0 > ZERO) h0 = tcrit - *t;
        }
        me->jstart = 0;
        if (me->miter > 0) RW(me->lwm) = sqrt(me->uround);
        me->ccmxj = pt2;
        me->msbj = 50;
        me->nhnil = 0;
        me->nst = 0;
        me->nje = 0;
        me->nni = 0;
        me->ncfn = 0;
        me->netf = 0;
        me->nlu = 0;
        me->nslj = 0;
        nslast = 0;
        me->hu = ZERO;
        me->nqu = 0;
        lf0 = me->lyh + me->nyh;
        me->f(me->ctx, me->n, *t, y, &RW(lf0));
        me->nfe = 1;
        dcopy(me->n, y, &RW(me->lyh));
        me->nq = 1;
        me->h = one;
        dewset(me->n, itol, rtol, atol, &RW(me->lyh), &RW(me->lewt));
        for (i = 1; i <= me->n; i++) {
            if (RW(i + me->lewt - 1) <= ZERO) ILLEGAL(21);
            RW(i + me->lewt - 1) = one / RW(i + me->lewt - 1);
        }
        if (h0 == ZERO) {
            dvhin(me, me->n, *t, &RW(me->lyh), &RW(lf0), tout, me->uround, &RW(me->lewt), itol,
                  atol, y, &RW(me->lacor), &h0, &niter, &ier);
            me->nfe = me->nfe + niter;
            if (ier != 0) ILLEGAL(22);
        }
        rh = fabs(h0) * me->hmxi;
        if (rh > one) h0 = h0 / rh;
        me->h = h0;
        dscal(me->n, h0, &RW(lf0));
        goto L200;
    } else {
        /* istate = 3, M:1383-1390 */
        me->jstart = -1;
        if (me->nq > me->maxord) dcopy(me->n, &RW(me->lwm), &RW(me->lsavf));
        if (me->miter > 0) RW(me->lwm) = sqrt(me->uround);
    }

L50: /* block d, M:1399-1483 */
    nslast = me->nst;
    me->kuth = 0;
    switch (itask) {
    case 2: goto L100;
    case 3:
        tp = me->tn - me->hu * (one + hun * me->uround);
        if ((tp - tout) * me->h > ZERO) ILLEGAL(23);
        if ((me->tn - tout) * me->h >= ZERO) {
            SET_Y_AND_T();
            CONTINUE_RETURN();
        }
        goto L100;
    case 4:
        tcrit = RW(1);
        if ((me->tn - tcrit) * me->h > ZERO) ILLEGAL(24);
        if ((tcrit - tout) * me->h < ZERO) ILLEGAL(25);
        if ((me->tn - tout) * me->h >= ZERO) {
            dvo_dvindy(me, tout, 0, &RW(me->lyh), me->nyh, y, &iflag);
            if (iflag != 0) ILLEGAL(27);
            *t = tout;
            CONTINUE_RETURN();
        }
        break;
    case 5:
        tcrit = RW(1);
        if ((me->tn - tcrit) * me->h > ZERO) ILLEGAL(24);
        break;
    default:
        if ((me->tn - tout) * me->h < ZERO) goto L100;
        dvo_dvindy(me, tout, 0, &RW(me->lyh), me->nyh, y, &iflag);
        if (iflag != 0) ILLEGAL(27);
        *t = tout;
        CONTINUE_RETURN();
    }
    /* itask = 4 or 5 only, M:1469-1483 */
    hmx = fabs(me->tn) + fabs(me->h);
    ihit = fabs(me->tn - tcrit) <= hun * me->uround * hmx;
    if (ihit) {
        SET_Y_AND_T();
        *t = tcrit;
        CONTINUE_RETURN();
    }
    tnext = me->tn + me->hnew * (one + four * me->uround);
    if ((tnext - tcrit) * me->h > ZERO) {
        me->h = (tcrit - me->tn) * (one - four * me->uround);
        me->kuth = 1;
    }

L100: /* block e, M:1496-1553 */
    if ((me->nst - nslast) >= me->mxstep) {
        me->last_msg = 201;
        *istate = -1;
        FINISH_RETURN();
    }
    dewset(me->n, itol, rtol, atol, &RW(me->lyh), &RW(me->lewt));
    for (i = 1; i <= me->n; i++) {
        if (RW(i + me->lewt - 1) <= ZERO) {
            me->last_msg = 202;
            *istate = -6;
            FINISH_RETURN();
        }
        RW(i + me->lewt - 1) = one / RW(i + me->lewt - 1);
    }
L200:
    tolsf = me->uround * dvnorm(me->n, &RW(me->lyh), &RW(me->lewt));
    if (tolsf <= one) {
        if ((me->tn + me->h) == me->tn) {
            me->nhnil = me->nhnil + 1;
            if (me->nhnil <= me->mxhnil) me->last_msg = 101;
        }
        dvstep(me, y, &RW(me->lyh), me->nyh, &RW(me->lyh), &RW(me->lewt), &RW(me->lsavf), y,
               &RW(me->lacor), &RW(me->lwm), &IW(me->liwm));
        if (me->kflag == -1) {
            me->last_msg = 204;
            *istate = -4;
            COMPUTE_IMXER();
            FINISH_RETURN();
        } else if (me->kflag == -2) {
            me->last_msg = 205;
            *istate = -5;
            COMPUTE_IMXER();
            FINISH_RETURN();
        } else {
            /* block f, M:1586-1623 */
            me->init = 1;
            me->kuth = 0;
            switch (itask) {
            case 2: break;
            case 3:
                if ((me->tn - tout) * me->h < ZERO) goto L100;
                break;
            case 4:
                if ((me->tn - tout) * me->h < ZERO) {
                    hmx = fabs(me->tn) + fabs(me->h);
                    ihit = fabs(me->tn - tcrit) <= hun * me->uround * hmx;
                    if (!ihit) {
                        tnext = me->tn + me->hnew * (one + four * me->uround);
                        if ((tnext - tcrit) * me->h > ZERO) {
                            me->h = (tcrit - me->tn) * (one - four * me->uround);
                            me->kuth = 1;
                        }
                        goto L100;
                    }
                } else {
                    dvo_dvindy(me, tout, 0, &RW(me->lyh), me->nyh, y, &iflag);
                    *t = tout;
                    CONTINUE_RETURN();
                }
                break;
            case 5:
                hmx = fabs(me->tn) + fabs(me->h);
                ihit = fabs(me->tn - tcrit) <= hun * me->uround * hmx;
                break;
            default:
                if ((me->tn - tout) * me->h < ZERO) goto L100;
                dvo_dvindy(me, tout, 0, &RW(me->lyh), me->nyh, y, &iflag);
                *t = tout;
                CONTINUE_RETURN();
            }
        }
    } else {
        /* M:1625-1645 */
        tolsf = tolsf * two;
        if (me->nst == 0) {
            me->last_msg = 26;
            RW(14) = tolsf;
            *istate = -3;
            return;
        } else {
            me->last_msg = 203;
            RW(14) = tolsf;
            *istate = -2;
            FINISH_RETURN();
        }
    }

    /* block g, M:1648-1652 */
    SET_Y_AND_T();
    if (itask == 4 || itask == 5) {
        if (ihit) *t = tcrit;
    }
    CONTINUE_RETURN();
#undef RW
#undef IW
}
