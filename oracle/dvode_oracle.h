/*
 * dvode_oracle.h -- CPU restatement of jacobwilliams/dvode (REAL64) in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the B200 batched
 * VODE path.  Nothing in the product path (dvode_b200/, include/) may call,
 * link or import it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity pin: the restatement is pinned against the reference's own golden
 * transcript test/dvdemo.txt (all 24 runs: printed digits, NQ, H, RWORK/IWORK
 * sizes and all counters) and against test/example.f90's trailing comment
 * (Robertson).  See tests/test_oracle_golden.py.  The reference itself cannot
 * be compiled here (no Fortran compiler in the image).
 *
 * Citations below are file:line into /root/reference (M = src/dvode_module.F90,
 * LP = src/dvode_linpack_module.F90, BL = src/dvode_blas_module.F90).
 */
#ifndef DVODE_ORACLE_H
#define DVODE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* the working precision: the reference's `wp` (K:29-31).  -DDVO_REAL32 builds liboracle32.so, the REAL32
 * twin that libbvode32.so is compared with.  DVO_R(x) types a literal like the reference's `x_wp`. */
#ifdef DVO_REAL32
typedef float real;
#else
typedef double real;
#endif
#define DVO_R(x) ((real)(x))

/* user callbacks, M:316-334.  ctx replaces the Fortran `me` host association. */
typedef void (*dvo_f_fn)(void *ctx, int neq, real t, const real *y, real *ydot);
typedef void (*dvo_jac_fn)(void *ctx, int neq, real t, real *y, int ml, int mu,
                           real *pd, int nrowpd);
/* user-replaceable error-weight and norm routines (M:431-503) */
typedef void (*dvo_ewt_fn)(void *ctx, int n, int itol, const real *rtol, const real *atol,
                           const real *ycur, real *ewt);
typedef real (*dvo_norm_fn)(void *ctx, int n, const real *v, const real *w);

/* dvode_data_t (M:40-153) + dvode_t (M:251-270) */
typedef struct dvo_solver {
    /* ex-common dvod01 */
    real acnrm, ccmxj, conp, crate, drc, el[14], eta, etamax, h, hmin, hmxi, hnew,
           hscal, prl1, rc, rl1, tau[14], tq[6], tn, uround;
    int icf, init, ipup, jcur, jstart, jsv, kflag, kuth, l, lmax, lyh, lewt, lacor,
        lsavf, lwm, liwm, locjs, maxord, meth, miter, msbj, mxhnil, mxstep, n, newh,
        newq, nhnil, nq, nqnyh, nqwait, nslj, nslp, nyh;
    /* ex-common dvod02 */
    real hu;
    int ncfn, netf, nfe, nje, nlu, nni, nqu, nst;
    /* dvode_t */
    real etaq, etaqm1;
    dvo_f_fn f;
    dvo_jac_fn jac;
    dvo_ewt_fn dewset;   /* NULL = dewset_default, M:3238 */
    dvo_norm_fn dvnorm;  /* NULL = dvnorm_default, M:3295 */
    void *ctx;
    int last_msg;     /* number of the last xerrwd message (M:3351), 0 if none */
    int pow_mode;     /* 0 = libm pow (reference semantics); 1 = portable pow (dvo_ppow) */
} dvo_solver;

/* test instrumentation: singular Newton matrices seen (dvjac's ierpj != 0) since the last reset */
void dvo_singular_events_add(void);
int dvo_singular_events(int reset);

/* initialize, M:377-505; the optional dewset / dvnorm procedure arguments (M:431-503) are set
 * with dvo_set_user_norms (NULL keeps the default) */
void dvo_initialize(dvo_solver *s, dvo_f_fn f, dvo_jac_fn jac, void *ctx);
void dvo_set_user_norms(dvo_solver *s, dvo_ewt_fn dewset, dvo_norm_fn dvnorm);

/* dvode / solve, M:604-1725.  rwork and iwork are 1-based in the reference;
 * here the caller passes plain C arrays of length lrw / liw and slot k of the
 * reference is element [k-1]. */
void dvo_solve(dvo_solver *s, int neq, real *y, real *t, real tout, int itol,
               const real *rtol, const real *atol, int itask, int *istate, int iopt,
               real *rwork, int lrw, int *iwork, int liw, int mf);

/* dvindy, M:1878-1959.  yh = &rwork[20] (rwork(21)), ldyh = nyh. */
void dvo_dvindy(dvo_solver *s, real t, int k, const real *yh, int ldyh, real *dky,
                int *iflag);

/* dvsrco, M:3198-3216 */
void dvo_dvsrco(dvo_solver *s, dvo_solver *sav, int job);

/* LINPACK, LP:46-522 (job = 0 branches), column-major, 1-based pivots. */
void dvo_dgefa(real *a, int lda, int n, int *ipvt, int *info);
void dvo_dgesl(const real *a, int lda, int n, const int *ipvt, real *b);
void dvo_dgbfa(real *abd, int lda, int n, int ml, int mu, int *ipvt, int *info);
void dvo_dgbsl(const real *abd, int lda, int n, int ml, int mu, const int *ipvt, real *b);

/* x**n for integer n the way gfortran lowers it (__powidf2: binary powering). */
real dvo_powi(real x, int n);
/* portable, deterministic x**y (x>0) built from + - * / only; mirrored on the device
 * by the strict build so that GPU and oracle can be compared bit for bit. */
/* REALIFY-OFF */
double dvo_ppow(double x, double y);
/* REALIFY-ON */

#ifdef __cplusplus
}
#endif
#endif
