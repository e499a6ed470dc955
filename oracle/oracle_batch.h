/*
 * oracle_batch.h -- batch driver around the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * One independent dvo_solver per system; OpenMP over systems.  Host layout is
 * array-of-systems: params[nsys][npar], y0[nsys][neq], outputs [nsys][nout][...].
 */
#ifndef ORACLE_BATCH_H
#define ORACLE_BATCH_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dvo_batch_cfg {
    int problem;               /* 0 robertson, 1 vdp, 2 banded advection, 3 mech32 */
    int mf, itol, itask, iopt; /* as in dvode, M:604 */
    int ml, mu;                /* iwork(1:2) */
    int maxord, mxstep, mxhnil;/* iwork(5:7), used when iopt = 1 */
    real h0, hmax, hmin;     /* rwork(5:7), used when iopt = 1 */
    real tcrit;              /* rwork(1) */
    int pow_mode;              /* 0 libm pow, 1 portable pow (strict GPU comparison) */
} dvo_batch_cfg;

int dvo_problem_info(int problem, int *neq, int *npar);
void dvo_problem_f(int problem, const real *par, real t, const real *y, real *ydot);
void dvo_work_sizes(int neq, int mf, int ml, int mu, int maxord_in, int *lrw, int *liw);

/* returns the number of systems that ended with istate < 0 (or -1 on bad problem).
 * rstats[.][.][4] = rwork(11:14); istats[.][.][12] = iwork(11:22). */
int dvo_batch_solve(const dvo_batch_cfg *cfg, int nsys, const real *params, const real *y0,
                    real t0, int nout, const real *touts, const real *rtol,
                    const real *atol, real *y_out, real *t_out, int *istate_out,
                    real *rstats, int *istats, int nthreads);

/* neq of the calls from the istate = 3 call on in dvo_batch_solve2 (0 = unchanged; M:1127-1135) */
void dvo_set_neq2(int neq2);

int dvo_batch_solve2(const dvo_batch_cfg *cfg, const dvo_batch_cfg *cfg2, int nsplit, int nsys,
                     const real *params, const real *y0, real t0, int nout,
                     const real *touts, const real *rtol, const real *atol,
                     const real *rtol2, const real *atol2, real *y_out, real *t_out,
                     int *istate_out, real *rstats, int *istats, int nthreads);

/* threads the last dvo_batch_solve / dvo_batch_solve2 call actually ran on */
int dvo_last_threads(void);

int dvo_single_dense(const dvo_batch_cfg *cfg, const real *par, const real *y0, real t0,
                     real tout, const real *rtol, const real *atol, int k, int nq,
                     const real *tqs, real *dky_out, int *iflag_out, real *tcur,
                     real *hu);
#ifdef __cplusplus
}
#endif
#endif
