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
    double h0, hmax, hmin;     /* rwork(5:7), used when iopt = 1 */
    double tcrit;              /* rwork(1) */
    int pow_mode;              /* 0 libm pow, 1 portable pow (strict GPU comparison) */
} dvo_batch_cfg;

int dvo_problem_info(int problem, int *neq, int *npar);
void dvo_problem_f(int problem, const double *par, double t, const double *y, double *ydot);
void dvo_work_sizes(int neq, int mf, int ml, int mu, int maxord_in, int *lrw, int *liw);

/* returns the number of systems that ended with istate < 0 (or -1 on bad problem).
 * rstats[.][.][4] = rwork(11:14); istats[.][.][12] = iwork(11:22). */
int dvo_batch_solve(const dvo_batch_cfg *cfg, int nsys, const double *params, const double *y0,
                    double t0, int nout, const double *touts, const double *rtol,
                    const double *atol, double *y_out, double *t_out, int *istate_out,
                    double *rstats, int *istats, int nthreads);

int dvo_batch_solve2(const dvo_batch_cfg *cfg, const dvo_batch_cfg *cfg2, int nsplit, int nsys,
                     const double *params, const double *y0, double t0, int nout,
                     const double *touts, const double *rtol, const double *atol,
                     const double *rtol2, const double *atol2, double *y_out, double *t_out,
                     int *istate_out, double *rstats, int *istats, int nthreads);

/* threads the last dvo_batch_solve / dvo_batch_solve2 call actually ran on */
int dvo_last_threads(void);

int dvo_single_dense(const dvo_batch_cfg *cfg, const double *par, const double *y0, double t0,
                     double tout, const double *rtol, const double *atol, int k, int nq,
                     const double *tqs, double *dky_out, int *iflag_out, double *tcur,
                     double *hu);
#ifdef __cplusplus
}
#endif
#endif
