/*
 * bvode.h -- C ABI of the B200-native batched VODE integrator.
 *
 * Drop-in boundary for ONE path of jacobwilliams/dvode: "integrate a batch of
 * independent ODE systems with VODE's variable-order, variable-step BDF / Adams methods".
 * Every entry point names the reference interface it replaces (file:line into
 * /root/reference; M = src/dvode_module.F90).  Plain C types only; all pointers
 * are HOST pointers unless the name says `_device`.
 *
 * Conventions that differ from the scalar reference because there are nsys systems:
 *   y      [nsys][neq]   system-major (each system's y is contiguous, as in one dvode call)
 *   t      [nsys]        per-system independent variable
 *   istate [nsys]        per-system state flag, same integer codes as M:832-908
 *   rwork  [nsys][lrw]   lrw >= 20: per-system copy of the reference's rwork(1:20) header.
 *                        Optional INPUTS rwork(1,5,6,7) are batch-uniform and are read from
 *                        system 0's row; optional OUTPUTS rwork(11:14) are written per system.
 *                        The reference's rwork(21:) (Nordsieck array, Newton matrix, ...) is
 *                        the device-resident state owned by the handle.
 *   iwork  [nsys][liw]   liw >= 30: same for iwork(1:30): inputs iwork(1,2,5,6,7) from row 0,
 *                        outputs iwork(11:22) per system.
 *   rtol, atol           scalars or arrays of length neq according to itol (M:663-668),
 *                        shared by all systems of the batch.
 * The user callbacks f / jac (M:316-334) cannot be host function pointers: they are
 * __device__ functors compiled into the library and selected by `problem` id.
 *
 * Return value of every function: 0 on success, a negative BVODE_E* code on API misuse
 * (the library never aborts, M:1083-1089's `stop` is replaced by BVODE_EISTATE).
 * Per-system integration outcomes are reported in istate[], never in the return value.
 */
#ifndef BVODE_H
#define BVODE_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BVODE_API __attribute__((visibility("default")))
#else
#define BVODE_API
#endif

typedef struct bvode_solver bvode_solver;

enum {
    BVODE_OK = 0,
    BVODE_EINVAL = -1,    /* illegal argument (the reference's istate = -3 input checks, M:1120-1303) */
    BVODE_ENOTIMPL = -2,  /* legal in the reference but not built (e.g. istate = 3 with another mf) */
    BVODE_ECUDA = -3,     /* CUDA runtime error; bvode_last_error() has the text */
    BVODE_ENOMEM = -4,
    BVODE_EISTATE = -5    /* negative istate on input: M:1083-1089 aborts the run here */
};

/* flags for bvode_create */
enum {
    BVODE_FLAG_STRICT = 1, /* run the -fmad=false / portable-pow build (bit-for-bit parity runs) */
    BVODE_FLAG_COMPACT = 2 /* thread-per-system kernels, itask = 1: launch only as many threads as are
                            * resident at once and give a thread whose system is finished the next
                            * unclaimed system of the batch (compaction of finished systems).  Results
                            * are identical; pays off when the systems differ widely in cost. */
};

/* ---- problem registry: stands in for the procedure arguments of initialize (M:377-505) ---- */
BVODE_API int bvode_problem_count(void);
BVODE_API int bvode_problem_lookup(const char *name);                        /* id >= 0, or -1 */
BVODE_API int bvode_problem_info(int problem, int *neq, int *npar, const char **name);

/* Register a user problem at run time: the f / jac (and optional dewset / dvnorm) procedure arguments of
 * initialize (M:377-505) as a __device__ functor compiled into a library of its own with
 * dvode_b200/csrc/bvode_user_problem.cuh (examples/decay_chain.cu).  entry_fast /
 * entry_strict are what that library's bvode_user_entry_fast() / bvode_user_entry_strict() return
 * (entry_strict may be NULL: BVODE_FLAG_STRICT handles are then refused), abi its bvode_user_abi_fast().
 * Returns the new problem id (>= bvode_problem_count() before the call) or a negative BVODE_E* code. */
BVODE_API int bvode_register_problem(const void *entry_fast, const void *entry_strict, int abi);
BVODE_API int bvode_plugin_abi(void);

/* ---- lifetime: replaces `type(dvode_t) :: solver` + `call solver%initialize(f, jac)` (M:377) ----
 * One handle owns the device-resident state of nsys systems on one GPU. */
BVODE_API int bvode_create(bvode_solver **out, int problem, int nsys, int device, unsigned flags);
BVODE_API void bvode_destroy(bvode_solver *s);
BVODE_API const char *bvode_last_error(void);

/* Per-system parameters read by f / jac: params[nsys][npar] (what the Fortran callbacks
 * would reach by host association, e.g. Robertson's three rate constants). */
BVODE_API int bvode_set_params(bvode_solver *s, const double *params);

/* ---- solve == dvode (M:604-605), for the whole batch.  Same argument meaning as the
 * reference.  tout: one value for the batch (tout_per_system = 0) or tout[nsys] (= 1).
 *   itask  1-5 (M:811-831); 4 and 5 read tcrit = rwork(1) of row 0.
 *   istate 1 (first call), 2 (continue), 3 (continue with changed itol / rtol / atol / itask /
 *          optional inputs incl. a lower maxord; must then be 3 for every system; mf may change
 *          within a method family -- miter / jsv, the state moves into the new kernel's layout and
 *          the Jacobian is re-evaluated -- but not meth, ml / mu of a banded flag, or neq).
 *   mf     jsv*(10*meth + miter), M:999-1046: meth 1 (Adams) or 2 (BDF); miter 0 (functional
 *          iteration), 1 / 2 (dense analytic / finite-difference J; thread-per-system problems,
 *          miter 2 also for the dense warp problem), 3 (diagonal approximation), 4 / 5 (banded
 *          analytic / FD J, band problems); mf < 0 = no saved Jacobian copy. */
BVODE_API int bvode_solve(bvode_solver *s, int neq, double *y, double *t, const double *tout,
                int tout_per_system, int itol, const double *rtol, const double *atol,
                int itask, int *istate, int iopt, double *rwork, int lrw, int *iwork, int liw,
                int mf);

/* The reference's usage loop (test/example.f90:56-66: nout calls with increasing tout) as ONE
 * call and one kernel launch (itask 1 or 4; itask 2, 3, 5 return at mesh points and are used one
 * call at a time through bvode_solve): equivalent to
 *     for k in 0..nout-1: bvode_solve(..., tout = touts[k], ...); y_out[k] = y
 * including the per-call mxstep budget.  y holds y0 on entry (istate = 1) and the last output
 * on return; y_out[nout][nsys][neq] receives every output point (may be NULL). */
BVODE_API int bvode_solve_schedule(bvode_solver *s, int neq, double *y, double *t, int nout,
                         const double *touts, int itol, const double *rtol, const double *atol,
                         int itask, int *istate, int iopt, double *rwork, int lrw, int *iwork,
                         int liw, int mf, double *y_out);

/* ---- dvindy (M:1878-1959): k-th derivative of the interpolating polynomial at t.
 * t: one value (t_per_system = 0) or t[nsys]; dky[nsys][neq]; iflag[nsys] = 0 / -1 / -2. */
BVODE_API int bvode_dvindy(bvode_solver *s, const double *t, int t_per_system, int k, double *dky,
                 int *iflag);

/* Optional outputs without a solve call: rout[nsys][4] = rwork(11:14) = hu, hcur, tcur, tolsf;
 * iout[nsys][12] = iwork(11:22) = nst, nfe, nje, nqu, nqcur, imxer, lenrw, leniw, nlu, nni,
 * ncfn, netf (M:764-783, 978-997).  Either pointer may be NULL. */
BVODE_API int bvode_get_stats(bvode_solver *s, double *rout, int *iout);

/* dvsrco (M:3198-3216): save (job = 1) / restore (job = 2) the whole batch state to / from a
 * host buffer of bvode_state_bytes() bytes: checkpoint / resume. */
BVODE_API long long bvode_state_bytes(const bvode_solver *s);
BVODE_API int bvode_dvsrco(bvode_solver *s, void *sav, int job);

/* ---- xerrwd (M:3351-3394) as a host-side message stream.  The kernels never print and never
 * stop; the per-system outcome is istate.  bvode_istate_summary compacts the batch: counts[0] =
 * systems without error (istate 1 / 2), counts[k] = systems with istate = -k (k = 1..6),
 * counts[7] = anything else.  bvode_format_message writes the reference's message text for
 * system `sys` (its "in above message, r1 = ..." lines filled from tcur / hcur / tolsf) into buf
 * and returns its length (0 when that system ended without error). */
BVODE_API int bvode_istate_summary(bvode_solver *s, int counts[8]);
BVODE_API int bvode_format_message(bvode_solver *s, int sys, char *buf, int buflen);

/* ---- device-resident variants (inputs already in HBM; used for the kernel-only metric).
 * Layout on the device is structure-of-arrays: y_device[neq][nsys], params_device[npar][nsys],
 * y_out_device[nout][neq][nsys].  stream is a cudaStream_t passed as void* (0 = default). */
BVODE_API int bvode_set_params_device(bvode_solver *s, const double *params_device);
BVODE_API int bvode_solve_schedule_device(bvode_solver *s, int neq, double *y_device, double *t_device,
                                int nout, const double *touts, int itol, const double *rtol,
                                const double *atol, int itask, int *istate_device, int iopt,
                                const double *rwork_opts, const int *iwork_opts, int mf,
                                double *y_out_device, void *stream);
/* One call (M:604-605) with inputs in HBM and a tout PER SYSTEM, tout_device[nsys] (the device-pointer form of
 * bvode_solve with tout_per_system = 1); optional outputs through bvode_get_stats. */
BVODE_API int bvode_solve_device(bvode_solver *s, int neq, double *y_device, double *t_device,
                       const double *tout_device, int itol, const double *rtol, const double *atol,
                       int itask, int *istate_device, int iopt, const double *rwork_opts,
                       const int *iwork_opts, int mf, void *stream);
/* number of kernels the last solve call launched (bench.py's gpu_launches) */
BVODE_API int bvode_last_launch_count(const bvode_solver *s);

/* ---- DFMA peak microbenchmark (roofline denominator, FP64 vector pipe).  Runs `iters`
 * dependent-chain FMAs x 8 chains per thread on every SM and returns TFLOP/s. */
BVODE_API int bvode_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
