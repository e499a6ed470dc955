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
 *                        system 0's row (from every row with BVODE_FLAG_PER_SYSTEM); optional
 *                        OUTPUTS rwork(11:14) are written per system.
 *                        The reference's rwork(21:) (Nordsieck array, Newton matrix, ...) is
 *                        the device-resident state owned by the handle.
 *   iwork  [nsys][liw]   liw >= 30: same for iwork(1:30): inputs iwork(1,2,5,6,7) from row 0,
 *                        outputs iwork(11:22) per system.
 *   rtol, atol           scalars or arrays of length neq according to itol (M:663-668),
 *                        shared by all systems of the batch; with BVODE_FLAG_PER_SYSTEM one such
 *                        scalar / array PER SYSTEM: rtol[nsys][1 or neq], atol[nsys][1 or neq].
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

/* The working precision of the library: the reference's `wp` (src/dvode_kinds_module.F90:29-31).  libbvode.so
 * is the REAL64 build; libbvode32.so (compiled with -DBVODE_REAL32) is the REAL32 build of the same sources and
 * exports the same entry points with every `bvode_real` a float (strict arithmetic, thread-per-system problems).
 * An application selects one at compile time, like the reference's -DREAL32. */
#ifdef BVODE_REAL32
typedef float bvode_real;
#else
typedef double bvode_real;
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
    BVODE_FLAG_COMPACT = 2,/* thread-per-system kernels, itask = 1: launch only as many threads as are
                            * resident at once and give a thread whose system is finished the next
                            * unclaimed system of the batch (compaction of finished systems).  Results
                            * are identical; pays off when the systems differ widely in cost. */
    BVODE_FLAG_PER_SYSTEM = 4
                           /* every system is its own solver instance as far as tolerances and optional
                            * inputs go, as in the reference (M:729-751, 928-965): rtol[nsys][1 or neq],
                            * atol[nsys][1 or neq] (by itol), and the optional inputs rwork(1,5,6,7) /
                            * iwork(5,6,7) are read from EVERY system's row.  A system whose row is illegal
                            * gets istate = -3 and is left alone (M:1120-1303).  mf, ml / mu, itol, itask,
                            * iopt stay batch-uniform.  Runs the full-driver kernels. */
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
BVODE_API int bvode_set_params(bvode_solver *s, const bvode_real *params);

/* ---- solve == dvode (M:604-605), for the whole batch.  Same argument meaning as the
 * reference.  tout: one value for the batch (tout_per_system = 0) or tout[nsys] (= 1).
 *   itask  1-5 (M:811-831); 4 and 5 read tcrit = rwork(1) of row 0.
 *   istate 1 (first call), 2 (continue), 3 (continue with changed itol / rtol / atol / itask /
 *          optional inputs incl. a lower maxord; must then be 3 for every system; mf may change
 *          -- miter, jsv and meth (Adams <-> BDF): the state moves into the new kernel's layout and
 *          the Jacobian is re-evaluated -- ml / mu of a banded flag may change, and neq may be
 *          reduced (M:1127-1135: y then has neq entries per system, the dropped components stay
 *          what y held for them; never raised again).
 *   mf     jsv*(10*meth + miter), M:999-1046: meth 1 (Adams) or 2 (BDF); miter 0 (functional
 *          iteration), 1 / 2 (dense analytic / finite-difference J; thread-per-system problems,
 *          miter 2 also for the dense warp problem), 3 (diagonal approximation), 4 / 5 (banded
 *          analytic / FD J, band problems); mf < 0 = no saved Jacobian copy. */
BVODE_API int bvode_solve(bvode_solver *s, int neq, bvode_real *y, bvode_real *t, const bvode_real *tout,
                int tout_per_system, int itol, const bvode_real *rtol, const bvode_real *atol,
                int itask, int *istate, int iopt, bvode_real *rwork, int lrw, int *iwork, int liw,
                int mf);

/* The reference's usage loop (test/example.f90:56-66: nout calls with increasing tout) as ONE
 * call and one kernel launch (itask 1 or 4; itask 2, 3, 5 return at mesh points and are used one
 * call at a time through bvode_solve): equivalent to
 *     for k in 0..nout-1: bvode_solve(..., tout = touts[k], ...); y_out[k] = y
 * including the per-call mxstep budget.  y holds y0 on entry (istate = 1) and the last output
 * on return; y_out[nout][nsys][neq] receives every output point (may be NULL). */
BVODE_API int bvode_solve_schedule(bvode_solver *s, int neq, bvode_real *y, bvode_real *t, int nout,
                         const bvode_real *touts, int itol, const bvode_real *rtol, const bvode_real *atol,
                         int itask, int *istate, int iopt, bvode_real *rwork, int lrw, int *iwork,
                         int liw, int mf, bvode_real *y_out);

/* ---- dvindy (M:1878-1959): k-th derivative of the interpolating polynomial at t.
 * t: one value (t_per_system = 0) or t[nsys]; dky[nsys][neq]; iflag[nsys] = 0 / -1 / -2. */
BVODE_API int bvode_dvindy(bvode_solver *s, const bvode_real *t, int t_per_system, int k, bvode_real *dky,
                 int *iflag);

/* Optional outputs without a solve call: rout[nsys][4] = rwork(11:14) = hu, hcur, tcur, tolsf;
 * iout[nsys][12] = iwork(11:22) = nst, nfe, nje, nqu, nqcur, imxer, lenrw, leniw, nlu, nni,
 * ncfn, netf (M:764-783, 978-997).  Either pointer may be NULL. */
BVODE_API int bvode_get_stats(bvode_solver *s, bvode_real *rout, int *iout);

/* dvsrco (M:3198-3216): save (job = 1) / restore (job = 2) the whole batch state to / from a
 * host buffer of bvode_state_bytes() bytes: checkpoint / resume.  The blob carries the persistent
 * state, the optional outputs, the handle's istate / t / y of the last host-buffer call and the
 * optional inputs that persist between calls (maxord, mxstep).  bvode_dvsrco_n checks the buffer
 * length `nbytes` against what it reads or writes; bvode_dvsrco trusts the caller (the reference's
 * argument list has no length). */
BVODE_API long long bvode_state_bytes(const bvode_solver *s);
BVODE_API int bvode_dvsrco_n(bvode_solver *s, void *sav, long long nbytes, int job);
BVODE_API int bvode_dvsrco(bvode_solver *s, void *sav, int job);
/* Drop the integration state of the handle (the next call must be a first call, istate = 1): needed
 * before a device-resident solve with another mf / ml / mu on the same handle. */
BVODE_API int bvode_reset(bvode_solver *s);

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
BVODE_API int bvode_set_params_device(bvode_solver *s, const bvode_real *params_device);
BVODE_API int bvode_solve_schedule_device(bvode_solver *s, int neq, bvode_real *y_device, bvode_real *t_device,
                                int nout, const bvode_real *touts, int itol, const bvode_real *rtol,
                                const bvode_real *atol, int itask, int *istate_device, int iopt,
                                const bvode_real *rwork_opts, const int *iwork_opts, int mf,
                                bvode_real *y_out_device, void *stream);
/* One call (M:604-605) with inputs in HBM and a tout PER SYSTEM, tout_device[nsys] (the device-pointer form of
 * bvode_solve with tout_per_system = 1); optional outputs through bvode_get_stats. */
BVODE_API int bvode_solve_device(bvode_solver *s, int neq, bvode_real *y_device, bvode_real *t_device,
                       const bvode_real *tout_device, int itol, const bvode_real *rtol, const bvode_real *atol,
                       int itask, int *istate_device, int iopt, const bvode_real *rwork_opts,
                       const int *iwork_opts, int mf, void *stream);
/* number of kernels the last solve call launched (bench.py's gpu_launches) */
BVODE_API int bvode_last_launch_count(const bvode_solver *s);

/* ---- multi-GPU: the batch shards by system index, each GPU integrates its block on its
 * own, and the only exchange is a final gather of solutions and statistics to one GPU over NVLink (NCCL,
 * loaded at run time).  No reference analogue: dvode is a serial code, one solver instance per thread.
 *
 * bvode_shard_range: rank r of `world` owns systems [first, last); block sizes differ by at most one.
 *
 * One process per GPU: rank 0 calls bvode_comm_unique_id, the 128 bytes travel to the other ranks by
 * whatever the application has (MPI_Bcast, a file, torch.distributed), every rank calls
 * bvode_comm_create; or bvode_comm_adopt wraps an ncclComm_t the application made itself.
 *
 * bvode_gather_device: after a solve, gather the final y / t / istate and the optional outputs of every
 * rank's handle into arrays on the root's GPU, in system order and host layout: y_all[nsys_total][neq],
 * t_all[nsys_total], istate_all[nsys_total], rout_all[nsys_total][4], iout_all[nsys_total][12] (device
 * pointers on the root; on the other ranks any non-NULL value means "this array is wanted"; every rank must
 * pass the same set).  Enqueued on `stream` (a cudaStream_t as void*, 0 = the default stream, as in
 * bvode_solve_schedule_device: pass the stream the solve was enqueued on); returns without waiting.  bvode_gather: the same with HOST arrays on the root, and it waits. */
typedef struct bvode_comm bvode_comm;
BVODE_API void bvode_shard_range(long long nsys_total, int rank, int world, long long *first, long long *last);
BVODE_API int bvode_nccl_version(void);                       /* 0: NCCL could not be loaded */
BVODE_API int bvode_comm_unique_id(char id[128]);
BVODE_API int bvode_comm_create(bvode_comm **out, const char id[128], int world, int rank, int device);
BVODE_API int bvode_comm_adopt(bvode_comm **out, void *nccl_comm, int world, int rank, int device);
BVODE_API void bvode_comm_destroy(bvode_comm *c);
BVODE_API int bvode_gather_device(bvode_solver *s, bvode_comm *c, int root, long long nsys_total,
                                  bvode_real *y_all_device, bvode_real *t_all_device, int *istate_all_device,
                                  bvode_real *rout_all_device, int *iout_all_device, void *stream);
BVODE_API int bvode_gather(bvode_solver *s, bvode_comm *c, int root, long long nsys_total, bvode_real *y_all,
                           bvode_real *t_all, int *istate_all, bvode_real *rout_all, int *iout_all);

/* One process, several GPUs -- what a Fortran / C program without MPI calls to use the whole box.  The
 * batch of nsys_total systems is sharded over `ndev` devices (devices[ndev], NULL = 0..ndev-1); the solve
 * calls take arrays of nsys_total rows with exactly the arguments of bvode_solve / bvode_solve_schedule
 * and run every block on its own GPU at the same time (each GPU copies its rows over its own PCIe link).
 * bvode_multi_gather_device puts the final results of all blocks into one GPU's memory (NCCL over NVLink)
 * for callers that go on working on the device.  bvode_multi_handle gives the per-device handle
 * (bvode_dvindy, bvode_get_stats, ... on one block). */
typedef struct bvode_multi bvode_multi;
BVODE_API int bvode_multi_create(bvode_multi **out, int problem, long long nsys_total, int ndev,
                                 const int *devices, unsigned flags);
BVODE_API void bvode_multi_destroy(bvode_multi *m);
BVODE_API int bvode_multi_device_count(const bvode_multi *m);
BVODE_API bvode_solver *bvode_multi_handle(bvode_multi *m, int slot);
BVODE_API int bvode_multi_set_params(bvode_multi *m, const bvode_real *params);
BVODE_API int bvode_multi_solve(bvode_multi *m, int neq, bvode_real *y, bvode_real *t, const bvode_real *tout,
                                int tout_per_system, int itol, const bvode_real *rtol, const bvode_real *atol,
                                int itask, int *istate, int iopt, bvode_real *rwork, int lrw, int *iwork,
                                int liw, int mf);
BVODE_API int bvode_multi_solve_schedule(bvode_multi *m, int neq, bvode_real *y, bvode_real *t, int nout,
                                         const bvode_real *touts, int itol, const bvode_real *rtol,
                                         const bvode_real *atol, int itask, int *istate, int iopt,
                                         bvode_real *rwork, int lrw, int *iwork, int liw, int mf, bvode_real *y_out);
BVODE_API int bvode_multi_gather_device(bvode_multi *m, int root_slot, bvode_real *y_all_device,
                                        bvode_real *t_all_device, int *istate_all_device,
                                        bvode_real *rout_all_device, int *iout_all_device);

/* ---- the device LINPACK routines in isolation: known-answer entry for tests and diagnostics.  The reference
 * has no test of dgefa / dgesl / dgbfa / dgbsl (LP:46-522) on their own; this runs exactly the factor / solve code
 * of one kernel family on nmat matrices and one right-hand side each (all HOST pointers).
 *   kind 0  one matrix per thread, registers (n = 2, 3, 4): a[nmat][n*n] column-major; afac / ipvt come back in
 *           LINPACK's layout (ipvt 1-based)
 *   kind 1  one matrix per warp, row i on lane i, implicit pivoting (n = 5, 25, 32); kind 2: two rows per lane
 *           (n = 48, 64): afac[nmat][n][n] is the row image (row r never moves), ipvt[k] = the row that is the
 *           pivot of step k (0-based)
 *   kind 3  LINPACK band storage (n = 25): a[nmat][n][2*ml+mu+1] (abd(i, j) at a[j*lda + i]); afac / ipvt in
 *           LINPACK's layout
 * x[nmat][n] = the solution (zeros where info != 0); info[nmat] as dgefa / dgbfa (LP:87-88, 119).
 * flags: BVODE_FLAG_STRICT selects the bit-exact build. */
BVODE_API int bvode_lu_test(int kind, int n, int ml, int mu, int nmat, const bvode_real *a, const bvode_real *b,
                            bvode_real *afac, int *ipvt, bvode_real *x, int *info, int device, unsigned flags);

/* ---- DFMA peak microbenchmark (roofline denominator, FP64 vector pipe).  Runs `iters`
 * dependent-chain FMAs x 8 chains per thread on every SM and returns TFLOP/s. */
BVODE_API int bvode_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
