// bvode_internal.h -- plain structs shared by the two kernel builds (fast / strict) and the
// C-ABI layer.  Nothing here is namespaced per build.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "bvode_real.h"

#define BVODE_MAX_NOUT 16

// Batch-uniform solver options: optional inputs after defaulting (M:729-751, 1174-1238).
struct BvOpts {
    real h0, hmax_inv, hmin;   // rwork(5), 1/rwork(6) (0 = no limit), rwork(7)
    int maxord, mxstep, mxhnil;  // iwork(5:7)
    int itol;
    int ml, mu;                  // iwork(1:2) (band problems)
    int itask;                   // 1..5 (M:811-831)
    int jreset;                  // istate = 3 with a changed maxord and a saved Jacobian: the saved
                                 // copy is not trusted (the reference's work-array segments move,
                                 // M:1246-1268, and its copy is lost) -> re-evaluate at the next dvjac
    real tcrit;                // rwork(1), itask 4 / 5
    int lmaxchg;                 // istate = 3 with another lmax = maxord + 1 (maxord or the method family changed): the
                                 // reference parks the acor copy in yh(:, lmax) (M:2259, 2298), so a pending order
                                 // increase now finds the Nordsieck column yh(:, lmax_new) there instead
    int nact;                    // equations actually integrated (0 = all neq): an istate = 3 call may pass a smaller
                                 // neq (M:1127-1135); the dropped components stay frozen in the state
    int stash;                   // istate = 3 from Adams to BDF: the Nordsieck array shrinks from 13 to 6 columns; the
                                 // old column maxord + 2, which the order reduction of M:1385, 2044-2079 reads, travels
                                 // in the acor slot of the converted state
};

// Arguments of one solve launch.  All pointers are device pointers, SoA layout.
struct KArgs {
    int nsys, nout;                // nsys: systems of the batch = stride of every SoA array
    int sys0, count;               // this launch integrates systems [sys0, sys0 + count)
    real touts[BVODE_MAX_NOUT];  // output points of this launch (nout >= 1)
    const real *tout_ps;         // per-system tout[nsys] (only with nout == 1) or nullptr
    real *dstate;                // [nd][nsys] persistent real state  (the reference's rwork(21:))
    int *istate_store;             // [ni][nsys] persistent integer state (iwork(31:) + dvode_data_t)
    const real *params;          // [npar][nsys]
    real *y;                     // [neq][nsys] in/out
    real *t;                     // [nsys] in/out
    int *istate;                   // [nsys] in/out
    const real *rtol, *atol;     // [neq] each, already expanded according to itol
    real rtolv[8], atolv[8];     // the same, by value, for thread-per-system kernels with neq <= 8
                                   // (larger thread problems read rtol / atol above)
    // Per-system tolerances and optional inputs (BVODE_FLAG_PER_SYSTEM; the reference takes them per
    // solver instance, M:729-751, 928-965).  nullptr = batch-uniform (rtol / atol / o above).  Read by
    // the full-driver kernels only (general = 1).
    const real *tol_ps;          // [2*neq][nsys]: rtol then atol, expanded according to itol
    const real *optd_ps;         // [4][nsys]: h0, 1/hmax (0 = none), hmin, tcrit
    const int *opti_ps;            // [3][nsys]: maxord, mxstep, jreset + 2 * lmaxchg
    BvOpts o;
    real *y_out;                 // [nout][neq][nsys] or nullptr
    real *rout;                  // [4][nsys]  rwork(11:14)
    int *iout;                     // [12][nsys] iwork(11:22)
    int lenrw, leniw;              // iwork(17:18), M:1268-1273
    int general;                   // 1: full-driver kernel (itask 2-5, istate = 3); 0: lean itask = 1
    int *next_sys;                 // thread-per-system kernels: counter handing out the systems beyond the
                                   // grid (zeroed before the launch), or nullptr: one system per thread
};

struct DkyArgs {
    int nsys, k, ml, mu;
    real t;
    const real *t_ps;
    const real *dstate;
    const int *istate_store;
    real *dky;  // [neq][nsys]
    int *iflag;   // [nsys]
};

struct ProblemEntry {
    const char *name;
    int neq, npar;
    int mode;  // 0 = one system per thread, 1 = one system per warp
    int (*supports)(int mf, int ml, int mu);
    // total doubles / ints of persistent state for a batch of nsys systems
    void (*state_size)(int mf, int ml, int mu, size_t nsys, size_t *nd, size_t *ni);
    int (*launch)(int mf, const KArgs &a, cudaStream_t st);
    int (*launch_dky)(int mf, const DkyArgs &a, cudaStream_t st);
    // istate = 3 with another method flag of the SAME family (meth): carry the part of the state that
    // does not depend on miter / jsv -- Nordsieck array, acor copy, tau, scalars, counters -- over into
    // the layout of mf_new (zero-filled by the caller; the Newton matrix is rebuilt at the restart,
    // M:2748).  Returns 0, or -1000 if the pair is not convertible.
    int (*state_convert)(int mf_old, int mf_new, int ml, int mu, size_t nsys, const real *d_old,
                         const int *i_old, real *d_new, int *i_new, cudaStream_t st);
};

// Tag a problem library passes to bvode_register_problem: it must have been compiled against
// these very structs (bump the leading number when KArgs / DkyArgs / ProblemEntry change meaning).
#define BVODE_PLUGIN_ABI ((int)(5 * 1000000 + (sizeof(KArgs) % 1000) * 1000 + sizeof(ProblemEntry) % 1000))
