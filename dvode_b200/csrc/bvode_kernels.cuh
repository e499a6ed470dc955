// bvode_kernels.cuh -- kernel entry points, state load/store and the per-build registry.
// Included by bvode_kernels_fast.cu and bvode_kernels_strict.cu (same source, two builds).
#pragma once
#include <stdlib.h>
#include "bvode_internal.h"
#include "policy_thread.cuh"
#if !BVODE_REAL_IS_FLOAT  // the REAL32 build (libbvode32.so) carries the thread-per-system kernels only
#include "policy_warp.cuh"
#include "policy_warp2.cuh"
#endif
#include "problems.cuh"

namespace BVODE_NS {

// ------------------------------------------------------------ state layout ------
// The reference keeps the solver state in the caller's rwork(21:) / iwork(31:) and in
// dvode_data_t.  Here it is structure-of-arrays in HBM: slot k of system s is at
// dstate[k * nsys + s], so a warp's 32 loads of one slot are one coalesced 256-byte line.
// visit_state enumerates the persistent fields in a fixed order; Loader / Storer / Counter
// are the three visitors.
template <class ENG, class V>
BV_DEV void visit_state(ENG &e, V &v) {
    Ctl &c = e.c;
#pragma unroll
    for (int j = 0; j < ENG::kLmax; j++)
#pragma unroll
        for (int i = 0; i < ENG::NLOC; i++) v.dv(e.hot.yh(j, i));
    ENG::POLICY::visit_acsav(e, v);
#pragma unroll
    for (int i = 0; i < ENG::NLOC; i++) v.dv(e.acor[i]);
    ENG::POLICY::visit_lin(e, v);
#pragma unroll
    for (int j = 1; j <= ENG::kLmax; j++) v.d(e.hot.tau(j));
    v.d(c.h); v.d(c.hscal); v.d(e.hot.sd(SD_HNEW)); v.d(e.hot.sd(SD_HU)); v.d(c.eta); v.d(e.hot.sd(SD_ETAMAX)); v.d(c.tn);
    v.d(c.rc); v.d(c.prl1); v.d(e.hot.sd(SD_CRATE)); v.d(e.hot.sd(SD_CONP));
    v.d(e.hot.sd(SD_ETAQ)); v.d(e.hot.sd(SD_ETAQM1)); v.d(e.hot.tq(1));
    v.i(c.nq); v.i(c.nqwait); v.i(c.newq); v.i(c.newh); v.i(c.jstart); v.i(c.icf); v.i(c.ipup);
    v.i(e.hot.si(SI_NSLJ)); v.i(e.hot.si(SI_NSLP)); v.i(e.hot.si(SI_NHNIL)); v.i(e.hot.si(SI_INIT));
    v.i(e.hot.si(SI_NST)); v.i(e.hot.si(SI_NFE)); v.i(e.hot.si(SI_NJE)); v.i(e.hot.si(SI_NLU)); v.i(e.hot.si(SI_NNI)); v.i(e.hot.si(SI_NCFN)); v.i(e.hot.si(SI_NETF));
    v.i(e.hot.si(SI_NQU));
}

// (T is double / int, or their volatile forms for fields a policy parks in shared memory)
struct Loader {
    const real *dp;
    const int *ip;
    size_t stride;
    int kd = 0, ki = 0;
    template <class T> BV_DEV void d(T &x) { x = dp[(size_t)kd * stride]; kd++; }
    template <class T> BV_DEV void i(T &x) { x = ip[(size_t)ki * stride]; ki++; }
    BV_DEV void skip_d(int n) { kd += n; }
    template <class T> BV_DEV void dv(T &x) { d(x); }  // thread-per-system: vector elements are plain slots
    template <class T> BV_DEV void iv(T &x) { i(x); }
};
struct Storer {
    real *dp;
    int *ip;
    size_t stride;
    int kd = 0, ki = 0;
    template <class T> BV_DEV void d(T &x) { dp[(size_t)kd * stride] = x; kd++; }
    template <class T> BV_DEV void i(T &x) { ip[(size_t)ki * stride] = x; ki++; }
    BV_DEV void skip_d(int n) { kd += n; }
    template <class T> BV_DEV void dv(T &x) { d(x); }
    template <class T> BV_DEV void iv(T &x) { i(x); }
};

// Per-system optional inputs (BVODE_FLAG_PER_SYSTEM): what the reference reads from each solver
// instance's rwork(1), rwork(5:7), iwork(5:6) (M:729-751, 1174-1238), defaulted on the host.
BV_DEV void load_opts_ps(BvOpts &o, const KArgs &a, size_t ns, size_t sys) {
    o.h0 = a.optd_ps[sys];
    o.hmax_inv = a.optd_ps[ns + sys];
    o.hmin = a.optd_ps[2 * ns + sys];
    o.tcrit = a.optd_ps[3 * ns + sys];
    o.maxord = a.opti_ps[sys];
    o.mxstep = a.opti_ps[ns + sys];
    o.jreset = a.opti_ps[2 * ns + sys] & 1;
    o.lmaxchg = (a.opti_ps[2 * ns + sys] >> 1) & 1;
}

// ------------------------------------------------------- thread-per-system ------
#ifndef BVODE_THREAD_TPB
#define BVODE_THREAD_TPB 32
#endif
constexpr int kThreadTPB = BVODE_THREAD_TPB;

// Each thread starts with system blockIdx*TPB + tid.  When the grid is smaller than the batch
// (a.next_sys != nullptr: the launcher sizes it to what is resident at once), a thread whose
// system is finished takes the next unclaimed index from a global counter and goes on inside
// the same step loop -- the finished systems of a warp are replaced instead of idling until the
// slowest one is done (mean / max attempts in a warp is 0.91 on the headline batch; far less
// on heavy-tailed batches).  Which thread integrates which system does not change any result.
// REFILL = false is the default instantiation (the hand-over code outside the step loop); the
// compacting one is chosen by BVODE_FLAG_COMPACT: it pays off when the systems of a batch differ
// widely in cost (measured: see DESIGN.md), and loses ~15 % on the homogeneous headline batch,
// where replacing lanes at different times de-correlates the rare paths of a warp's 32 systems.
// BVODE_THREAD_MAXNREG (experiments): a register cap between ptxas's own steps (168 at 12 warps / SM, 128 above)
#ifdef BVODE_THREAD_MAXNREG
#define BVODE_THREAD_BOUNDS __maxnreg__(BVODE_THREAD_MAXNREG)
#else
#define BVODE_THREAD_BOUNDS __launch_bounds__(kThreadTPB, POL::kMinBlocks)
#endif
template <class POL, bool GENERAL, bool REFILL>
__global__ void BVODE_THREAD_BOUNDS vode_thread_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ real smem_thread[];
    constexpr int N = POL::N;
    using Prob = typename POL::Prob;

    Engine<POL> e;
    e.hot.bind(smem_thread, threadIdx.x);
    e.o = a.o;
    if constexpr (GENERAL) e.nact = (a.o.nact > 0 && a.o.nact < N) ? a.o.nact : N;
#pragma unroll
    for (int i = 0; i < N; i++) {  // N <= 8: by-value kernel parameters; larger: the [neq] arrays in HBM
        e.tol.rv[i] = (N <= 8) ? a.rtolv[i < 8 ? i : 0] : a.rtol[i];
        e.tol.av[i] = (N <= 8) ? a.atolv[i < 8 ? i : 0] : a.atol[i];
    }

    int sys = a.sys0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int sys_end = a.sys0 + a.count;  // this launch integrates systems [sys0, sys0 + count)
    int nemit = 0;
    real tout0 = BV_R(0.0);
    const size_t ns = (size_t)a.nsys;

    // ---- the kernel epilogue for one system: y, t, istate, the persistent state, optional outputs
    auto store = [&](int ist, real t) {
        if (ist == IST_ILLEGAL_ENTRY) {  // illegal input at a call entry (messages 23-25, 27)
            a.istate[sys] = IST_ILLEGAL;
            if (nemit == 0) return;      // nothing was done: y, t and the state stay as they were
            ist = IST_ILLEGAL;
        }
        if (a.y_out) {  // a system that failed keeps its last state in the remaining output slots
            for (int k2 = nemit; k2 < a.nout; k2++) {
#pragma unroll
                for (int i = 0; i < N; i++) a.y_out[((size_t)k2 * N + i) * ns + sys] = e.y[i];
            }
        }
        real tolsf = BV_R(0.0);
        if (ist == IST_TOLSF || (ist == IST_ILLEGAL && e.hot.si(SI_NST) == 0))
            tolsf = BV_R(2.0) * kUround * e.norm_y0();  // M:1626, 1632, 1641
#pragma unroll
        for (int i = 0; i < N; i++) a.y[(size_t)i * ns + sys] = e.y[i];
        a.t[sys] = t;
        a.istate[sys] = ist;
        {
            Storer st{a.dstate + sys, a.istate_store + sys, ns};
            visit_state(e, st);
        }
        // optional outputs, M:1670-1684 (success) / M:1708-1723 (failure)
        a.rout[0 * ns + sys] = e.hot.sd(SD_HU);
        a.rout[1 * ns + sys] = (ist == IST_OK) ? e.hot.sd(SD_HNEW) : e.c.h;
        a.rout[2 * ns + sys] = e.c.tn;
        a.rout[3 * ns + sys] = tolsf;
        a.iout[0 * ns + sys] = e.hot.si(SI_NST);
        a.iout[1 * ns + sys] = e.hot.si(SI_NFE);
        a.iout[2 * ns + sys] = e.hot.si(SI_NJE);
        a.iout[3 * ns + sys] = e.hot.si(SI_NQU);
        a.iout[4 * ns + sys] = (ist == IST_OK) ? e.c.newq : e.c.nq;
        a.iout[5 * ns + sys] =
            (ist == IST_ERRTEST || ist == IST_CONVFAIL) ? POL::imxer(e.acor, e.ewt) : 0;
        a.iout[6 * ns + sys] = a.lenrw + (GENERAL ? N * (e.o.maxord - a.o.maxord) : 0);  // M:1246: 20 + nyh*(maxord+1) + ...
        a.iout[7 * ns + sys] = a.leniw;
        a.iout[8 * ns + sys] = e.hot.si(SI_NLU);
        a.iout[9 * ns + sys] = e.hot.si(SI_NNI);
        a.iout[10 * ns + sys] = e.hot.si(SI_NCFN);
        a.iout[11 * ns + sys] = e.hot.si(SI_NETF);
    };
    // ---- the kernel prologue for system `sys`: 0 = nothing to do for it, 1 = first call, 2 = continuation
    auto load = [&]() -> int {
        const int ist = a.istate[sys];
        if (ist != 1 && ist != 2 && !(GENERAL && ist == 3)) {  // failed / illegal: left untouched
            // a schedule split over several launches (nout > BVODE_MAX_NOUT): a system that failed in an
            // earlier launch keeps its last state in the output slots of this one as well
            if (ist < 0 && a.y_out) {
                for (int k2 = 0; k2 < a.nout; k2++) {
#pragma unroll
                    for (int i = 0; i < N; i++) a.y_out[((size_t)k2 * N + i) * ns + sys] = a.y[(size_t)i * ns + sys];
                }
            }
            return 0;
        }
#pragma unroll
        for (int k = 0; k < Prob::NPAR; k++) e.par[k] = a.params[(size_t)k * ns + sys];
        const real t = a.t[sys];
        tout0 = a.tout_ps ? a.tout_ps[sys] : a.touts[0];
        nemit = 0;
        if constexpr (GENERAL) {  // per-system tolerances / optional inputs (M:729-751, 928-965)
            if (a.tol_ps) {
#pragma unroll
                for (int i = 0; i < N; i++) {
                    e.tol.rv[i] = a.tol_ps[(size_t)i * ns + sys];
                    e.tol.av[i] = a.tol_ps[(size_t)(N + i) * ns + sys];
                }
            }
            if (a.optd_ps) load_opts_ps(e.o, a, ns, (size_t)sys);
        }
        if (ist == 1) {
            if (tout0 == t) return 0;  // M:1101
#pragma unroll
            for (int i = 0; i < N; i++) e.y[i] = a.y[(size_t)i * ns + sys];
            const int r = e.init_problem(t, tout0);
            if (r < 0) {  // illegal input detected on the device (messages 21, 22): y, t untouched
                a.istate[sys] = r;
                return 0;
            }
            return 1;
        }
        Loader ld{a.dstate + sys, a.istate_store + sys, ns};
        visit_state(e, ld);
        e.c.L = e.c.nq + 1;
        e.c.kflag = 0;
        e.c.jcur = 0;
        if (e.hot.si(SI_INIT) != 1) {  // M:1102-1106
            a.istate[sys] = IST_ILLEGAL;
            return 0;
        }
        if constexpr (GENERAL) {
            if (e.nact != N) {
                // neq was reduced (M:1127-1135): the dropped components are what the caller's y array holds for
                // them -- the last output values -- and stay that: f of the kept equations may read them
#pragma unroll
                for (int i = 0; i < N; i++)
                    if (i >= e.nact) e.hot.yh(0, i) = a.y[(size_t)i * ns + sys];
            }
        }
        if (GENERAL && ist == 3) e.restart_istate3(true);  // M:1383-1390
        return 2;
    };

    e.template run<GENERAL, REFILL>(
        a.nout, [&](int k) { return (k == 0) ? tout0 : a.touts[k]; },
        [&](int k, const real (&yk)[N]) {
            nemit = k + 1;
            if (a.y_out) {
#pragma unroll
                for (int i = 0; i < N; i++) a.y_out[((size_t)k * N + i) * ns + sys] = yk[i];
            }
        },
        [&](bool have, int ist, real t) -> int {
            if constexpr (REFILL) {
                if (have) {
                    store(ist, t);
                    sys = a.sys0 + (int)(gridDim.x * blockDim.x) + atomicAdd(a.next_sys, 1);
                }
                int got = 0;
#pragma unroll 1
                while (got == 0 && sys < sys_end) {
                    got = load();
                    if (got == 0) sys = a.sys0 + (int)(gridDim.x * blockDim.x) + atomicAdd(a.next_sys, 1);
                }
                return got;
            } else {  // one system per thread and launch
                if (have) {
                    store(ist, t);
                    return 0;
                }
                return (sys < sys_end) ? load() : 0;
            }
        });
}

// dvindy for any k (M:1878-1959) from the stored state; one thread per system.
template <class POL>
__global__ void __launch_bounds__(kThreadTPB) dvindy_thread_kernel(const __grid_constant__ DkyArgs a) {
    extern __shared__ real smem_thread[];
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= a.nsys) return;
    constexpr int N = POL::N;
    Engine<POL> e;
    e.hot.bind(smem_thread, threadIdx.x);
    Loader ld{a.dstate + sys, a.istate_store + sys, (size_t)a.nsys};
    visit_state(e, ld);
    e.c.L = e.c.nq + 1;
    const real t = a.t_ps ? a.t_ps[sys] : a.t;
    real dky[N];
    const int iflag = e.dvindy_k(t, a.k, dky);
#pragma unroll
    for (int i = 0; i < N; i++) a.dky[(size_t)i * a.nsys + sys] = dky[i];
    a.iflag[sys] = iflag;
}

#if !BVODE_REAL_IS_FLOAT
// ---------------------------------------------------------- warp-per-system ------
// State layout in HBM for one-system-per-warp kernels (all in the same dstate / istate_store
// allocations, regions in this order):
//   scalars   [kScalarD][nsys]            one copy per system (lane 0 writes, all lanes read)
//   per-lane  [kLaneSlots][nsys][32]      lane l of system s at (k*nsys + s)*32 + l: coalesced
//   blocks    [nsys][nblock]              shared-memory image of the band arrays
template <class POL>
__host__ __device__ constexpr int scalar_d() { return POL::kLmax + 14; }  // tau(1:lmax) + the 14 scalars of visit_state
constexpr int kScalarI = 19;
#ifndef BVODE_WARP_TPB
#define BVODE_WARP_TPB 128
#endif
constexpr int kWarpTPB = BVODE_WARP_TPB;  // threads per CTA = 32 x systems per CTA (a policy may choose its own)
template <class POL, class = void>
struct WarpTpb { static constexpr int value = kWarpTPB; };
template <class POL>
struct WarpTpb<POL, decltype((void)POL::kTPB)> { static constexpr int value = POL::kTPB; };

struct WarpLoader {
    const real *ds, *dl, *db;
    const int *is, *il;
    size_t nsys, sys;
    int ln, nblock;
    int kd = 0, kl = 0, ki = 0, kil = 0;
    template <class T> BV_DEV void d(T &x) { x = ds[(size_t)kd * nsys + sys]; kd++; }
    template <class T> BV_DEV void i(T &x) { x = is[(size_t)ki * nsys + sys]; ki++; }
    template <class T> BV_DEV void dv(T &x) { x = dl[((size_t)kl * nsys + sys) * 32 + ln]; kl++; }
    template <class T> BV_DEV void iv(T &x) { x = il[((size_t)kil * nsys + sys) * 32 + ln]; kil++; }
    BV_DEV void skip_dv(int n) { kl += n; }
    BV_DEV void block(real *sm, int n) {
        for (int k = ln; k < n; k += 32) sm[k] = db[sys * (size_t)nblock + k];
        __syncwarp();
    }
};
struct WarpStorer {
    real *ds, *dl, *db;
    int *is, *il;
    size_t nsys, sys;
    int ln, nblock;
    int kd = 0, kl = 0, ki = 0, kil = 0;
    template <class T> BV_DEV void d(T &x) { const real v = x; if (ln == 0) ds[(size_t)kd * nsys + sys] = v; kd++; }
    template <class T> BV_DEV void i(T &x) { const int v = x; if (ln == 0) is[(size_t)ki * nsys + sys] = v; ki++; }
    template <class T> BV_DEV void dv(T &x) { dl[((size_t)kl * nsys + sys) * 32 + ln] = x; kl++; }
    template <class T> BV_DEV void iv(T &x) { il[((size_t)kil * nsys + sys) * 32 + ln] = x; kil++; }
    BV_DEV void skip_dv(int n) { kl += n; }
    BV_DEV void block(real *sm, int n) {
        __syncwarp();
        for (int k = ln; k < n; k += 32) db[sys * (size_t)nblock + k] = sm[k];
    }
};

template <class POL>
struct WarpLayout {
    static constexpr int N = POL::N;
    // doubles of one system's saved block (band policy: the band image without the scratch of the inverse)
    static int nblock(int ml, int mu) {
        if constexpr (POL::kBand) return POL::image_doubles(ml, mu, POL::SAVEJ);
        else return POL::smem_doubles(ml, mu, POL::SAVEJ);
    }
    // doubles of shared memory per system
    static int nsmem(int ml, int mu) { return POL::smem_doubles(ml, mu, POL::SAVEJ); }
    static void sizes(int ml, int mu, size_t nsys, size_t *nd, size_t *ni) {
        *nd = nsys * ((size_t)scalar_d<POL>() + (size_t)POL::kLaneSlots * 32 + (size_t)nblock(ml, mu));
        *ni = nsys * ((size_t)kScalarI + (size_t)POL::kLaneIntSlots * 32);
    }
};

// The three state regions of a warp kernel (see the layout above), recomputed from the kernel arguments where a
// system is loaded or stored instead of being kept in registers across the step loop.
template <class POL>
struct WarpRegions {
    const real *dl, *db;
    const int *il;
    int nblock;
    BV_DEV WarpRegions(const real *dstate, const int *istore, size_t nsys, int ml, int mu) {
        dl = dstate + (size_t)scalar_d<POL>() * nsys;
        db = dl + (size_t)POL::kLaneSlots * nsys * 32;
        il = istore + (size_t)kScalarI * nsys;
        if constexpr (POL::kBand) nblock = POL::band_ld(ml, mu) * POL::N + (POL::SAVEJ ? (ml + mu + 1) * POL::N : 0);
        else if constexpr (POL::kPlain) nblock = 0;
        else nblock = POL::kBlockDoubles;
    }
};

template <class POL>
BV_DEV void warp_bind(Engine<POL> &e, const real *dstate, const int *istore, size_t nsys,
                      size_t sys, int ln, int warp, int ml, int mu, real *smem) {
    const WarpRegions<POL> rg(dstate, istore, nsys, ml, mu);
    int nb = rg.nblock;  // shared-memory doubles per system
    if constexpr (POL::kBand) nb = POL::smem_doubles(ml, mu, POL::SAVEJ);
    // per-warp control state: after the kWarpTPB/32 matrix images of the CTA
    e.hot.bind(smem + (size_t)(WarpTpb<POL>::value / 32) * nb + (size_t)warp * POL::kWarpHotDoubles, ln);
    if constexpr (POL::kBand) {
        e.lin.ml = ml;
        e.lin.mu = mu;
        e.lin.ld = POL::band_ld(ml, mu);
        e.lin.abd = smem + (size_t)warp * nb;
        e.lin.wjs = e.lin.abd + e.lin.ld * POL::N;
        e.lin.inv = e.lin.abd + ((POL::image_doubles(ml, mu, POL::SAVEJ) + 1) & ~1);
    } else if constexpr (POL::kPlain) {
        (void)ml; (void)mu; (void)smem; (void)warp; (void)sys; (void)ln;
    } else {
        e.lin.p = smem + (size_t)warp * nb;
        e.lin.perm = reinterpret_cast<int *>(e.lin.p + POL::kPermOff);
        e.lin.wjs.p = const_cast<real *>(rg.dl) + ((size_t)POL::kLaneSlotWjs * nsys + sys) * 32 + ln;
        e.lin.wjs.stride = nsys * 32;
    }
}

// One-system-per-warp kernels are latency-bound (dependent shuffle / FP64 chains, scalar
// control replicated on 32 lanes): resident warps beat registers.  Measured on B200, 65536
// systems, systems/s by CTAs per SM (POL::kMinBlocks; 4 warps each):
//                      4        5        6        8
//   mech32 mf=22      73.2k    83.1k    91.6k    87.0k
//   advection mf=24   1.37M    1.41M    1.60M    1.75M
// (the replicated control state lives once per warp in shared memory, see WarpCommon::Hot;
// before that: 6 CTAs 88.8k / 1.53M).
// BVODE_WARP_MAXNREG (experiments): a register cap between ptxas's own steps (see BVODE_THREAD_MAXNREG)
#ifdef BVODE_WARP_MAXNREG
#define BVODE_WARP_BOUNDS __maxnreg__(BVODE_WARP_MAXNREG)
#else
#define BVODE_WARP_BOUNDS __launch_bounds__(WarpTpb<POL>::value, POL::kMinBlocks)
#endif
template <class POL, bool GENERAL>
__global__ void BVODE_WARP_BOUNDS
vode_warp_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ real smem[];
    const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    const size_t sys = (size_t)a.sys0 + (size_t)blockIdx.x * (WarpTpb<POL>::value / 32) + warp;
    const size_t nsys = (size_t)a.nsys;
    if (sys >= (size_t)(a.sys0 + a.count)) return;  // whole warp
    constexpr int N = POL::N;
    using Prob = typename POL::Prob;

    Engine<POL> e;
    e.o = a.o;
    if constexpr (GENERAL) e.nact = (a.o.nact > 0 && a.o.nact < N) ? a.o.nact : N;
    constexpr int NL = POL::NLOC;  // components per lane: component ln + 32*i in slot i
    if constexpr (NL == 1) {
        e.tol.r = (ln < N) ? a.rtol[ln] : BV_R(0.0);
        e.tol.a = (ln < N) ? a.atol[ln] : BV_R(1.0);  // padding lanes: finite weight, value always 0
    } else {
#pragma unroll
        for (int i = 0; i < NL; i++) {
            e.tol.r[i] = (ln + 32 * i < N) ? a.rtol[ln + 32 * i] : BV_R(0.0);
            e.tol.a[i] = (ln + 32 * i < N) ? a.atol[ln + 32 * i] : BV_R(1.0);
        }
    }
    warp_bind(e, a.dstate, a.istate_store, nsys, sys, ln, warp, a.o.ml, a.o.mu, smem);
    int nemit = 0;
    real tout0 = BV_R(0.0);

    // ---- the kernel prologue: 0 = nothing to do for this system, 1 = first call, 2 = continuation
    auto load = [&]() -> int {
        const int ist = a.istate[sys];
        if (ist != 1 && ist != 2 && !(GENERAL && ist == 3)) {
            if (ist < 0 && a.y_out) {  // failed in an earlier launch of a split schedule: see the thread kernel
                for (int k2 = 0; k2 < a.nout; k2++) {
#pragma unroll
                    for (int i = 0; i < NL; i++)
                        if (ln + 32 * i < N)
                            a.y_out[((size_t)k2 * N + ln + 32 * i) * nsys + sys] = a.y[(size_t)(ln + 32 * i) * nsys + sys];
                }
            }
            return 0;
        }
        if constexpr (GENERAL) {  // per-system tolerances / optional inputs (M:729-751, 928-965)
            if (a.tol_ps) {
                if constexpr (NL == 1) {
                    if (ln < N) {
                        e.tol.r = a.tol_ps[(size_t)ln * nsys + sys];
                        e.tol.a = a.tol_ps[(size_t)(N + ln) * nsys + sys];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NL; i++) {
                        if (ln + 32 * i < N) {
                            e.tol.r[i] = a.tol_ps[(size_t)(ln + 32 * i) * nsys + sys];
                            e.tol.a[i] = a.tol_ps[(size_t)(N + ln + 32 * i) * nsys + sys];
                        }
                    }
                }
            }
            if (a.optd_ps) load_opts_ps(e.o, a, nsys, sys);
        }
        Prob::load_par(e.par, a.params, nsys, sys, ln);
        const real t = a.t[sys];
        tout0 = a.tout_ps ? a.tout_ps[sys] : a.touts[0];
        if (ist == 1) {
            if (tout0 == t) return 0;  // M:1101
#pragma unroll
            for (int i = 0; i < NL; i++) e.y[i] = (ln + 32 * i < N) ? a.y[(size_t)(ln + 32 * i) * nsys + sys] : BV_R(0.0);
            const int r = e.init_problem(t, tout0);
            if (r < 0) {
                if (ln == 0) a.istate[sys] = r;
                return 0;
            }
            return 1;
        }
        const WarpRegions<POL> rg(a.dstate, a.istate_store, nsys, a.o.ml, a.o.mu);
        WarpLoader ld{a.dstate, rg.dl, rg.db, a.istate_store, rg.il, nsys, sys, ln, rg.nblock};
        visit_state(e, ld);
        e.c.L = e.c.nq + 1;
        e.c.kflag = 0;
        e.c.jcur = 0;
        if (e.hot.si(SI_INIT) != 1) {
            if (ln == 0) a.istate[sys] = IST_ILLEGAL;
            return 0;
        }
        if constexpr (POL::kBand) POL::after_load(e);  // fast build: the inverse is not part of the saved state
        if constexpr (GENERAL) {
            if (e.nact != N) {  // neq was reduced: see the thread kernel
#pragma unroll
                for (int i = 0; i < NL; i++)
                    if (ln + 32 * i >= e.nact && ln + 32 * i < N) e.hot.yh(0, i) = a.y[(size_t)(ln + 32 * i) * nsys + sys];
            }
        }
        if (GENERAL && ist == 3) e.restart_istate3(ln == 0);
        return 2;
    };
    // ---- the kernel epilogue
    auto store = [&](int ist, real t) {
        if (ist == IST_ILLEGAL_ENTRY) {  // illegal input at a call entry (messages 23-25, 27)
            if (ln == 0) a.istate[sys] = IST_ILLEGAL;
            if (nemit == 0) return;
            ist = IST_ILLEGAL;
        }
        if (a.y_out) {
            for (int k2 = nemit; k2 < a.nout; k2++) {
#pragma unroll
                for (int i = 0; i < NL; i++)
                    if (ln + 32 * i < N) a.y_out[((size_t)k2 * N + ln + 32 * i) * nsys + sys] = e.y[i];
            }
        }
        real tolsf = BV_R(0.0);
        if (ist == IST_TOLSF || (ist == IST_ILLEGAL && e.hot.si(SI_NST) == 0))
            tolsf = BV_R(2.0) * kUround * e.norm_y0();
        const int imx = (ist == IST_ERRTEST || ist == IST_CONVFAIL) ? POL::imxer(e.acor, e.ewt) : 0;

#pragma unroll
        for (int i = 0; i < NL; i++)
            if (ln + 32 * i < N) a.y[(size_t)(ln + 32 * i) * nsys + sys] = e.y[i];
        {
            const WarpRegions<POL> rg(a.dstate, a.istate_store, nsys, a.o.ml, a.o.mu);
            WarpStorer st{a.dstate, const_cast<real *>(rg.dl), const_cast<real *>(rg.db),
                          a.istate_store, const_cast<int *>(rg.il), nsys, sys, ln, rg.nblock};
            visit_state(e, st);
        }
        if (ln == 0) {
            a.t[sys] = t;
            a.istate[sys] = ist;
            a.rout[0 * nsys + sys] = e.hot.sd(SD_HU);
            a.rout[1 * nsys + sys] = (ist == IST_OK) ? e.hot.sd(SD_HNEW) : e.c.h;
            a.rout[2 * nsys + sys] = e.c.tn;
            a.rout[3 * nsys + sys] = tolsf;
            a.iout[0 * nsys + sys] = e.hot.si(SI_NST);
            a.iout[1 * nsys + sys] = e.hot.si(SI_NFE);
            a.iout[2 * nsys + sys] = e.hot.si(SI_NJE);
            a.iout[3 * nsys + sys] = e.hot.si(SI_NQU);
            a.iout[4 * nsys + sys] = (ist == IST_OK) ? e.c.newq : e.c.nq;
            a.iout[5 * nsys + sys] = imx;
            a.iout[6 * nsys + sys] = a.lenrw + (GENERAL ? N * (e.o.maxord - a.o.maxord) : 0);
            a.iout[7 * nsys + sys] = a.leniw;
            a.iout[8 * nsys + sys] = e.hot.si(SI_NLU);
            a.iout[9 * nsys + sys] = e.hot.si(SI_NNI);
            a.iout[10 * nsys + sys] = e.hot.si(SI_NCFN);
            a.iout[11 * nsys + sys] = e.hot.si(SI_NETF);
        }
    };

    e.template run<GENERAL, false>(
        a.nout, [&](int k) { return (k == 0) ? tout0 : a.touts[k]; },
        [&](int k, const real (&yk)[NL]) {
            nemit = k + 1;
            if (a.y_out) {
#pragma unroll
                for (int i = 0; i < NL; i++)
                    if (ln + 32 * i < N) a.y_out[((size_t)k * N + ln + 32 * i) * nsys + sys] = yk[i];
            }
        },
        [&](bool have, int ist, real t) -> int {  // one system per warp and launch
            if (have) {
                store(ist, t);
                return 0;
            }
            return load();
        });
}

template <class POL>
__global__ void __launch_bounds__(WarpTpb<POL>::value) dvindy_warp_kernel(const __grid_constant__ DkyArgs a) {
    extern __shared__ real smem[];
    const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    const size_t sys = (size_t)blockIdx.x * (WarpTpb<POL>::value / 32) + warp;
    const size_t nsys = (size_t)a.nsys;
    if (sys >= nsys) return;
    constexpr int N = POL::N;
    Engine<POL> e;
    warp_bind(e, a.dstate, a.istate_store, nsys, sys, ln, warp, a.ml, a.mu, smem);
    const WarpRegions<POL> rg(a.dstate, a.istate_store, nsys, a.ml, a.mu);
    WarpLoader ld{a.dstate, rg.dl, rg.db, a.istate_store, rg.il, nsys, sys, ln, rg.nblock};
    visit_state(e, ld);
    e.c.L = e.c.nq + 1;
    const real t = a.t_ps ? a.t_ps[sys] : a.t;
    real dky[POL::NLOC];
    const int iflag = e.dvindy_k(t, a.k, dky);
#pragma unroll
    for (int i = 0; i < POL::NLOC; i++)
        if (ln + 32 * i < N) a.dky[(size_t)(ln + 32 * i) * nsys + sys] = dky[i];
    if (ln == 0) a.iflag[sys] = iflag;
}

#endif  // !BVODE_REAL_IS_FLOAT
// --------------------------------------------------------------- launchers ------
// Grid of the compacting thread-per-system kernel: exactly the resident capacity (SMs x CTAs per
// SM); finished threads claim the remaining systems through a.next_sys.
template <class K>
static int resident_grid(K kernel, size_t smem) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreadTPB, smem);
    return (sms > 0 && per_sm > 0) ? sms * per_sm : 0;
}
template <class POL>
static int launch_thread(const KArgs &a0, cudaStream_t st) {
    KArgs a = a0;
    static_assert(POL::kTPB == kThreadTPB, "shared-memory stride");
    size_t smem = POL::kSmemBytes;
    if (const char *pad = getenv("BVODE_PAD_SMEM")) smem += (size_t)atol(pad);  // occupancy experiments
    const int need = (a.count + kThreadTPB - 1) / kThreadTPB;
    if (a.general) {
        a.next_sys = nullptr;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(vode_thread_kernel<POL, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        vode_thread_kernel<POL, true, false><<<need, kThreadTPB, smem, st>>>(a);
        return (int)cudaGetLastError();
    }
    if (a.next_sys) {  // BVODE_FLAG_COMPACT
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(vode_thread_kernel<POL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        static int cap_of[64];  // per instantiation and device (0 = not asked yet); benign if two threads race
        int dev = 0;
        cudaGetDevice(&dev);
        int capacity = (dev >= 0 && dev < 64) ? cap_of[dev] : 0;
        if (capacity == 0) {
            capacity = resident_grid(vode_thread_kernel<POL, false, true>, smem);
            if (dev >= 0 && dev < 64) cap_of[dev] = capacity;
        }
        if (capacity > 0 && need > capacity) {
            vode_thread_kernel<POL, false, true><<<capacity, kThreadTPB, smem, st>>>(a);
            return (int)cudaGetLastError();
        }
        a.next_sys = nullptr;  // the whole batch is resident at once: nothing to hand out
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(vode_thread_kernel<POL, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vode_thread_kernel<POL, false, false><<<need, kThreadTPB, smem, st>>>(a);
    return (int)cudaGetLastError();
}
template <class POL>
static int launch_dky_thread(const DkyArgs &a, cudaStream_t st) {
    const int grid = (a.nsys + kThreadTPB - 1) / kThreadTPB;
    if (POL::kSmemBytes > 48 * 1024)
        cudaFuncSetAttribute(dvindy_thread_kernel<POL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POL::kSmemBytes);
    dvindy_thread_kernel<POL><<<grid, kThreadTPB, POL::kSmemBytes, st>>>(a);
    return (int)cudaGetLastError();
}
#if !BVODE_REAL_IS_FLOAT
template <class POL>
static int launch_warp(const KArgs &a, cudaStream_t st) {
    const int spb = WarpTpb<POL>::value / 32;
    const int grid = (a.count + spb - 1) / spb;
    const size_t smem = sizeof(real) * spb * ((size_t)WarpLayout<POL>::nsmem(a.o.ml, a.o.mu) + POL::kWarpHotDoubles +
                                               ScratchDoublesOf<typename POL::Prob>::value);
    // Shared memory / L1 split: ask for just what POL::kMinBlocks resident CTAs need (+1 KB each that the
    // system reserves), so that the rest of the 256 KB stays L1 -- the kernels' register spills and the
    // problems' coefficient tables live there (ncu r2d: with the maximum carve-out spills went to L2).
    int carve = (int)((POL::kMinBlocks * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (const char *cv = getenv("BVODE_CARVEOUT")) carve = atoi(cv);  // experiments
    if (carve > 100) carve = 100;
    if (a.general) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(vode_warp_kernel<POL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (carve >= 0) cudaFuncSetAttribute(vode_warp_kernel<POL, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        vode_warp_kernel<POL, true><<<grid, WarpTpb<POL>::value, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(vode_warp_kernel<POL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (carve >= 0) cudaFuncSetAttribute(vode_warp_kernel<POL, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        vode_warp_kernel<POL, false><<<grid, WarpTpb<POL>::value, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}
template <class POL>
static int launch_dky_warp(const DkyArgs &a, cudaStream_t st) {
    const int spb = WarpTpb<POL>::value / 32;
    const int grid = (a.nsys + spb - 1) / spb;
    const size_t smem = sizeof(real) * spb * ((size_t)WarpLayout<POL>::nsmem(a.ml, a.mu) + POL::kWarpHotDoubles);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dvindy_warp_kernel<POL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dvindy_warp_kernel<POL><<<grid, WarpTpb<POL>::value, smem, st>>>(a);
    return (int)cudaGetLastError();
}

#endif  // !BVODE_REAL_IS_FLOAT
// ------------------------------------------------------------------ dispatch ------
// mf = jsv * (10*meth + miter), M:999-1046.  jsv only matters where a Jacobian is stored
// (miter 1, 2, 4, 5).
struct MfParts {
    int meth, miter;
    bool savej;
};
static inline MfParts mf_parts(int mf) {
    const int a = mf < 0 ? -mf : mf;
    MfParts p;
    p.meth = a / 10;
    p.miter = a - 10 * p.meth;
    p.savej = (mf > 0) && (p.miter == 1 || p.miter == 2 || p.miter == 4 || p.miter == 5);
    return p;
}

// ---- thread-per-system problems: meth 1, 2; miter 0 (functional), 1, 2 (dense), 3 (diagonal)
template <class PROB, int METH>
static int thread_dispatch_m(const MfParts &m, const KArgs &a, cudaStream_t st) {
    switch (m.miter) {
    case 0: return launch_thread<ThreadPolicy<PROB, 0, false, METH>>(a, st);
    case 1: return m.savej ? launch_thread<ThreadPolicy<PROB, 1, true, METH>>(a, st)
                           : launch_thread<ThreadPolicy<PROB, 1, false, METH>>(a, st);
    case 2: return m.savej ? launch_thread<ThreadPolicy<PROB, 2, true, METH>>(a, st)
                           : launch_thread<ThreadPolicy<PROB, 2, false, METH>>(a, st);
    case 3: return launch_thread<ThreadPolicy<PROB, 3, false, METH>>(a, st);
    default: return -1000;
    }
}
template <class PROB>
static int thread_dispatch(int mf, const KArgs &a, cudaStream_t st) {
    const MfParts m = mf_parts(mf);
    if (m.meth == 1) return thread_dispatch_m<PROB, 1>(m, a, st);
    if (m.meth == 2) return thread_dispatch_m<PROB, 2>(m, a, st);
    return -1000;
}
template <class PROB>
static int thread_dky_dispatch(int mf, const DkyArgs &a, cudaStream_t st) {
    // the state layout depends on meth (lmax) and on SAVEJ through the Lin block only
    const MfParts m = mf_parts(mf);
    if (m.meth == 1)
        return m.savej ? launch_dky_thread<ThreadPolicy<PROB, 1, true, 1>>(a, st)
                       : launch_dky_thread<ThreadPolicy<PROB, 1, false, 1>>(a, st);
    return m.savej ? launch_dky_thread<ThreadPolicy<PROB, 1, true, 2>>(a, st)
                   : launch_dky_thread<ThreadPolicy<PROB, 1, false, 2>>(a, st);
}
template <class PROB>
static void thread_state_size(int mf, int ml, int mu, size_t nsys, size_t *nd, size_t *ni) {
    (void)ml; (void)mu;
    constexpr int N = PROB::N;
    const MfParts m = mf_parts(mf);
    const int lm = kLmaxOf(m.meth);
    *nd = nsys * (size_t)(lm * N + 2 * N + N * N + (m.savej ? N * N : 0) + lm + 14);
    *ni = nsys * (size_t)(N + 19);
}
// Copy `rows` SoA rows of `width` entries each (D2D, async).
template <class T>
static void copy_rows(T *dst, const T *src, size_t rows, size_t width, cudaStream_t st) {
    if (rows) cudaMemcpyAsync(dst, src, sizeof(T) * rows * width, cudaMemcpyDeviceToDevice, st);
}
// istate = 3 with another method flag: carry the method-independent part of the state over into the layout of
// mf_new (zero-filled by the caller).  Within a family only miter / jsv change; across families (Adams <-> BDF)
// the Nordsieck array and the tau history have another number of columns (lmax = 13 / 6): the common columns are
// copied (a Nordsieck array means the same under both methods, M:2031-2079 goes on from it), and when the array
// shrinks its old column maxord_new + 2 -- which the order reduction of M:1385 / 2044-2079 reads -- goes into the
// acor slot (free at a restart: dvnlsd zeroes acor before its next use).
template <class PROB>
static int thread_state_convert(int mf_old, int mf_new, int, int, size_t nsys, const real *d_old,
                                const int *i_old, real *d_new, int *i_new, cudaStream_t st) {
    constexpr size_t N = PROB::N;
    const MfParts a = mf_parts(mf_old), b = mf_parts(mf_new);
    const size_t la = kLmaxOf(a.meth), lb = kLmaxOf(b.meth), lmin = la < lb ? la : lb;
    // doubles: yh[lm][N], acor copy[N], acor[N], P[N][N] | saved J if jsv > 0 | tau[lm], 14 scalars; see thread_state_size
    const size_t acs_a = la * N, acs_b = lb * N;
    const size_t tau_a = acs_a + 2 * N + N * N + (a.savej ? N * N : 0), tau_b = acs_b + 2 * N + N * N + (b.savej ? N * N : 0);
    copy_rows(d_new, d_old, lmin * N, nsys, st);                                              // yh
    copy_rows(d_new + acs_b * nsys, d_old + acs_a * nsys, N, nsys, st);                       // acor copy
    if (la > lb) copy_rows(d_new + (acs_b + N) * nsys, d_old + lb * N * nsys, N, nsys, st);    // acor slot <- old yh(:, lmax_new + 1)
    else copy_rows(d_new + (acs_b + N) * nsys, d_old + (acs_a + N) * nsys, N, nsys, st);       // acor
    copy_rows(d_new + tau_b * nsys, d_old + tau_a * nsys, lmin, nsys, st);                    // tau
    copy_rows(d_new + (tau_b + lb) * nsys, d_old + (tau_a + la) * nsys, 14, nsys, st);         // scalars
    copy_rows(i_new, i_old, N + 19, nsys, st);
    return (int)cudaGetLastError();
}
template <class PROB>
static int thread_supports(int mf, int ml, int mu) {
    (void)ml; (void)mu;
    const MfParts m = mf_parts(mf);
    return (m.meth == 1 || m.meth == 2) && m.miter >= 0 && m.miter <= 3;
}
#define BVODE_THREAD_PROBLEM(NAME, PROB)                                                    \
    ProblemEntry {                                                                          \
        NAME, PROB::N, PROB::NPAR, 0, thread_supports<PROB>, thread_state_size<PROB>,       \
            thread_dispatch<PROB>, thread_dky_dispatch<PROB>, thread_state_convert<PROB>    \
    }

#if !BVODE_REAL_IS_FLOAT
template <class P, class = void>
struct HasJacRowMember { static constexpr bool value = false; };
template <class P>
struct HasJacRowMember<P, decltype((void)&P::jac_row)> { static constexpr bool value = true; };
template <class P>
struct HasJacRow { static constexpr bool value = HasJacRowMember<P>::value || HasJacDenseFull<P>::value; };
template <class P>
constexpr bool kHasBandJac = HasJacBandCol<P>::value || HasJacBandFull<P>::value;

// ---- warp-per-system problems ---------------------------------------------------------------
// KIND 1: dense finite-difference Jacobian available (miter 2); KIND 2: banded (miter 4, 5).
// Both kinds also run miter 0 (functional iteration) and 3 (diagonal approximation).
template <class PROB, int KIND, int METH>
static int warp_dispatch_m(const MfParts &m, const KArgs &a, cudaStream_t st) {
    switch (m.miter) {
    case 0: return launch_warp<WarpPlainPolicy<PROB, 0, METH>>(a, st);
    case 3: return launch_warp<WarpPlainPolicy<PROB, 3, METH>>(a, st);
    case 1:  // dense analytic Jacobian: problems with a jac_row (or whole-matrix jac_dense_full) member
        if constexpr (KIND == 1 && HasJacRow<PROB>::value)
            return m.savej ? launch_warp<WarpDensePolicy<PROB, 1, true, METH>>(a, st)
                           : launch_warp<WarpDensePolicy<PROB, 1, false, METH>>(a, st);
        return -1000;
    case 2:
        if constexpr (KIND == 1)
            return m.savej ? launch_warp<WarpDensePolicy<PROB, 2, true, METH>>(a, st)
                           : launch_warp<WarpDensePolicy<PROB, 2, false, METH>>(a, st);
        return -1000;
    case 4:  // banded analytic Jacobian: problems with jac_band_col (or whole-matrix jac_band_full)
        if constexpr (KIND == 2 && kHasBandJac<PROB>)
            return m.savej ? launch_warp<WarpBandPolicy<PROB, 4, true, METH>>(a, st)
                           : launch_warp<WarpBandPolicy<PROB, 4, false, METH>>(a, st);
        return -1000;
    case 5:
        if constexpr (KIND == 2)
            return m.savej ? launch_warp<WarpBandPolicy<PROB, 5, true, METH>>(a, st)
                           : launch_warp<WarpBandPolicy<PROB, 5, false, METH>>(a, st);
        return -1000;
    default: return -1000;
    }
}
template <class PROB, int KIND>
static int warp_dispatch(int mf, const KArgs &a, cudaStream_t st) {
    const MfParts m = mf_parts(mf);
    if (m.meth == 1) return warp_dispatch_m<PROB, KIND, 1>(m, a, st);
    if (m.meth == 2) return warp_dispatch_m<PROB, KIND, 2>(m, a, st);
    return -1000;
}
// F is a generic functor called with a policy tag whose state layout matches this mf
template <class PROB, int KIND, int METH, class F>
static int warp_layout_m(const MfParts &m, F &&fn) {
    if (m.miter == 0 || m.miter == 3) return fn(WarpPlainPolicy<PROB, 0, METH>{});
    if constexpr (KIND == 1) {
        if (m.miter == 2 || (m.miter == 1 && HasJacRow<PROB>::value))  // one state layout for both
            return m.savej ? fn(WarpDensePolicy<PROB, 2, true, METH>{})
                           : fn(WarpDensePolicy<PROB, 2, false, METH>{});
    } else {
        if (m.miter == 4 || m.miter == 5)
            return m.savej ? fn(WarpBandPolicy<PROB, 4, true, METH>{})
                           : fn(WarpBandPolicy<PROB, 4, false, METH>{});
    }
    return -1000;
}
template <class PROB, int KIND, class F>
static int warp_layout(int mf, F &&fn) {
    const MfParts m = mf_parts(mf);
    if (m.meth == 1) return warp_layout_m<PROB, KIND, 1>(m, fn);
    if (m.meth == 2) return warp_layout_m<PROB, KIND, 2>(m, fn);
    return -1000;
}
template <class PROB, int KIND>
static int warp_dky_dispatch(int mf, const DkyArgs &a, cudaStream_t st) {
    return warp_layout<PROB, KIND>(mf, [&](auto pol) { return launch_dky_warp<decltype(pol)>(a, st); });
}
template <class PROB, int KIND>
static void warp_state_size(int mf, int ml, int mu, size_t nsys, size_t *nd, size_t *ni) {
    *nd = *ni = 0;
    warp_layout<PROB, KIND>(mf, [&](auto pol) {
        WarpLayout<decltype(pol)>::sizes(ml, mu, nsys, nd, ni);
        return 0;
    });
}
// The same for the one-system-per-warp layouts: scalars[lm + 14][nsys], then the lane slots yh[lm][NLOC], acor
// copy[NLOC], acor[NLOC] (rows of nsys * 32); everything else (matrix images, pivots) is rebuilt at the restart.
template <int NLOC>
static int warp_convert_generic(int mf_old, int mf_new, size_t nsys, const real *d_old, const int *i_old,
                                real *d_new, int *i_new, cudaStream_t st) {
    const MfParts a = mf_parts(mf_old), b = mf_parts(mf_new);
    const size_t la = kLmaxOf(a.meth), lb = kLmaxOf(b.meth), lmin = la < lb ? la : lb;
    copy_rows(d_new, d_old, lmin, nsys, st);                                     // tau
    copy_rows(d_new + lb * nsys, d_old + la * nsys, 14, nsys, st);               // scalars
    const real *lo = d_old + (la + 14) * nsys;
    real *ln = d_new + (lb + 14) * nsys;
    const size_t w = nsys * 32;
    copy_rows(ln, lo, lmin * NLOC, w, st);                                       // yh
    copy_rows(ln + lb * NLOC * w, lo + la * NLOC * w, NLOC, w, st);              // acor copy
    if (la > lb) copy_rows(ln + (lb + 1) * NLOC * w, lo + lb * NLOC * w, NLOC, w, st);  // acor slot <- old yh(:, lmax_new + 1)
    else copy_rows(ln + (lb + 1) * NLOC * w, lo + (la + 1) * NLOC * w, NLOC, w, st);    // acor
    copy_rows(i_new, i_old, (size_t)kScalarI, nsys, st);
    return (int)cudaGetLastError();
}
template <class PROB, int KIND>
static int warp_state_convert(int mf_old, int mf_new, int, int, size_t nsys, const real *d_old,
                              const int *i_old, real *d_new, int *i_new, cudaStream_t st) {
    return warp_convert_generic<1>(mf_old, mf_new, nsys, d_old, i_old, d_new, i_new, st);
}
template <class PROB, int KIND>
static int warp_supports(int mf, int ml, int mu) {
    const MfParts m = mf_parts(mf);
    if (m.meth != 1 && m.meth != 2) return 0;
    if (m.miter == 0 || m.miter == 3) return 1;
    if (KIND == 1) return m.miter == 2 || (m.miter == 1 && HasJacRow<PROB>::value);
    if (m.miter == 4 && !kHasBandJac<PROB>) return 0;
    return (m.miter == 4 || m.miter == 5) && ml >= 0 && mu >= 0 && ml + mu + 1 <= 32 &&
           ml < PROB::N && mu < PROB::N;
}
// ---- one system per warp, two components per lane (32 < n <= 64): dense FD Jacobian only so far
template <class PROB>
static int warp2_dispatch(int mf, const KArgs &a, cudaStream_t st) {
    const MfParts m = mf_parts(mf);
    if (m.meth == 1) {
        if (m.miter == 0) return launch_warp<WarpPlain2Policy<PROB, 0, 1>>(a, st);
        if (m.miter == 3) return launch_warp<WarpPlain2Policy<PROB, 3, 1>>(a, st);
        if (m.miter == 2)
            return m.savej ? launch_warp<WarpDense2Policy<PROB, 2, true, 1>>(a, st) : launch_warp<WarpDense2Policy<PROB, 2, false, 1>>(a, st);
    }
    if (m.meth == 2) {
        if (m.miter == 0) return launch_warp<WarpPlain2Policy<PROB, 0, 2>>(a, st);
        if (m.miter == 3) return launch_warp<WarpPlain2Policy<PROB, 3, 2>>(a, st);
        if (m.miter == 2)
            return m.savej ? launch_warp<WarpDense2Policy<PROB, 2, true, 2>>(a, st) : launch_warp<WarpDense2Policy<PROB, 2, false, 2>>(a, st);
    }
    return -1000;
}
template <class PROB, class F>
static int warp2_layout(int mf, F &&fn) {
    const MfParts m = mf_parts(mf);
    if (m.meth != 1 && m.meth != 2) return -1000;
    if (m.miter == 0 || m.miter == 3)  // one state layout for both
        return (m.meth == 1) ? fn(WarpPlain2Policy<PROB, 0, 1>{}) : fn(WarpPlain2Policy<PROB, 0, 2>{});
    if (m.miter != 2) return -1000;
    if (m.meth == 1) return m.savej ? fn(WarpDense2Policy<PROB, 2, true, 1>{}) : fn(WarpDense2Policy<PROB, 2, false, 1>{});
    return m.savej ? fn(WarpDense2Policy<PROB, 2, true, 2>{}) : fn(WarpDense2Policy<PROB, 2, false, 2>{});
}
template <class PROB>
static int warp2_dky_dispatch(int mf, const DkyArgs &a, cudaStream_t st) {
    return warp2_layout<PROB>(mf, [&](auto pol) { return launch_dky_warp<decltype(pol)>(a, st); });
}
template <class PROB>
static void warp2_state_size(int mf, int ml, int mu, size_t nsys, size_t *nd, size_t *ni) {
    *nd = *ni = 0;
    warp2_layout<PROB>(mf, [&](auto pol) {
        WarpLayout<decltype(pol)>::sizes(ml, mu, nsys, nd, ni);
        return 0;
    });
}
template <class PROB>
static int warp2_supports(int mf, int, int) {
    const MfParts m = mf_parts(mf);
    return (m.meth == 1 || m.meth == 2) && (m.miter == 0 || m.miter == 2 || m.miter == 3);
}
static int warp2_state_convert(int mf_old, int mf_new, int, int, size_t nsys, const real *d_old, const int *i_old,
                               real *d_new, int *i_new, cudaStream_t st) {
    return warp_convert_generic<2>(mf_old, mf_new, nsys, d_old, i_old, d_new, i_new, st);
}
#define BVODE_WARP2_DENSE_FD_PROBLEM(NAME, PROB)                                            \
    ProblemEntry {                                                                          \
        NAME, PROB::N, PROB::NPAR, 1, warp2_supports<PROB>, warp2_state_size<PROB>,         \
            warp2_dispatch<PROB>, warp2_dky_dispatch<PROB>, warp2_state_convert             \
    }
#define BVODE_WARP_DENSE_FD_PROBLEM(NAME, PROB)                                             \
    ProblemEntry {                                                                          \
        NAME, PROB::N, PROB::NPAR, 1, warp_supports<PROB, 1>, warp_state_size<PROB, 1>,     \
            warp_dispatch<PROB, 1>, warp_dky_dispatch<PROB, 1>, warp_state_convert<PROB, 1> \
    }
#define BVODE_WARP_BAND_PROBLEM(NAME, PROB)                                                 \
    ProblemEntry {                                                                          \
        NAME, PROB::N, PROB::NPAR, 1, warp_supports<PROB, 2>, warp_state_size<PROB, 2>,     \
            warp_dispatch<PROB, 2>, warp_dky_dispatch<PROB, 2>, warp_state_convert<PROB, 2> \
    }

#endif  // !BVODE_REAL_IS_FLOAT

}  // namespace BVODE_NS
