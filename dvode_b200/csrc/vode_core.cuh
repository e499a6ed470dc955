// vode_core.cuh -- the batched VODE integrator core (device code).
//
// One Engine<POL> instance integrates ONE system.  The policy POL decides how the
// system is mapped onto the machine:
//   ThreadPolicy (policy_thread.cuh): one system per thread, every vector in registers
//                                     (tiny systems: Robertson n = 3).
//   WarpPolicy   (policy_warp.cuh)  : one system per warp, component i on lane i,
//                                     scalars replicated on all lanes (n <= 32).
// The scalar control path (step size / order selection, coefficient generation,
// corrector logic) is written once, here, and is identical for both mappings.
//
// What is restated here, from /root/reference/src/dvode_module.F90 (M):
//   driver blocks C, D (itask = 1), E, F, G      M:1304-1381, 1399-1468, 1496-1623, 1664-1723
//   dvhin  M:1738-1851      dvindy (k = 0)  M:1878-1952     dvstep  M:1975-2377
//   dvset  (BDF) M:2440-2490, 2559          dvjust (BDF)    M:2581-2651
//   dvnlsd (chord) M:2703-2879              dewset/dvnorm   M:3238-3311
// dvjac / dvsol and the LINPACK factor/solve live in the policies.
//
// Design differences from the reference (results are identical operation for operation):
//  * The Nordsieck array keeps the invariant "columns above L are exactly zero"; the
//    acor copy that the reference parks in yh(:,lmax) (M:2259, 2298) lives in its own
//    vector `acsav`.  With that invariant the Pascal-triangle predict / retract, the
//    rescale, the accept update and dvindy run at full order with no dependence on nq,
//    so threads of a warp at different orders execute the same instruction stream
//    (x + 0.0 == x, so the extra operations change nothing).
//  * dvstep's GOTO retry network (labels 300/400) is one loop whose iteration is one
//    step attempt; `entry` says where the attempt starts.  All lanes reconverge at the
//    loop head, and a lane whose system finished simply drops out.
#pragma once
#include "bvode_internal.h"
#include "bvode_math.cuh"

namespace BVODE_NS {

// mord = [12, 5] (M:38): maximum order of the Adams (meth = 1) and BDF (meth = 2) families.
// Every array that the reference dimensions by maxord is sized at compile time from the
// method: lmax = 13 (Adams) or 6 (BDF).
constexpr int kMaxOrdBdf = 5;
__host__ __device__ constexpr int kMaxOrdOf(int meth) { return meth == 1 ? 12 : 5; }
__host__ __device__ constexpr int kLmaxOf(int meth) { return kMaxOrdOf(meth) + 1; }
constexpr int kLmaxBdf = kMaxOrdBdf + 1;

// istate / error codes written per system (same integers as the reference, M:832-908)
enum : int {
    IST_FIRST = 1, IST_OK = 2,
    IST_MXSTEP = -1, IST_TOLSF = -2, IST_ILLEGAL = -3, IST_ERRTEST = -4, IST_CONVFAIL = -5,
    IST_EWT = -6,
    IST_ILLEGAL_ENTRY = -30  // internal: istate = -3 detected at call entry (y, t untouched)
};

// Scalar per-system state: the members of dvode_data_t (M:40-153) that are not
// batch-uniform.  el/tau/tq are 1-based like the reference (element 0 unused).
// Hot scalars (touched several times per step attempt) stay in this struct, i.e. in registers.
// The rest of dvode_data_t -- tau(1:6), el(1:6), tq(1:5), conp, hu, hnew, crate, etamax and
// the counters -- is reached through the policy's `Hot` storage (shared memory for the
// thread-per-system policy, registers for the warp policies) by the slot ids below.
struct Ctl {
    real h, hscal, eta, tn, rc, prl1, rl1, drc, acnrm;
    int nq, L, nqwait, newq, newh, jstart, icf, ipup, jcur, kflag;
};
// SD_ETAQ / SD_ETAQM1: the reference's etaq / etaqm1 (M:258-259), which only an istate = 3 call
// that lowers maxord below a pending order increase reads back (M:2059-2063).  Strict build:
// the values; fast build: the arguments xq = (bias2*dsm)^2, xm = (bias1*ddn)^2 of the step that
// chose the increase (the ratios are evaluated from them if they are ever needed).
enum : int { SD_CONP = 0, SD_HU, SD_HNEW, SD_CRATE, SD_ETAMAX, SD_ETAQ, SD_ETAQM1, SD_COUNT };
enum : int { SI_NST = 0, SI_NFE, SI_NJE, SI_NLU, SI_NNI, SI_NCFN, SI_NETF, SI_NQU, SI_NSLJ,
             SI_NSLP, SI_NHNIL, SI_INIT, SI_COUNT };

using Opts = ::BvOpts;  // batch-uniform optional inputs (M:729-751)

template <int LO, int HI>
BV_DEV real pick(const real *a, int idx) {
    real r = a[LO];
#pragma unroll
    for (int i = LO + 1; i <= HI; i++) r = (idx == i) ? a[i] : r;
    return r;
}

template <int LO, int HI, class HOT>
BV_DEV real pick_tau(HOT &hot, int idx) {
    real r = hot.tau(LO);
#pragma unroll
    for (int i = LO + 1; i <= HI; i++) r = (idx == i) ? hot.tau(i) : r;
    return r;
}

// ---------------------------------------------------------------- dvset (BDF) ----
// M:2440-2490, 2559.  Loops are unrolled to the compile-time maximum order and guarded
// by nq.  el / tq are built in local registers and committed to the policy's storage once at
// the end (a policy may keep them in shared memory: they are produced here and consumed after
// the corrector, and registers across the corrector are what limits occupancy).
template <class HOT>
BV_DEV void dvset_bdf(Ctl &c, HOT &hot) {
    const int nq = c.nq, L = c.L;
    const real h = c.h;
    real el[kLmaxBdf + 1], tq[6];
#pragma unroll
    for (int i = 3; i <= kLmaxBdf; i++) el[i] = BV_R(0.0);
    el[1] = BV_R(1.0);
    el[2] = BV_R(1.0);
    real tauv[kLmaxBdf + 1];
#pragma unroll
    for (int i = 1; i <= kLmaxBdf; i++) tauv[i] = hot.tau(i);
    real alph0 = -BV_R(1.0), ahatn0 = -BV_R(1.0), hsum = h, rxi = BV_R(1.0), rxis = BV_R(1.0);
    if (nq != 1) {
#pragma unroll
        for (int j = 1; j <= kMaxOrdBdf - 2; j++) {
            if (j <= nq - 2) {
                hsum = hsum + tauv[j];
                rxi = qdiv(h, hsum);
                alph0 = alph0 - BV_R(1.0) / (real)(j + 1);
#pragma unroll
                for (int iback = 1; iback <= j + 1; iback++) {
                    const int i = (j + 3) - iback;
                    el[i] = el[i] + el[i - 1] * rxi;
                }
            }
        }
        alph0 = alph0 - recip_int(nq);
        rxis = -el[2] - alph0;
        hsum = hsum + pick<1, kMaxOrdBdf>(tauv, nq - 1);
        rxi = qdiv(h, hsum);
        ahatn0 = -el[2] - rxi;
#pragma unroll
        for (int i = kLmaxBdf; i >= 2; i--) {
            if (i <= nq + 1) el[i] = el[i] + el[i - 1] * rxis;
        }
    }
    const real t1 = BV_R(1.0) - ahatn0 + alph0;
    const real t2 = BV_R(1.0) + (real)nq * t1;
    const real elL = pick<2, kLmaxBdf>(el, L);
    tq[2] = fabs(qdiv(alph0 * t2, t1));
    tq[1] = BV_R(0.0);
    tq[3] = BV_R(0.0);
#if BVODE_STRICT
    tq[5] = fabs(t2 / (elL * rxi / rxis));
    if (c.nqwait == 1) {
        const real cnqm1 = rxis / elL;
        const real t3 = alph0 + recip_int(nq);
        const real t4 = ahatn0 + rxi;
        real elp = t3 / (BV_R(1.0) - t4 + t3);
        tq[1] = fabs(elp / cnqm1);
        hsum = hsum + pick<1, kLmaxBdf>(tauv, nq);
        rxi = h / hsum;
        const real t5 = alph0 - recip_int(nq + 1);
        const real t6 = ahatn0 - rxi;
        elp = t2 / (BV_R(1.0) - t6 + t5);
        tq[3] = fabs(elp * rxi * ((real)L + BV_R(1.0)) * t5);
    }
#else
    // fast build: the nested quotients of M:2478-2489 as single quotients
    //   t2 / (elL*rxi/rxis)             = t2*rxis / (elL*rxi)
    //   (t3/(1-t4+t3)) / (rxis/elL)     = t3*elL / ((1-t4+t3)*rxis)
    //   (t2/(1-t6+t5)) * rxi*(L+1)*t5   = t2*rxi*(L+1)*t5 / (1-t6+t5)
    tq[5] = fabs(qdiv(t2 * rxis, elL * rxi));
    if (c.nqwait == 1) {
        const real t3 = alph0 + recip_int(nq);
        const real t4 = ahatn0 + rxi;
        tq[1] = fabs(qdiv(t3 * elL, (BV_R(1.0) - t4 + t3) * rxis));
        hsum = hsum + pick<1, kLmaxBdf>(tauv, nq);
        rxi = qdiv(h, hsum);
        const real t5 = alph0 - recip_int(nq + 1);
        const real t6 = ahatn0 - rxi;
        tq[3] = fabs(qdiv(t2 * rxi * ((real)L + BV_R(1.0)) * t5, BV_R(1.0) - t6 + t5));
    }
#endif
    tq[4] = BV_R(0.1) * tq[2];
#pragma unroll
    for (int i = 1; i <= kLmaxBdf; i++) hot.el(i) = el[i];
    // tq(1), tq(3) are only read by the order selection of the step whose dvset ran with
    // nqwait == 1 (M:2272-2292), so they need no storage otherwise
    hot.tq(2) = tq[2];
    hot.tq(4) = tq[4];
    hot.tq(5) = tq[5];
    if (c.nqwait == 1) {
        hot.tq(1) = tq[1];
        hot.tq(3) = tq[3];
    }
}

// dvjust coefficient part for both directions (the yh updates are done by the engine):
// builds el(3:) from the tau history, M:2589-2616 (iord = +1) and M:2629-2643 (iord = -1).
// The two branches of the reference share the recurrence el(i) = el(i)*x + el(i-1); they are
// one code path here so that lanes raising and lowering their order execute together.
// el is returned in a local array (the reference leaves it in el(:), which dvset overwrites
// before its next use).  Returns t1 = (-alph0 - alph1)/prod for the order increase.
template <class HOT>
BV_DEV real dvjust_coeffs_bdf(Ctl &c, HOT &hot, bool up, real (&el)[kLmaxBdf + 1]) {
    const int nq = c.nq;
#pragma unroll
    for (int j = 0; j <= kLmaxBdf; j++) el[j] = BV_R(0.0);
    el[3] = BV_R(1.0);
    real alph0 = -BV_R(1.0), alph1 = BV_R(1.0), prod = BV_R(1.0), xiold = BV_R(1.0);
    real hsum = up ? c.hscal : BV_R(0.0);
    const int jmax = up ? nq - 1 : nq - 2;
#if !BVODE_STRICT
    const real rhscal = qrcp(c.hscal);
#endif
    real tauv[kLmaxBdf + 1];
#pragma unroll
    for (int i = 1; i <= kLmaxBdf; i++) tauv[i] = hot.tau(i);
#pragma unroll
    for (int j = 1; j <= kMaxOrdBdf - 2; j++) {
        if (j <= jmax) {
            hsum = hsum + (up ? tauv[j + 1] : tauv[j]);
#if BVODE_STRICT
            const real xi = hsum / c.hscal;
            const real rxi = up ? BV_R(1.0) / xi : BV_R(0.0);
#else
            const real xi = hsum * rhscal;  // one reciprocal of hscal, and 1/xi = hscal/hsum
            const real rxi = up ? c.hscal * qrcp(hsum) : BV_R(0.0);
#endif
            if (up) {
                prod = prod * xi;
                alph0 = alph0 - BV_R(1.0) / (real)(j + 1);
                alph1 = alph1 + rxi;
            }
            const real x = up ? xiold : xi;
#pragma unroll
            for (int iback = 1; iback <= j + 1; iback++) {
                const int i = (j + 4) - iback;
                el[i] = el[i] * x + el[i - 1];
            }
            xiold = xi;
        }
    }
    return qdiv(-alph0 - alph1, prod);
}

// dvjust(-1), BDF branch (M:2629-2643), for an order above the BDF maximum: reached once, when an istate = 3
// call switches an Adams history at order 6 to BDF (maxord = nq - 1, M:2059-2068).  Run-time loops, out of
// line.  el(3:nq) such that yh(:,j) -= yh(:,L)*el(j).
template <class HOT>
static __device__ __noinline__ void dvjust_coeffs_bdf_down_wide(Ctl &c, HOT &hot, real (&el)[kLmaxBdf + 1]) {
    const int nq = c.nq;  // == kMaxOrdBdf + 1
    real w[kLmaxOf(1) + 2];
    for (int j = 0; j < kLmaxOf(1) + 2; j++) w[j] = BV_R(0.0);
    w[3] = BV_R(1.0);
    real hsum = BV_R(0.0);
    for (int j = 1; j <= nq - 2; j++) {
        hsum = hsum + hot.tau(j);
        const real xi = hsum / c.hscal;
        const int jp1 = j + 1;
        for (int iback = 1; iback <= jp1; iback++) {
            const int i = (j + 4) - iback;
            w[i] = w[i] * xi + w[i - 1];
        }
    }
    for (int j = 0; j <= kLmaxBdf; j++) el[j] = w[j];
}

// ---------------------------------------------------------------- dvset (Adams) ----
// M:2492-2558.  Run-time loops over the order (up to 12): the Adams family is a coverage
// path, not the tuned one, so el / em live in local arrays indexed at run time.
template <class HOT>
BV_DEV void dvset_adams(Ctl &c, HOT &hot) {
    constexpr int LM = kLmaxOf(1);
    const int nq = c.nq, L = c.L, nqm1 = nq - 1;
    const real h = c.h;
    real el[LM + 1], em[LM + 2], tq[6], tauv[LM + 1];
    for (int i = 1; i <= LM; i++) { tauv[i] = hot.tau(i); el[i] = BV_R(0.0); }  // el above L: zero (see accept update)
    tq[1] = BV_R(0.0); tq[2] = BV_R(0.0); tq[3] = BV_R(0.0); tq[5] = BV_R(0.0);
    const real flotl = (real)L;
    if (nq != 1) {
        real hsum = h;
        em[1] = BV_R(1.0);
        const real flotnq = flotl - BV_R(1.0);
        for (int i = 2; i <= L; i++) em[i] = BV_R(0.0);
        for (int j = 1; j <= nqm1; j++) {
            if ((j == nqm1) && (c.nqwait == 1)) {
                real s = BV_R(1.0), csum = BV_R(0.0);
                for (int i = 1; i <= nqm1; i++) {
                    csum = csum + s * em[i] / (real)(i + 1);
                    s = -s;
                }
                tq[1] = em[nqm1] / (flotnq * csum);
            }
            const real rxi = h / hsum;
            for (int iback = 1; iback <= j; iback++) {
                const int i = (j + 2) - iback;
                em[i] = em[i] + em[i - 1] * rxi;
            }
            hsum = hsum + tauv[j];
        }
        real s = BV_R(1.0), em0 = BV_R(0.0), csum = BV_R(0.0);
        for (int i = 1; i <= nq; i++) {
            const real floti = (real)i;
            em0 = em0 + s * em[i] / floti;
            csum = csum + s * em[i] / (floti + BV_R(1.0));
            s = -s;
        }
        s = BV_R(1.0) / em0;
        el[1] = BV_R(1.0);
        for (int i = 1; i <= nq; i++) el[i + 1] = s * em[i] / (real)i;
        const real xi = hsum / h;
        tq[2] = xi * em0 / csum;
        tq[5] = xi / el[L];
        if (c.nqwait == 1) {
            const real rxi = BV_R(1.0) / xi;
            for (int iback = 1; iback <= nq; iback++) {
                const int i = (L + 1) - iback;
                em[i] = em[i] + em[i - 1] * rxi;
            }
            s = BV_R(1.0);
            csum = BV_R(0.0);
            for (int i = 1; i <= L; i++) {
                csum = csum + s * em[i] / (real)(i + 1);
                s = -s;
            }
            tq[3] = flotl * em0 / csum;
        }
    } else {
        el[1] = BV_R(1.0);
        el[2] = BV_R(1.0);
        tq[1] = BV_R(1.0);
        tq[2] = BV_R(2.0);
        tq[3] = BV_R(6.0) * tq[2];
        tq[5] = BV_R(1.0);
    }
    tq[4] = BV_R(0.1) * tq[2];
    for (int i = 1; i <= LM; i++) hot.el(i) = el[i];
    hot.tq(2) = tq[2];
    hot.tq(4) = tq[4];
    hot.tq(5) = tq[5];
    if (c.nqwait == 1 || nq == 1) {
        hot.tq(1) = tq[1];
        hot.tq(3) = tq[3];
    }
}

// dvjust, Adams order decrease (M:2669-2692): el(3:nq) such that yh(:,j) -= yh(:,L)*el(j).
// (The Adams order increase only zeroes the new column, M:2661-2666.)
template <class HOT>
BV_DEV void dvjust_coeffs_adams_down(Ctl &c, HOT &hot, real (&el)[kLmaxOf(1) + 1]) {
    constexpr int LM = kLmaxOf(1);
    const int nq = c.nq;
    for (int j = 0; j <= LM; j++) el[j] = BV_R(0.0);
    el[2] = BV_R(1.0);
    real hsum = BV_R(0.0);
    for (int j = 1; j <= nq - 2; j++) {
        hsum = hsum + hot.tau(j);
        const real xi = hsum / c.hscal;
        for (int iback = 1; iback <= j + 1; iback++) {
            const int i = (j + 3) - iback;
            el[i] = el[i] * xi + el[i - 1];
        }
    }
    for (int j = 2; j <= nq - 1; j++) el[j + 1] = (real)nq * el[j] / (real)j;
}

template <int METH, class HOT>
BV_DEV void dvset(Ctl &c, HOT &hot) {
    if constexpr (METH == 2) dvset_bdf(c, hot);
    else dvset_adams(c, hot);
}

#if !BVODE_STRICT
// Which step-ratio candidate wins (M:2294-2317), from log-ratios l = log2(x)/k (smaller l =
// larger eta): 0 = current order, 1 = order - 1, 2 = order + 1.
template <class T>
BV_DEV int choose_ratio(T lq, T lm, T lp) {
    if (lq > lp) return (lp >= lm) ? 1 : 2;
    return (lq > lm) ? 1 : 0;
}
// The same decision from double-precision logs: the rare case where the single-precision
// screen is too close to call.  Out of line to keep the step loop small.
static __device__ __noinline__ int choose_ratio_f64(real xq, real xm, real xp, real scq, real scm,
                                             real scp, bool has_m, bool has_p) {
    const real inf = __longlong_as_double(0x7ff0000000000000LL);
    const real lq = log2_fast(xq) * scq;
    const real lm = has_m ? log2_fast(xm) * scm : inf;
    const real lp = has_p ? log2_fast(xp) * scp : inf;
    return choose_ratio(lq, lm, lp);
}
#endif

// Detection of a problem's optional members `ewt` / `norm`: the user-replaceable dewset /
// dvnorm of initialize (M:431-503).
template <class P, class = void>
struct HasUserEwt { static constexpr bool value = false; };
template <class P>
struct HasUserEwt<P, decltype((void)&P::ewt)> { static constexpr bool value = true; };
template <class P, class = void>
struct HasUserNorm { static constexpr bool value = false; };
template <class P>
struct HasUserNorm<P, decltype((void)&P::norm)> { static constexpr bool value = true; };

template <class P, class = void>
struct HasLaneNorm { static constexpr bool value = false; };  // the lane-wise form (norm_term / combine / final)
template <class P>
struct HasLaneNorm<P, decltype((void)&P::norm_term)> { static constexpr bool value = true; };

// =================================================================== Engine =====
// POL provides:
//   Prob, N (equations), NLOC (vector elements held by this thread), MITER
//   norm(v, w)                       weighted rms norm over the whole system (M:3295)
//   any_le_zero(v)                   true if any component of the system is <= 0
//   f(t, y, ydot, par)               the user's right-hand side
//   Lin                              linear-algebra state (P = LU, pivots, saved J)
//   AcSav                            storage of the acor copy (registers or HBM), get/set
//   Tol                              rtol(i) / atol(i) of this thread's components
//   jac_setup(eng, ierpj)            dvjac, M:2897-3104
//   solve(lin, x)                    dvsol, M:3139-3191
//   imxer(acor, ewt)                 index of the largest weighted error (M:1694-1706)
//
// Fast build only (BVODE_STRICT == 0), each mathematically equivalent to the reference
// expression and differing from it by rounding only (like FMA contraction does):
//   * the three step-ratio candidates etaq / etaqm1 / etaqp1 (M:2275-2292) are compared
//     through log(x)/k instead of 1/(x**(1/k) + addon) (same ordering: both are monotone),
//     and exp() is evaluated only for the winner when the ratio can exceed `thresh`;
//   * a/tq(2), b/tq(4) use one shared reciprocal of tq(2).
template <class POL>
struct Engine {
    using POLICY = POL;
    using Prob = typename POL::Prob;
    static constexpr int N = POL::N;
    static constexpr int NLOC = POL::NLOC;
    static constexpr int MITER = POL::MITER;   // 0 functional, 1/2 dense, 3 diagonal, 4/5 banded
    static constexpr int METH = POL::METH;     // 1 Adams, 2 BDF
    static constexpr int kMaxOrd = kMaxOrdOf(METH);
    static constexpr int kLmax = kLmaxOf(METH);

    Ctl c;
    typename POL::Hot hot;   // yh(j,i): Nordsieck history, column j = h^j y^(j)/j!; tau(1:6);
                             // (thread policy) the Newton matrix -- registers or shared memory
    real ewt[NLOC];        // stored inverted (M:1359, 1522)
    real savf[NLOC], acor[NLOC], y[NLOC];
    typename POL::Par par;   // per-system parameters of f / jac, as this thread holds them
    typename POL::AcSav acsav;  // the reference's yh(:,lmax) scratch use (M:2259, 2298)
    typename POL::Tol tol;
    typename POL::Lin lin;
    Opts o;
    // Equations actually integrated: N, or fewer after an istate = 3 call with a smaller neq (M:1127-1135; the
    // full-driver kernels only).  Components >= nact are frozen: f is masked to zero for them, their history is
    // cleared at the restart, they carry unit weight and zero error, and the norms divide by nact.
    int nact = N;
    BV_DEV bool active(int i) const { return POL::comp(i) < nact; }
    BV_DEV void feval(real t, const real (&yy)[NLOC], real (&out)[NLOC]) {  // the user's f for the first nact equations
        POL::f(t, yy, out, par);
        if (nact != N) {
#pragma unroll
            for (int i = 0; i < NLOC; i++) out[i] = active(i) ? out[i] : BV_R(0.0);
        }
    }

    // ---- small vector helpers --------------------------------------------------
    BV_DEV void dewset_invert(bool &bad) {  // M:3238-3273 then M:1351-1360 / 1516-1523
        real e[NLOC];
#pragma unroll
        for (int i = 0; i < NLOC; i++) e[i] = POL::ewt_of(i, hot.yh(0, i), tol.rtol(i), tol.atol(i));
        if (nact != N) {
#pragma unroll
            for (int i = 0; i < NLOC; i++) e[i] = active(i) ? e[i] : BV_R(1.0);
        }
        bad = POL::any_le_zero(e);
#pragma unroll
        for (int i = 0; i < NLOC; i++) ewt[i] = qrcp(e[i]);
    }

#if !BVODE_STRICT
    BV_DEV real msq_y0() {
        real y0v[NLOC];
#pragma unroll
        for (int i = 0; i < NLOC; i++) y0v[i] = active(i) ? hot.yh(0, i) : BV_R(0.0);
        return POL::msq(y0v, ewt, nact);
    }
#endif
    BV_DEV real norm_y0() {  // dvnorm(yh(:,1), ewt), M:1525
        real y0v[NLOC];
#pragma unroll
        for (int i = 0; i < NLOC; i++) y0v[i] = active(i) ? hot.yh(0, i) : BV_R(0.0);
        return POL::norm(y0v, ewt, nact);
    }

    BV_DEV void predict() {  // M:2148-2154 at full order (zero-column invariant)
#pragma unroll
        for (int jb = 1; jb <= kMaxOrd; jb++) {
#pragma unroll
            for (int j = kMaxOrd + 1 - jb; j <= kMaxOrd; j++) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = hot.yh(j - 1, i) + hot.yh(j, i);
            }
        }
    }
    BV_DEV void retract() {  // M:2181-2187, 2342-2348
#pragma unroll
        for (int jb = 1; jb <= kMaxOrd; jb++) {
#pragma unroll
            for (int j = kMaxOrd + 1 - jb; j <= kMaxOrd; j++) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = hot.yh(j - 1, i) - hot.yh(j, i);
            }
        }
    }
    // label 300, M:2131-2139.  Runs on every attempt: lanes that enter at label 400 pass
    // eta = 1, which changes nothing (x * 1.0 == x and h == hscal there), so the warp does
    // not split on "was h changed?".
    BV_DEV void rescale(real eta) {
        real r = BV_R(1.0);
#pragma unroll
        for (int j = 2; j <= kLmax; j++) {
            r = r * eta;
#pragma unroll
            for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = r * hot.yh(j - 1, i);
        }
        c.h = c.hscal * eta;
        c.hscal = c.h;
        c.rc = c.rc * eta;
    }
    // column `col` (1-based) of yh selected at run time
    BV_DEV void get_col(int col, real (&out)[NLOC]) {
#pragma unroll
        for (int i = 0; i < NLOC; i++) out[i] = hot.yh(0, i);
#pragma unroll
        for (int j = 2; j <= kLmax; j++) {
            if (col == j) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) out[i] = hot.yh(j - 1, i);
            }
        }
    }

    // istate = 3, M:1383-1390 (the optional inputs were re-read on the host; this is the device
    // part).  `first_comp`: this thread holds component 1 of the system.
    // Two consequences of the reference's work-array layout are mirrored / resolved here: with a
    // lowered maxord the segment pointer lwm = lyh + (maxord+1)*nyh moves INTO the old yh, and
    //  * M:1386 `rwork(lwm) = sqrt(uround)` (miter > 0) lands on yh(1, maxord+2) before dvstep's
    //    dvjust(-1) reads that column (case maxord = nq-1, M:2060-2068) -- mirrored, since the
    //    results depend on it;
    //  * a saved Jacobian (mf > 0) is then read from memory that holds other data -- resolved
    //    by re-evaluating it (o.jreset, set by the host when maxord changed).
    BV_DEV void restart_istate3(bool first_comp) {
        c.jstart = -1;
        xcol_mode = 0;
        if (nact != N) {  // neq was reduced (M:1127-1135, 2040-2049): the dropped equations take no further part
#pragma unroll
            for (int j = 1; j < kLmax; j++) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) hot.yh(j, i) = active(i) ? hot.yh(j, i) : BV_R(0.0);
            }
#pragma unroll
            for (int i = 0; i < NLOC; i++) {
                if (!active(i)) { acor[i] = BV_R(0.0); POL::acsav_set(*this, i, BV_R(0.0)); }
            }
        }
        if (c.nq > o.maxord) {
            // What dvjust(-1) will find in yh(:, maxord + 2) (case maxord = nq - 1, M:2060-2068) is decided by the
            // reference's work-array layout: the segments after yh start at lwm = lyh + (maxord+1)*nyh, i.e. ON that
            // column.  miter > 0: wm(1) = sqrt(uround) lands on its first component (M:1386); miter = 0: lenwm = 0,
            // so ewt itself (refreshed before the step, M:1516-1523) overlays the whole column.
            if (GENERAL_STASH && o.stash && o.maxord + 2 > kLmax) {
                // the method family changed (Adams -> BDF): yh(:, maxord + 2) of the old, wider array came along
                // in the acor slot (state_convert); dvjust(-1) reads it from savf
#pragma unroll
                for (int i = 0; i < NLOC; i++) savf[i] = acor[i];
                if (MITER > 0 && first_comp) xcol_mode = 1;
            } else {
                get_col(o.maxord + 2, savf);
                if (MITER > 0 && first_comp) {
#pragma unroll
                    for (int j = 2; j <= kLmax; j++) {
                        if (j == o.maxord + 2) hot.yh(j - 1, 0) = sqrt(kUround);
                    }
                }
            }
            if (MITER == 0) xcol_mode = 2;
        }
        if (nact != N) {
            // M:2045-2053: while n differs from the leading dimension of yh -- on every istate = 3 call from the
            // reduction on -- dvstep first clears yh(:, newq+2 : maxord+1), ALL components, "to avoid undefined
            // references".  Above the order these columns are zero here anyway; but with an order DECREASE pending
            // (newq = nq - 1) the cleared range starts at column nq + 1, the top Nordsieck column, which the
            // dvjust(-1) that follows (M:2114) would have folded into the lower ones, and it always takes in
            // yh(:, lmax), the acor copy of a pending order increase.  (Found on the one system in 40 of the n = 48
            // batch that restarts with an order decrease pending; the corrector then fails on the first step.)
            const bool top = (c.newq < c.nq) && (c.newq + 2 <= o.maxord + 1);
#pragma unroll
            for (int j = 2; j <= kLmax; j++) {
                if (top && j == c.nq + 1) {
#pragma unroll
                    for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = BV_R(0.0);
                }
            }
            if (c.newq + 2 <= o.maxord + 1) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) POL::acsav_set(*this, i, BV_R(0.0));
            }
        }
        if (o.jreset) hot.si(SI_NSLJ) = hot.si(SI_NST) - 51;  // forces jok = -1 (msbj = 50)
        // The acor copy of a pending order increase is parked in yh(:, lmax) (M:2259, 2298); with another lmax the
        // reference's dvjust(+1) / order test read yh(:, lmax_new) -- a Nordsieck column (zero above the order)
        // -- until the copy is saved again.  Mirrored for a column that exists in this array.
        if (o.lmaxchg && o.maxord + 1 <= kLmax) {
#pragma unroll
            for (int j = 1; j <= kLmax; j++) {
                if (j == o.maxord + 1) {
#pragma unroll
                    for (int i = 0; i < NLOC; i++) POL::acsav_set(*this, i, hot.yh(j - 1, i));
                }
            }
        }
    }
    static constexpr bool GENERAL_STASH = (METH == 2);  // only a BDF kernel can be entered from a wider array
    int xcol_mode = 0;  // until the restart's dvjust(-1) has run: 1 = component 1 of that column is sqrt(uround), 2 = the column is ewt

    // dvjust, M:2568-2694 (BDF), both directions in one pass over yh.
    //   down: yh(:,j) -= yh(:,L)*el(j), j = 3..nq   (nothing if nq == 2, M:2581); the freed
    //         column L is zeroed to keep the zero-column invariant.
    //   up:   yh(:,L+1) = t1*acsav ; yh(:,j) += el(j)*yh(:,L+1), j = 3..nq+1
    // WIDE (the istate = 3 restart only): an order above this family's maximum and the column overlays of the
    // reference's work array are possible; the step loop's own order changes never need them.
    template <bool WIDE = false>
    BV_DEV void order_change(bool up) {
        real el[kLmax + 1];
        real t1 = BV_R(0.0);
        if constexpr (METH == 2) {
            if (WIDE && !up && c.nq > kMaxOrdBdf) dvjust_coeffs_bdf_down_wide(c, hot, el);  // order 6, from an Adams history
            else t1 = dvjust_coeffs_bdf(c, hot, up, el);
        } else {
            // Adams (M:2656-2692): raising the order only zeroes the new column, which the
            // zero-column invariant already guarantees; lowering it needs el(3:nq)
            if (up) {
#pragma unroll
                for (int j = 0; j <= kLmax; j++) el[j] = BV_R(0.0);
            } else {
                dvjust_coeffs_adams_down(c, hot, el);
            }
        }
        real col[NLOC];
        if (up) {
#pragma unroll
            for (int i = 0; i < NLOC; i++) col[i] = (METH == 2) ? t1 * POL::acsav_get(*this, i) : BV_R(0.0);
        } else {
            if (WIDE && GENERAL_STASH && c.L > kLmax) {  // right after an Adams -> BDF restart at order 6 (see restart_istate3)
#pragma unroll
                for (int i = 0; i < NLOC; i++) col[i] = savf[i];
                if (xcol_mode == 1) col[0] = sqrt(kUround);
            } else {
                get_col(c.L, col);
            }
            if (WIDE && xcol_mode == 2) {  // miter = 0 at an istate = 3 restart: the column is overlaid by ewt
#pragma unroll
                for (int i = 0; i < NLOC; i++) col[i] = ewt[i];
            }
#pragma unroll
            for (int i = 0; i < NLOC; i++) col[i] = -col[i];
        }
        const int jhi = up ? ((METH == 2) ? c.nq + 1 : 0) : ((c.nq == 2) ? 0 : c.nq);
        const int jnew = up ? c.L + 1 : c.L;  // column written (up) or zeroed (down)
#pragma unroll
        for (int j = 3; j <= kLmax; j++) {
            if (j <= jhi) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = hot.yh(j - 1, i) + el[j] * col[i];
            }
            if (j == jnew) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = up ? col[i] : BV_R(0.0);
            }
        }
    }

    // ---- dvindy, k = 0 (M:1905-1952).  iflag as in the reference. ----------------
    BV_DEV int dvindy0(real t, real (&dky)[NLOC]) {
        const real tfuzz = BV_R(100.0) * kUround * dsign(fabs(c.tn) + fabs(hot.sd(SD_HU)), hot.sd(SD_HU));
        const real tp = c.tn - hot.sd(SD_HU) - tfuzz;
        const real tn1 = c.tn + tfuzz;
        if ((t - tp) * (t - tn1) > BV_R(0.0)) return -2;
        const real s = qdiv(t - c.tn, c.h);
#pragma unroll
        for (int i = 0; i < NLOC; i++) dky[i] = hot.yh(kLmax - 1, i);
#pragma unroll
        for (int j = kLmax - 2; j >= 0; j--) {
#pragma unroll
            for (int i = 0; i < NLOC; i++) dky[i] = hot.yh(j, i) + s * dky[i];
        }
        return 0;
    }

    // ---- dvindy for any k (M:1905-1957), from the current state --------------------------
    BV_DEV int dvindy_k(real t, int k, real (&dky)[NLOC]) {
        const int nq = c.nq, L = nq + 1;
#pragma unroll
        for (int i = 0; i < NLOC; i++) dky[i] = BV_R(0.0);
        if (k < 0 || k > nq) return -1;
        const real tfuzz = BV_R(100.0) * kUround * dsign(fabs(c.tn) + fabs(hot.sd(SD_HU)), hot.sd(SD_HU));
        const real tp = c.tn - hot.sd(SD_HU) - tfuzz;
        const real tn1 = c.tn + tfuzz;
        if ((t - tp) * (t - tn1) > BV_R(0.0)) return -2;
        const real s = (t - c.tn) / c.h;
        int ic = 1;
        if (k != 0)
            for (int jj = L - k; jj <= nq; jj++) ic = ic * jj;
        real cc = (real)ic;
        real col[NLOC];
        get_col(L, col);
#pragma unroll
        for (int i = 0; i < NLOC; i++) dky[i] = cc * col[i];
        if (k != nq) {
#pragma unroll 1
            for (int jb = 1; jb <= nq - k; jb++) {
                const int j = nq - jb, jp1 = j + 1;
                ic = 1;
                if (k != 0)
                    for (int jj = jp1 - k; jj <= j; jj++) ic = ic * jj;
                cc = (real)ic;
                get_col(jp1, col);
#pragma unroll
                for (int i = 0; i < NLOC; i++) dky[i] = cc * col[i] + s * dky[i];
            }
        }
        if (k != 0) {
            const real r = powi(c.h, -k);
#pragma unroll
            for (int i = 0; i < NLOC; i++) dky[i] = r * dky[i];
        }
        return 0;
    }

    // ---- dvhin (M:1738-1851).  y0 = yh[0], ydot = yh[1]; scratch y, acor. ---------
    BV_DEV int dvhin(real t0, real tout, real &h0, int &niter) {
        niter = 0;
        const real tdist = fabs(tout - t0);
        const real tround = kUround * dmax2(fabs(t0), fabs(tout));
        if (tdist < BV_R(2.0) * tround) return -1;
        const real hlb = BV_R(100.0) * tround;
        real hub = BV_R(0.1) * tdist;
        {
            real y0v[NLOC], ydv[NLOC];
#pragma unroll
            for (int i = 0; i < NLOC; i++) { y0v[i] = hot.yh(0, i); ydv[i] = hot.yh(1, i); }
            hub = POL::hin_upper_bound(hub, y0v, ydv, tol);
        }
        int iter = 0;
        real hg = sqrt(hlb * hub);
        if (hub < hlb) {
            h0 = hg;
        } else {
            real hnew = hg;
            bool more = true;
#pragma unroll 1
            while (more) {
                const real h = dsign(hg, tout - t0);
                const real t1 = t0 + h;
#pragma unroll
                for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i) + h * hot.yh(1, i);
                feval(t1, y, acor);
#pragma unroll
                for (int i = 0; i < NLOC; i++) acor[i] = (acor[i] - hot.yh(1, i)) / h;
                const real yddnrm = POL::norm(acor, ewt, nact);
                if (yddnrm * hub * hub > BV_R(2.0)) hnew = sqrt(BV_R(2.0) / yddnrm);
                else hnew = sqrt(hg * hub);
                iter = iter + 1;
                more = false;
                if (iter < 4) {
                    const real hrat = hnew / hg;
                    if ((hrat <= BV_R(0.5)) || (hrat >= BV_R(2.0))) {
                        if ((iter >= 2) && (hnew > BV_R(2.0) * hg)) {
                            hnew = hg;
                        } else {
                            hg = hnew;
                            more = true;
                        }
                    }
                }
            }
            h0 = hnew * BV_R(0.5);
            if (h0 < hlb) h0 = hlb;
            if (h0 > hub) h0 = hub;
        }
        h0 = dsign(h0, tout - t0);
        niter = iter;
        return 0;
    }

    // ---- block C: first call for the problem (M:1312-1381) -----------------------
    // y[] holds the initial values on entry.  Returns 0 or a negative istate.
    BV_DEV int init_problem(real t, real tout) {
        c.tn = t;
        c.jstart = 0;
        hot.si(SI_NHNIL) = 0;
        hot.si(SI_NST) = hot.si(SI_NJE) = hot.si(SI_NNI) = hot.si(SI_NCFN) = hot.si(SI_NETF) = hot.si(SI_NLU) = 0;
        hot.si(SI_NSLJ) = 0;
        hot.sd(SD_HU) = BV_R(0.0);
        hot.si(SI_NQU) = 0;
        hot.si(SI_INIT) = 0;
        c.icf = 0; c.ipup = 0; c.jcur = 0; c.kflag = 0; hot.si(SI_NSLP) = 0; c.newh = 0; c.newq = 0;
        c.eta = BV_R(0.0); hot.sd(SD_HNEW) = BV_R(0.0); hot.sd(SD_CONP) = BV_R(0.0); hot.sd(SD_CRATE) = BV_R(0.0); c.drc = BV_R(0.0); c.acnrm = BV_R(0.0);
        c.rl1 = BV_R(0.0); hot.sd(SD_ETAMAX) = BV_R(0.0); c.hscal = BV_R(0.0); c.rc = BV_R(0.0); c.prl1 = BV_R(0.0);
#pragma unroll
        for (int j = 1; j <= kLmax; j++) { hot.tau(j) = BV_R(0.0); hot.el(j) = BV_R(0.0); }
#pragma unroll
        for (int j = 1; j <= 5; j++) hot.tq(j) = BV_R(0.0);
#pragma unroll
        for (int j = 0; j < kLmax; j++) {
#pragma unroll
            for (int i = 0; i < NLOC; i++) hot.yh(j, i) = BV_R(0.0);
        }
#pragma unroll
        for (int i = 0; i < NLOC; i++) { POL::acsav_set(*this, i, BV_R(0.0)); acor[i] = BV_R(0.0); savf[i] = BV_R(0.0); }
        POL::lin_init(*this);
        feval(t, y, savf);
#pragma unroll
        for (int i = 0; i < NLOC; i++) hot.yh(1, i) = savf[i];
        hot.si(SI_NFE) = 1;
#pragma unroll
        for (int i = 0; i < NLOC; i++) hot.yh(0, i) = y[i];
        c.nq = 1;
        c.L = 2;
        c.h = BV_R(1.0);
        bool bad;
        dewset_invert(bad);
        if (bad) return IST_ILLEGAL;  // message 21
        real h0 = o.h0;
        if ((tout - t) * h0 < BV_R(0.0)) return IST_ILLEGAL;  // message 14, M:1214
        if (o.itask == 4 || o.itask == 5) {  // M:1314-1323
            if ((o.tcrit - tout) * (tout - t) < BV_R(0.0)) return IST_ILLEGAL;  // message 25
            if (h0 != BV_R(0.0) && (t + h0 - o.tcrit) * h0 > BV_R(0.0)) h0 = o.tcrit - t;
        }
        if (h0 == BV_R(0.0)) {
            int niter;
            const int ier = dvhin(t, tout, h0, niter);
            hot.si_add(SI_NFE, niter);
            if (ier != 0) return IST_ILLEGAL;  // message 22
        }
        const real rh = fabs(h0) * o.hmax_inv;
        if (rh > BV_R(1.0)) h0 = h0 / rh;
        c.h = h0;
#pragma unroll
        for (int i = 0; i < NLOC; i++) hot.yh(1, i) = h0 * hot.yh(1, i);
        return 0;
    }

    // ---- dvnlsd, chord iteration (M:2703-2879) -------------------------------------
    // Single-exit loops only: a `return` or `break` out of a divergent region moves the
    // warp's reconvergence point to the end of the function, and the 32 systems of a warp
    // would then run one after the other for the rest of the integration.
    BV_DEV int dvnlsd(int nflag, real rtq4) {
        constexpr int maxcor = 3, msbp = 20;
        constexpr real ccmax = BV_R(0.3), crdown = BV_R(0.3), rdiv = BV_R(2.0);
        if (c.jstart == 0) hot.si(SI_NSLP) = 0;
        if (nflag == 0) c.icf = 0;
        if (nflag == -2) c.ipup = MITER;
        if (c.jstart == 0 || c.jstart == -1) c.ipup = MITER;  // M:2748
        if constexpr (MITER == 0) {
            hot.sd(SD_CRATE) = BV_R(1.0);  // M:2751
        } else {
            c.drc = fabs(c.rc - BV_R(1.0));
            if (c.drc > ccmax || hot.si(SI_NST) >= hot.si(SI_NSLP) + msbp) c.ipup = MITER;
        }
        (void)rtq4;

        int result = 1;  // 1 = another corrector pass is needed
#pragma unroll 1
        while (result == 1) {
            int m = 0;
            real delp = BV_R(0.0), del = BV_R(0.0);
            bool converged = false, singular = false;
#pragma unroll
            for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i);
            feval(c.tn, y, savf);
            hot.si_add(SI_NFE, 1);
            if constexpr (MITER != 0) {
                if (c.ipup > 0) {
                    int ierpj;
                    POL::jac_setup(*this, ierpj);
                    c.ipup = 0;
                    c.rc = BV_R(1.0);
                    c.drc = BV_R(0.0);
                    hot.sd(SD_CRATE) = BV_R(1.0);
                    hot.si(SI_NSLP) = hot.si(SI_NST);
                    singular = (ierpj != 0);
                }
            }
            if (!singular) {
#pragma unroll
                for (int i = 0; i < NLOC; i++) acor[i] = BV_R(0.0);
                // 2/(1+rc), M:2813: rc is fixed during the iterations of one pass
                // 2/(1+rc), BDF only (M:2813): rc is fixed during the iterations of one pass
                const bool do_scale = (METH == 2) && (c.rc != BV_R(1.0));
                const real cscale = (METH == 2) ? qdiv(BV_R(2.0), BV_R(1.0) + c.rc) : BV_R(1.0);
                const real rl1h = c.rl1 * c.h;
                bool iterate = true;
#pragma unroll 1
                while (iterate) {
                    int iersl = 0;
                    if constexpr (MITER != 0) {
                        // chord iteration, M:2800-2821
#pragma unroll
                        for (int i = 0; i < NLOC; i++)
                            y[i] = rl1h * savf[i] - (c.rl1 * hot.yh(1, i) + acor[i]);
                        iersl = POL::solve(*this, y);
                        hot.si_add(SI_NNI, 1);
                    }
                    if (iersl > 0) {  // dvsol failed (diagonal approximation, M:2806)
                        iterate = false;
                        continue;
                    }
                    if constexpr (MITER != 0) {
                        if (do_scale) {
#pragma unroll
                            for (int i = 0; i < NLOC; i++) y[i] = cscale * y[i];
                        }
#if BVODE_STRICT
                        del = POL::norm(y, ewt, nact);
#else
                        // fast build: `del` and `delp` hold the MEAN SQUARE (norm^2); the tests
                        // of M:2842-2855 are applied to squares, sqrt only for the rate estimate
                        del = POL::msq(y, ewt, nact);
#endif
#pragma unroll
                        for (int i = 0; i < NLOC; i++) acor[i] = acor[i] + y[i];
#pragma unroll
                        for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i) + acor[i];
                    } else {
                        // functional iteration, M:2822-2837
#pragma unroll
                        for (int i = 0; i < NLOC; i++) savf[i] = c.rl1 * (c.h * savf[i] - hot.yh(1, i));
#pragma unroll
                        for (int i = 0; i < NLOC; i++) y[i] = savf[i] - acor[i];
#if BVODE_STRICT
                        del = POL::norm(y, ewt, nact);
#else
                        del = POL::msq(y, ewt, nact);
#endif
#pragma unroll
                        for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i) + savf[i];
#pragma unroll
                        for (int i = 0; i < NLOC; i++) acor[i] = savf[i];
                    }
#if BVODE_STRICT
                    if (m != 0) hot.sd(SD_CRATE) = dmax2(crdown * hot.sd(SD_CRATE), qdiv(del, delp));
                    const real dcon = qdiv(del * dmin2(BV_R(1.0), hot.sd(SD_CRATE)), hot.tq(4));
                    const bool diverging = !(m + 1 < 2 || del <= rdiv * delp);
#else
                    if (m != 0) hot.sd(SD_CRATE) = dmax2(crdown * hot.sd(SD_CRATE), sqrt(qdiv(del, delp)));
                    const real cfac = dmin2(BV_R(1.0), hot.sd(SD_CRATE)) * rtq4;
                    const real dcon = del * (cfac * cfac);
                    const bool diverging = !(m + 1 < 2 || del <= (rdiv * rdiv) * delp);
#endif
                    if (dcon <= BV_R(1.0)) {
                        converged = true;
                        iterate = false;
                    } else {
                        m = m + 1;
                        if (m == maxcor || diverging) {
                            iterate = false;
                        } else {
                            delp = del;
                            feval(c.tn, y, savf);
                            hot.si_add(SI_NFE, 1);
                        }
                    }
                }
            }
            if (converged) {
                c.jcur = 0;
                c.icf = 0;
#if BVODE_STRICT
                c.acnrm = (m == 0) ? del : POL::norm(acor, ewt, nact);
#else
                c.acnrm = (m == 0) ? del : POL::msq(acor, ewt, nact);  // squared, see above
#endif
                result = 0;
            } else if (MITER != 0 && !singular && c.jcur != 1) {
                c.icf = 1;
                c.ipup = MITER;
            } else {
                c.icf = 2;
                c.ipup = MITER;
                result = -1;
            }
        }
        return result;
    }

    // eta = 1 / (x**(1/k) + addon), M:2201, 2275, 2282, 2292
    BV_DEV static real eta_of(real x, int k) {
        return qrcp(rpow(x, recip_int(k)) + BV_R(1.0e-6));
    }

    // ---- the step loop: driver blocks D/E/F around dvstep (M:1399-1623, 1975-2377) --
    // Runs the reference's usage loop `do k = 1, nout; call dvode(..., tout(k), itask = 1, ...)`
    // for this system inside ONE loop: when an accepted step passes the current tout the
    // solution is interpolated (block G), handed to `emit(k, y)`, and the next "call" starts
    // (nslast = nst, M:1399; block D, M:1457-1467) without leaving the loop -- so the 32
    // systems of a warp are never made to wait for each other at an output point.
    // `first` is true when init_problem ran in this launch (goto 200, M:1381).
    // Returns the output istate; t_out receives the output t; y the output solution.
    //
    // GENERAL = false is the lean instantiation for itask = 1 (the hot path).  GENERAL = true
    // adds the rest of the driver: itask 2-5 with tcrit / kuth (blocks D and F, M:1399-1483,
    // 1586-1623; dvstep M:2100-2103) and the istate = 3 restart (jstart = -1, M:2031-2079).
    //
    // Systems come and go INSIDE the loop: `next(have, ist, t_out)` is called whenever this lane
    // has no system to work on -- at the start (have = false), and when its system is finished
    // (have = true; ist / t_out are its output istate and t, this->y its output solution).  It
    // stores the finished system, loads another one (init_problem for istate = 1, the saved
    // state otherwise) and returns 0 = no more work, 1 = a first call (init_problem ran in this
    // launch: goto 200, M:1381), 2 = a continuation call.  A kernel that gives each lane exactly
    // one system returns 0 the second time; the thread-per-system kernel hands finished lanes the
    // next system of the batch, so that a warp is not held by its slowest system alone
    // (compaction of finished systems).
    // REFILL = false: `next` runs once before the loop and once after it (the lean shape: the
    // hand-over code is outside the instruction-cache-sensitive step loop); REFILL = true: inside.
    template <bool GENERAL, bool REFILL, class TOUT, class EMIT, class NEXT>
    BV_DEV void run(int nout, TOUT tout_of, EMIT emit, NEXT next) {
        constexpr int kfc = -3, kfh = -7, mxncf = 10;
#ifndef BVODE_REFILL_GROUP
#define BVODE_REFILL_GROUP 4
#endif
        [[maybe_unused]] constexpr int kRefillGroup = BVODE_REFILL_GROUP;
        [[maybe_unused]] constexpr real addon = BV_R(1.0e-6);
        constexpr real bias1 = BV_R(6.0), bias2 = BV_R(6.0), bias3 = BV_R(10.0), etacf = BV_R(0.25),
                         etamin = BV_R(0.1), etamxf = BV_R(0.2), etamx1 = BV_R(1.0e4), etamx2 = BV_R(10.0),
                         etamx3 = BV_R(10.0), onepsm = BV_R(1.00001);
#if BVODE_STRICT
        constexpr real thresh = BV_R(1.5);
#else
        const lg_t l2thr = (lg_t)-BV_R(0.5849646647653407);  // log2(1/thresh - addon), thresh = 1.5
#endif
        enum { E_NEWSTEP = 0, E_RESCALE = 1, E_PREDICT = 2, E_DOWN_RESCALE = 3 };

        int nslast = 0;  // M:1399 (and 1337 on the first call)
        int lmax = o.maxord + 1;  // per system when the optional inputs are (set again at hand-over)
        int istate = 1;  // 0 = keep stepping; every loop below has a single exit (see dvnlsd).
                         // Non-zero with have == false: nothing loaded yet.
        bool have = false, alive = true;
        int kout = 0;
        real tout = BV_R(0.0);
        // "reached": tout lies inside the step just taken -> interpolate instead of stepping
        bool reached = false;
        // ---- GENERAL: what the current call does next (blocks D and F for every itask)
        enum { ACT_STEP = 0, ACT_RET_INTERP = 1, ACT_RET_YH = 2, ACT_ILLEGAL = 3 };
        [[maybe_unused]] int act = ACT_STEP, kuth = 0;
        [[maybe_unused]] bool ihit = false;
        [[maybe_unused]] real t_emit = tout;
        const int itask = GENERAL ? o.itask : 1;
        // tcrit handling shared by blocks D and F (M:1469-1483, 1600-1612): true = return now
        auto tcrit_logic = [&]() -> bool {
            const real hmx = fabs(c.tn) + fabs(c.h);
            ihit = fabs(c.tn - o.tcrit) <= BV_R(100.0) * kUround * hmx;
            if (ihit) return true;
            const real tnext = c.tn + hot.sd(SD_HNEW) * (BV_R(1.0) + BV_R(4.0) * kUround);
            if ((tnext - o.tcrit) * c.h > BV_R(0.0)) {
                c.h = (o.tcrit - c.tn) * (BV_R(1.0) - BV_R(4.0) * kUround);
                kuth = 1;
            }
            return false;
        };
        // block D, M:1399-1483: entry of a continuation call
        auto block_d = [&](real to) -> int {
            nslast = hot.si(SI_NST);
            kuth = 0;
            ihit = false;
            if (itask == 2) return ACT_STEP;
            if (itask == 3) {
                const real tp = c.tn - hot.sd(SD_HU) * (BV_R(1.0) + BV_R(100.0) * kUround);
                if ((tp - to) * c.h > BV_R(0.0)) return ACT_ILLEGAL;  // message 23
                if ((c.tn - to) * c.h >= BV_R(0.0)) return ACT_RET_YH;
                return ACT_STEP;
            }
            if (itask == 4 || itask == 5) {
                if ((c.tn - o.tcrit) * c.h > BV_R(0.0)) return ACT_ILLEGAL;  // message 24
                if (itask == 4) {
                    if ((o.tcrit - to) * c.h < BV_R(0.0)) return ACT_ILLEGAL;  // message 25
                    if ((c.tn - to) * c.h >= BV_R(0.0)) {
                        real dky[NLOC];
                        if (dvindy0(to, dky) != 0) return ACT_ILLEGAL;  // message 27
                        return ACT_RET_INTERP;
                    }
                }
                return tcrit_logic() ? ACT_RET_YH : ACT_STEP;
            }
            if ((c.tn - to) * c.h < BV_R(0.0)) return ACT_STEP;
            real dky[NLOC];
            if (dvindy0(to, dky) != 0) return ACT_ILLEGAL;  // message 27
            return ACT_RET_INTERP;
        };
        // block F, M:1586-1623: after a successful step
        auto block_f = [&](real to) -> int {
            kuth = 0;
            ihit = false;
            if (itask == 2) return ACT_RET_YH;
            if (itask == 3) return ((c.tn - to) * c.h < BV_R(0.0)) ? ACT_STEP : ACT_RET_YH;
            if (itask == 4) {
                if ((c.tn - to) * c.h < BV_R(0.0)) return tcrit_logic() ? ACT_RET_YH : ACT_STEP;
                return ACT_RET_INTERP;
            }
            if (itask == 5) {
                const real hmx = fabs(c.tn) + fabs(c.h);
                ihit = fabs(c.tn - o.tcrit) <= BV_R(100.0) * kUround * hmx;
                return ACT_RET_YH;
            }
            return ((c.tn - to) * c.h < BV_R(0.0)) ? ACT_STEP : ACT_RET_INTERP;
        };
        int entry = E_NEWSTEP;
        bool skip_ewt = false;
        real told = BV_R(0.0);
        int ncf = 0, nflag = 0;

        auto handover = [&]() {
            // ---- this lane's system is finished (or none is loaded yet): hand it over, get the next
            int ist_ret = 0;
            real t_ret = BV_R(0.0);
            if (have) {
                if (istate == IST_OK) {
                    t_ret = GENERAL ? t_emit : tout;  // y was set by the last emit
                    ist_ret = IST_OK;
                } else if (istate == IST_ILLEGAL_ENTRY) {  // illegal input at a call entry: that call changes nothing
                    t_ret = (GENERAL && kout > 0) ? t_emit : tout;
                    ist_ret = istate;
                } else {
                    // ---- block H: unsuccessful return (M:1708-1723): y <- yh(:,1), t <- tn
#pragma unroll
                    for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i);
                    t_ret = c.tn;
                    ist_ret = istate;
                }
            }
            const int got = next(have, ist_ret, t_ret);
            if (got == 0) {
                alive = false;
            } else {
                // ---- entry of the first call for the new system (block D on a continuation call)
                const bool first = (got == 1);
                if constexpr (GENERAL) lmax = o.maxord + 1;
                have = true;
                istate = 0;
                kout = 0;
                tout = tout_of(0);
                nslast = hot.si(SI_NST);
                reached = !GENERAL && !first && !((c.tn - tout) * c.h < BV_R(0.0));
                if (reached) {
                    // block D on a continuation call checks dvindy's flag (message 27, M:1459-1464)
                    real dky[NLOC];
                    if (dvindy0(tout, dky) != 0) {
                        reached = false;
                        istate = IST_ILLEGAL_ENTRY;
                    }
                }
                if constexpr (GENERAL) {
                    act = ACT_STEP;
                    kuth = 0;
                    ihit = false;
                    t_emit = tout;
                    if (!first) act = block_d(tout);
                }
                entry = E_NEWSTEP;
                skip_ewt = first;
                told = c.tn;
                ncf = 0;
                nflag = 0;
            }
        };
        if constexpr (!REFILL) handover();
#pragma unroll 1
        while (REFILL ? (__any_sync(0xffffffffu, alive) != 0) : (alive && istate == 0)) {
            hot.sync();  // (warp policies) last attempt's reads of the shared control state before this one's writes
            if constexpr (REFILL) {
                // Finished lanes are replaced in groups: a lane waits until kRefillGroup lanes of
                // its warp are idle (or nothing else runs), so that the hand-over code runs for
                // several lanes at once and the systems that start together stay in step.
                const unsigned idle = __ballot_sync(0xffffffffu, alive && istate != 0);
                const unsigned busy = __ballot_sync(0xffffffffu, alive && istate == 0);
                if (alive && istate != 0 && (__popc(idle) >= kRefillGroup || busy == 0u)) handover();
            }
            // ---- block G / D: output points reached by the last step (M:1457-1467, 1618-1622)
#pragma unroll 1
            while (REFILL ? (reached && istate == 0) : reached) {
                dvindy0(tout, y);
                emit(kout, y);
                kout = kout + 1;
                if (kout < nout) {
                    tout = tout_of(kout);
                    nslast = hot.si(SI_NST);
                    reached = !((c.tn - tout) * c.h < BV_R(0.0));
                } else {
                    reached = false;
                    istate = IST_OK;
                }
            }
            if constexpr (GENERAL) {
                // ---- returns of the current call (block G, M:1648-1652) and the next call's entry
#pragma unroll 1
                while (act != ACT_STEP && istate == 0) {
                    if (act == ACT_ILLEGAL) {
                        istate = IST_ILLEGAL_ENTRY;
                    } else {
                        if (act == ACT_RET_INTERP) {
                            dvindy0(tout, y);
                            t_emit = tout;
                        } else {
#pragma unroll
                            for (int i = 0; i < NLOC; i++) y[i] = hot.yh(0, i);
                            t_emit = ihit ? o.tcrit : c.tn;
                        }
                        emit(kout, y);
                        kout = kout + 1;
                        if (kout < nout) {
                            tout = tout_of(kout);
                            act = block_d(tout);
                        } else {
                            istate = IST_OK;
                        }
                    }
                }
            }
            int change = 0;  // pending order change: -1 / +1
            if (istate == 0) {
                if (entry == E_NEWSTEP) {
                    // ---- label 100 / 200, M:1496-1543
                    if (!skip_ewt) {
                        if ((hot.si(SI_NST) - nslast) >= o.mxstep) {
                            istate = IST_MXSTEP;
                        } else {
                            bool bad;
                            dewset_invert(bad);
                            if (bad) istate = IST_EWT;
                        }
                    }
                    skip_ewt = false;
                    if (istate == 0) {
#if BVODE_STRICT
                        const real tolsf = kUround * norm_y0();
                        const bool too_much = (tolsf > BV_R(1.0));
#else
                        // tolsf > 1  <=>  uround^2 * mean-square > 1 (no sqrt on the main path)
                        const bool too_much = ((kUround * kUround) * msq_y0() > BV_R(1.0));
#endif
                        if (too_much) istate = (hot.si(SI_NST) == 0) ? IST_ILLEGAL : IST_TOLSF;
                    }
                    if (istate == 0) {
                        if ((c.tn + c.h) == c.tn) hot.si_add(SI_NHNIL, 1);
                        // ---- dvstep prologue, M:2024-2128
                        c.kflag = 0;
                        told = c.tn;
                        ncf = 0;
                        c.jcur = 0;
                        nflag = 0;
                        if (c.jstart == 0) {
                            c.nq = 1;
                            c.L = 2;
                            hot.tau(1) = c.h;
                            c.prl1 = BV_R(1.0);
                            c.rc = BV_R(0.0);
                            hot.sd(SD_ETAMAX) = etamx1;
                            c.nqwait = 2;
                            c.hscal = c.h;
                            entry = E_PREDICT;
                        } else {
                            if constexpr (GENERAL) {
                                if (c.jstart == -1) {
                                    // ---- istate = 3 restart, M:2044-2079 (savf holds the old
                                    // yh(:, maxord + 2) when nq > maxord, M:1385)
                                    const int maxord = o.maxord;
                                    if (c.newq > maxord) {
                                        const real flotl = (real)lmax;
                                        if (maxord < c.nq - 1 || (maxord == c.nq - 1 && c.newq == c.nq)) {
                                            const real ddn = POL::norm(savf, ewt, nact) / hot.tq(1);
                                            c.eta = BV_R(1.0) / (rpow(bias1 * ddn, BV_R(1.0) / flotl) + addon);
                                        }
#if BVODE_STRICT
                                        if (maxord == c.nq && c.newq == c.nq + 1) c.eta = hot.sd(SD_ETAQ);
                                        if (maxord == c.nq - 1 && c.newq == c.nq + 1) c.eta = hot.sd(SD_ETAQM1);
#else
                                        if (maxord == c.nq && c.newq == c.nq + 1)
                                            c.eta = BV_R(1.0) / (rpow(hot.sd(SD_ETAQ), BV_R(0.5) * recip_int(c.L)) + addon);
                                        if (maxord == c.nq - 1 && c.newq == c.nq + 1)
                                            c.eta = BV_R(1.0) / (rpow(hot.sd(SD_ETAQM1), BV_R(0.5) * recip_int(c.L - 1)) + addon);
#endif
                                        if (maxord == c.nq - 1) order_change<true>(false);  // dvjust(-1)
                                        c.eta = dmin2(c.eta, BV_R(1.0));
                                        c.nq = maxord;
                                        c.L = lmax;
                                        // columns above the new order: zero-column invariant
#pragma unroll
                                        for (int j = 2; j <= kLmax; j++) {
                                            if (j > lmax) {
#pragma unroll
                                                for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = BV_R(0.0);
                                            }
                                        }
                                        c.newq = c.nq;  // (the reference jumps to label 300 here)
                                    }
                                    xcol_mode = 0;
                                    if (kuth == 1) c.eta = dmin2(c.eta, fabs(c.h / c.hscal));
                                    if (kuth == 0) c.eta = dmax2(c.eta, o.hmin / fabs(c.hscal));
                                    c.eta = c.eta / dmax2(BV_R(1.0), fabs(c.hscal) * o.hmax_inv * c.eta);
                                    c.newh = 1;
                                    c.nqwait = c.L;
                                } else if (kuth == 1) {  // h was cut to land on tcrit, M:2100-2103
                                    c.eta = dmin2(c.eta, c.h / c.hscal);
                                    c.newh = 1;
                                }
                            }
                            if (c.newh == 0) {
                                entry = E_PREDICT;
                            } else {
                                if (c.newq < c.nq) change = -1;
                                else if (c.newq > c.nq) change = 1;
                                entry = E_RESCALE;
                            }
                        }
                    }
                } else if (entry == E_DOWN_RESCALE) {
                    change = -1;  // third (or later) error-test failure, M:2230-2235
                    entry = E_RESCALE;
                }
            }
            if (istate == 0) {
                // order change: dvjust + the bookkeeping of M:2118-2128 / M:2231-2234
                if (change != 0) {
                    order_change(change > 0);
                    c.nq = c.nq + change;
                    c.L = c.nq + 1;
                    c.nqwait = c.L;
                }
                rescale(entry == E_RESCALE ? c.eta : BV_R(1.0));

                // ---- label 400: predict, coefficients (M:2147-2158)
                c.tn = c.tn + c.h;
                predict();
                dvset<METH>(c, hot);
                hot.sync();
                c.rl1 = qrcp(hot.el(2));
                c.rc = c.rc * qdiv(c.rl1, c.prl1);
                c.prl1 = c.rl1;
#if BVODE_STRICT
                const real rtq2 = BV_R(0.0), rtq4 = BV_R(0.0);
                (void)rtq2;
#else
                const real rtq2 = qrcp(hot.tq(2));
                const real rtq4 = BV_R(10.0) * rtq2;  // tq(4) = 0.1 * tq(2), M:2559
#endif
                nflag = dvnlsd(nflag, rtq4);

                const bool conv = (nflag == 0);
#if BVODE_STRICT
                const real dsm = conv ? qdiv(c.acnrm, hot.tq(2)) : BV_R(0.0);
#else
                const real dsm = conv ? c.acnrm * (rtq2 * rtq2) : BV_R(0.0);  // dsm^2
#endif
                const bool accept = conv && !(dsm > BV_R(1.0));
                if (!accept) {  // both kinds of failure undo the prediction (M:2180-2187, 2341-2348)
                    c.tn = told;
                    retract();
                }
                // ---- step-size ratio candidates (M:2201, 2275, 2282, 2292).
                // Strict build: eta = 1/(x**(1/k) + addon) evaluated where the reference does.
                // Fast build: the candidates are ordered through log2(x)/k, first in single
                // precision (hardware lg2; a decision is only taken from it when the margin is
                // far above its error, else the double-precision logs decide -- so the choices
                // are those of the double-precision comparison), and ONE double log2 / exp2 pair
                // per attempt serves whichever lane needs a ratio: the failed error test, or
                // the accepted step whose ratio reaches `thresh` (phase B below).
#if BVODE_STRICT
                const real etaq = conv ? eta_of(bias2 * dsm, c.L) : BV_R(0.0);
#else
                const real xq = (bias2 * bias2) * dsm;  // dsm holds dsm^2: log2(x)/(2L)
                int pend = 0;  // 1: ratio for a failed error test, 2: ratio after an accepted step
                int wsel = 0;  // winning candidate: 0 current order, 1 order - 1, 2 order + 1
                real xsel = xq, scsel = BV_R(0.5) * recip_int(c.L);
                float lwf = 0.0f;
#endif
                // hmin/|h| of the failure paths (M:2202, 2221, 2360), one evaluation for all
                const real hmr = accept ? BV_R(0.0) : qdiv(o.hmin, fabs(c.h));
                if (conv && !accept) {
                    // ---- error test failed (M:2177-2237)
                    c.kflag = c.kflag - 1;
                    hot.si_add(SI_NETF, 1);
                    nflag = -2;
                    if (fabs(c.h) <= o.hmin * onepsm) {
                        c.kflag = -1;
                        c.jstart = 1;
                        istate = IST_ERRTEST;
                    } else {
                        hot.sd(SD_ETAMAX) = BV_R(1.0);
                        if (c.kflag > kfc) {
#if BVODE_STRICT
                            c.eta = etaq;
                            c.eta = dmax2(dmax2(c.eta, hmr), etamin);
                            if ((c.kflag <= -2) && (c.eta > etamxf)) c.eta = etamxf;
#else
                            pend = 1;
#endif
                            entry = E_RESCALE;
                        } else if (c.kflag == kfh) {
                            c.kflag = -1;
                            c.jstart = 1;
                            istate = IST_ERRTEST;
                        } else if (c.nq == 1) {
                            c.eta = dmax2(etamin, hmr);
                            c.h = c.h * c.eta;
                            c.hscal = c.h;
                            hot.tau(1) = c.h;
                            feval(c.tn, y, savf);
                            hot.si_add(SI_NFE, 1);
#pragma unroll
                            for (int i = 0; i < NLOC; i++) hot.yh(1, i) = c.h * savf[i];
                            c.nqwait = 10;
                            entry = E_PREDICT;
                        } else {
                            c.eta = dmax2(etamin, hmr);
                            entry = E_DOWN_RESCALE;
                        }
                    }
                } else if (!conv) {
                    // ---- corrector failed to converge (M:2338-2367)
                    ncf = ncf + 1;
                    hot.si_add(SI_NCFN, 1);
                    hot.sd(SD_ETAMAX) = BV_R(1.0);
                    if (fabs(c.h) <= o.hmin * onepsm || ncf == mxncf) {
                        c.kflag = -2;
                        c.jstart = 1;
                        istate = IST_CONVFAIL;
                    } else {
                        c.eta = dmax2(etacf, hmr);
                        nflag = -1;
                        entry = E_RESCALE;
                    }
                } else {
                    // ---- successful step (M:2245-2330, 2370-2375)
                    c.kflag = 0;
                    hot.si_add(SI_NST, 1);
                    hot.sd(SD_HU) = c.h;
                    hot.si(SI_NQU) = c.nq;
                    {   // tau(i) = tau(i-1), i = L..2 (M:2249-2253): all reads, then all writes
                        real tv[kLmax + 1];
#pragma unroll
                        for (int i = 1; i < kLmax; i++) tv[i] = hot.tau(i);
                        hot.sync();
#pragma unroll
                        for (int i = kLmax; i >= 2; i--) {
                            if (i <= c.nq + 1) hot.tau(i) = tv[i - 1];
                        }
                        hot.tau(1) = c.h;
                        hot.sync();
                    }
#pragma unroll
                    for (int j = 1; j <= kLmax; j++) {
                        const real elj = hot.el(j);
#pragma unroll
                        for (int i = 0; i < NLOC; i++) hot.yh(j - 1, i) = hot.yh(j - 1, i) + elj * acor[i];
                    }
                    c.nqwait = c.nqwait - 1;
                    if ((c.L != lmax) && (c.nqwait == 1)) {
#pragma unroll
                        for (int i = 0; i < NLOC; i++) POL::acsav_set(*this, i, acor[i]);
                        hot.sd(SD_CONP) = hot.tq(5);
                    }
                    if (hot.sd(SD_ETAMAX) != BV_R(1.0)) {
                        int newq = c.nq;
                        bool save_acor = false;
#if BVODE_STRICT
                        real eta = etaq;
                        hot.sd(SD_ETAQ) = etaq;
#else
                        pend = 2;
                        const float scqf = (float)scsel;
                        lwf = log2_f32(xq) * scqf;
#endif
                        if (c.nqwait == 0) {
                            c.nqwait = 2;
                            real ddn = BV_R(0.0), dup = BV_R(0.0);
                            if (c.nq != 1) {
                                real top[NLOC];
                                get_col(c.L, top);
#if BVODE_STRICT
                                ddn = qdiv(POL::norm(top, ewt, nact), hot.tq(1));
#else
                                {
                                    const real r1 = qrcp(hot.tq(1));
                                    ddn = POL::msq(top, ewt, nact) * (r1 * r1);  // ddn^2
                                }
#endif
                            }
                            if (c.L != lmax) {
                                const real cnquot =
                                    qdiv(hot.tq(5), hot.sd(SD_CONP)) * powi_small<kLmax>(qdiv(c.h, hot.tau(2)), c.L);
#pragma unroll
                                for (int i = 0; i < NLOC; i++)
                                    savf[i] = acor[i] - cnquot * POL::acsav_get(*this, i);
#if BVODE_STRICT
                                dup = qdiv(POL::norm(savf, ewt, nact), hot.tq(3));
#else
                                {
                                    const real r3 = qrcp(hot.tq(3));
                                    dup = POL::msq(savf, ewt, nact) * (r3 * r3);  // dup^2
                                }
#endif
                            }
#if BVODE_STRICT
                            const real etaqm1 = (c.nq != 1) ? eta_of(bias1 * ddn, c.L - 1) : BV_R(0.0);
                            const real etaqp1 = (c.L != lmax) ? eta_of(bias3 * dup, c.L + 1) : BV_R(0.0);
                            hot.sd(SD_ETAQM1) = etaqm1;
                            if (etaq < etaqp1) {
                                if (etaqp1 <= etaqm1) {
                                    eta = etaqm1;  // label 420
                                    newq = c.nq - 1;
                                } else {
                                    eta = etaqp1;
                                    newq = c.nq + 1;
                                    save_acor = true;
                                }
                            } else if (etaq < etaqm1) {
                                eta = etaqm1;  // label 420
                                newq = c.nq - 1;
                            }
#else
                            // eta_a < eta_b  <=>  log2(x_a)/k_a > log2(x_b)/k_b ; an order that is
                            // not available (ratio 0) has log-ratio +infinity
                            const bool has_m = (c.nq != 1), has_p = (c.L != lmax);
                            const real xm = (bias1 * bias1) * ddn, xp = (bias3 * bias3) * dup;
                            const real scm = BV_R(0.5) * recip_int(c.L - 1), scp = BV_R(0.5) * recip_int(c.L + 1);
                            auto choose = [](float lq_, float lm_, float lp_) -> int {
                                return choose_ratio(lq_, lm_, lp_);
                            };
                            const float lmf = has_m ? log2_f32(xm) * (float)scm : 1.0e30f;
                            const float lpf = has_p ? log2_f32(xp) * (float)scp : 1.0e30f;
                            wsel = choose(lwf, lmf, lpf);
                            if (near_f32(lwf, lpf) || near_f32(lpf, lmf) || near_f32(lwf, lmf)) {
                                wsel = choose_ratio_f64(xq, xm, xp, scsel, scm, scp, has_m, has_p);
                            }
                            if (wsel == 1) { newq = c.nq - 1; xsel = xm; scsel = scm; lwf = lmf; }
                            if (wsel == 2) {
                                newq = c.nq + 1; xsel = xp; scsel = scp; lwf = lpf; save_acor = true;
                                hot.sd(SD_ETAQ) = xq;
                                hot.sd(SD_ETAQM1) = xm;
                            }
#endif
                        }
                        if (save_acor) {
#pragma unroll
                            for (int i = 0; i < NLOC; i++) POL::acsav_set(*this, i, acor[i]);
                        }
#if BVODE_STRICT
                        // label 450, M:2320-2330 (etamax != 1 here)
                        if (eta < thresh) {
                            c.newq = c.nq;
                            c.newh = 0;
                            c.eta = BV_R(1.0);
                            hot.sd(SD_HNEW) = c.h;
                        } else {
                            c.newq = newq;
                            c.eta = dmin2(eta, hot.sd(SD_ETAMAX));
                            c.eta = qdiv(c.eta, dmax2(BV_R(1.0), fabs(c.h) * o.hmax_inv * c.eta));
                            c.newh = 1;
                            hot.sd(SD_HNEW) = c.h * c.eta;
                        }
#else
                        c.newq = newq;  // the candidate; phase B resets it when the step size is kept
                        c.eta = hot.sd(SD_ETAMAX);  // ... and needs this step's etamax, reset below
#endif
                    } else {
                        if (c.nqwait < 2) c.nqwait = 2;
                        c.newq = c.nq;
                        c.newh = 0;
                        c.eta = BV_R(1.0);
                        hot.sd(SD_HNEW) = c.h;
                    }
                    // label 500
                    hot.sd(SD_ETAMAX) = etamx3;
                    if (hot.si(SI_NST) <= 10) hot.sd(SD_ETAMAX) = etamx2;
                    {
#if BVODE_STRICT
                        const real r = qrcp(hot.tq(2));
#else
                        const real r = rtq2;
#endif
#pragma unroll
                        for (int i = 0; i < NLOC; i++) acor[i] = r * acor[i];
                    }
                    c.jstart = 1;
                    // ---- block F, itask = 1 (M:1586-1622)
                    hot.si(SI_INIT) = 1;
                    entry = E_NEWSTEP;
                    if constexpr (!GENERAL) reached = !((c.tn - tout) * c.h < BV_R(0.0));
                }
#if !BVODE_STRICT
                // ---- phase B: the one ratio evaluation of this attempt
                if (pend != 0) {
                    const float l2thr_f = -0.58496466f;
                    const bool keepf = (lwf > l2thr_f);
                    const bool unsure = (pend == 2) && near_f32(lwf, l2thr_f);
                    real lsel = BV_R(0.0);
                    if (pend == 1 || !keepf || unsure) lsel = log2_fast(xsel) * scsel;
                    const bool keep = (pend == 2) && (unsure ? (lsel > l2thr) : keepf);
                    if (keep) {  // label 450, M:2320-2324
                        c.newq = c.nq;
                        c.newh = 0;
                        c.eta = BV_R(1.0);
                        hot.sd(SD_HNEW) = c.h;
                    } else {
                        const real eta = qrcp(exp2_fast(lsel) + addon);
                        if (pend == 1) {  // M:2201-2204
                            c.eta = dmax2(dmax2(eta, hmr), etamin);
                            if ((c.kflag <= -2) && (c.eta > etamxf)) c.eta = etamxf;
                        } else {  // M:2325-2330; c.eta holds this step's etamax
                            c.eta = dmin2(eta, c.eta);
                            const real d = dmax2(BV_R(1.0), fabs(c.h) * o.hmax_inv * c.eta);
                            if (d != BV_R(1.0)) c.eta = qdiv(c.eta, d);
                            c.newh = 1;
                            hot.sd(SD_HNEW) = c.h * c.eta;
                        }
                    }
                }
#endif
                // block F of the full driver (after dvstep has set hnew, which it reads)
                if constexpr (GENERAL) {
                    if (accept) act = block_f(tout);
                }
            }  // if (istate == 0)
        }
        if constexpr (!REFILL) {
            if (have) handover();
        }
    }
};

}  // namespace BVODE_NS
