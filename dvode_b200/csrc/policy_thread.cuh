// policy_thread.cuh -- one ODE system per thread (tiny systems, e.g. Robertson n = 3).
//
// Every vector of the system (N doubles) and the N x N Newton matrix live in the
// thread's registers; all loops are unrolled to the compile-time N, and run-time
// pivot rows are handled with selects, so a warp executes one instruction stream
// whatever each of its 32 systems decides.
//
// Restates, from /root/reference/src:
//   dvjac miter = 1, 2     dvode_module.F90:2930-3001
//   dvsol default branch   dvode_module.F90:3187-3188
//   dgefa                  dvode_linpack_module.F90:46-120
//   dgesl (job = 0)        dvode_linpack_module.F90:206-225
//   idamax tie rule        dvode_blas_module.F90:311-359 (first maximum wins)
#pragma once
#include "vode_core.cuh"

namespace BVODE_NS {

template <class PROB, int MITER_, bool SAVEJ_, int METH_ = 2>
struct ThreadPolicy {
    using Prob = PROB;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = N;
    static constexpr int MITER = MITER_;     // 0 functional iteration, 1 analytic dense J,
                                             // 2 finite-difference dense J, 3 diagonal approximation
    static constexpr int METH = METH_;       // 1 Adams, 2 BDF
    static constexpr bool SAVEJ = SAVEJ_;    // jsv = +1 (mf > 0): keep a copy of J (miter 1, 2)
    static constexpr int kLmax = kLmaxOf(METH);
    static_assert(MITER >= 0 && MITER <= 3, "thread-per-system policy: miter 0..3");
    static_assert(N >= 2, "miter = 3 keeps wm(2) and the diagonal in the N x N matrix slots");

    typedef real Par[PROB::NPAR > 0 ? PROB::NPAR : 1];
    struct AcSav {};  // the acor copy the reference parks in yh(:,lmax) (M:2259, 2298) is hot.acs(i)
    // rtol / atol expanded to N entries.  Batch-uniform tolerances are copies of kernel parameters
    // (the compiler reads them straight from the constant bank); per-system ones (KArgs::tol_ps,
    // full-driver kernel only) are loaded with the system.
    struct Tol {
        real rv[N], av[N];
        BV_DEV real rtol(int i) const { return rv[i]; }
        BV_DEV real atol(int i) const { return av[i]; }
    };

    // Hot per-system data.  The Nordsieck array always lives in registers.  Everything that is
    // produced in one part of a step attempt and consumed in another -- tau(1:6), the LU factors
    // and pivots, el / tq (dvset -> accept update / order selection), conp, hu, hnew, crate,
    // etamax and the counters -- is parked in shared memory (BVODE_THREAD_SMEM = 1, default),
    // element e of thread t at sm[e * TPB + t] (conflict-free, compile-time offsets): that cuts
    // the registers live across the corrector from ~250 to the 168 that let three blocks (12
    // warps) share an SM instead of two.  The accesses are volatile so that the compiler does
    // not promote them back into registers across the step loop.  BVODE_THREAD_SMEM = 0 keeps
    // everything in registers (255, 8 warps / SM).
#ifndef BVODE_THREAD_SMEM
#define BVODE_THREAD_SMEM 1
#endif
#ifndef BVODE_THREAD_TPB
#define BVODE_THREAD_TPB 32
#endif
    // threads (= systems) per CTA: one warp, so that a finished warp frees its registers and shared
    // memory for the next CTA at once (measured 128 / 64 / 32: 8.72 / 8.75 / 8.80 M systems/s)
    static constexpr int kTPB = BVODE_THREAD_TPB;
    // doubles: tau[lmax], p[N][N], el[lmax], tq[5], sd[SD_COUNT], acs[N] (the acor copy of
    // M:2259), wj[N][N] (saved Jacobian wm(locjs), mf > 0); ints: si[SI_COUNT], ipvt[N]
    static constexpr int kOffTau = 0, kOffP = kOffTau + kLmax, kOffEl = kOffP + N * N,
                         kOffTq = kOffEl + kLmax, kOffSd = kOffTq + 5, kOffAcs = kOffSd + SD_COUNT,
                         kOffWj = kOffAcs + N, kHotDoubles = kOffWj + (SAVEJ ? N * N : 0),
                         kHotInts = SI_COUNT + N;
#if BVODE_THREAD_SMEM
    struct Hot {
        real yhv[kLmax][N];
        volatile real *b;  // this thread's element 0 of the double region
        volatile int *bi;    // ... of the int region
        BV_DEV real &yh(int j, int i) { return yhv[j][i]; }
        BV_DEV volatile real &tau(int i) const { return b[(kOffTau + (i - 1)) * kTPB]; }
        BV_DEV volatile real &p(int i, int j) const { return b[(kOffP + i * N + j) * kTPB]; }
        BV_DEV volatile real &el(int i) const { return b[(kOffEl + (i - 1)) * kTPB]; }
        BV_DEV volatile real &tq(int i) const { return b[(kOffTq + (i - 1)) * kTPB]; }
        BV_DEV volatile real &sd(int k) const { return b[(kOffSd + k) * kTPB]; }
        BV_DEV volatile real &acs(int i) const { return b[(kOffAcs + i) * kTPB]; }
        BV_DEV volatile real &wj(int k) const { return b[(kOffWj + k) * kTPB]; }
        BV_DEV volatile int &si(int k) const { return bi[k * kTPB]; }
        BV_DEV volatile int &ipvt(int i) const { return bi[(SI_COUNT + i) * kTPB]; }
        // this thread's own slot: one fire-and-forget shared-memory reduction instead of load -> add -> store (the
        // load's latency sat on the step loop's critical path: 2.3 % of the samples, ncu r2k)
        BV_DEV void si_add(int k, int d) const { atomicAdd(const_cast<int *>(bi + k * kTPB), d); }
        BV_DEV void sync() const {}
        BV_DEV void bind(real *smem, int tid) {
            b = smem + tid;
            bi = reinterpret_cast<int *>(smem + (size_t)kHotDoubles * kTPB) + tid;
        }
    };
    static constexpr size_t kSmemBytes = (sizeof(real) * kHotDoubles + sizeof(int) * kHotInts) * kTPB;
#ifndef BVODE_THREAD_MINBLOCKS
#define BVODE_THREAD_MINBLOCKS (384 / BVODE_THREAD_TPB)  // 12 warps per SM
#endif
    static constexpr int kMinBlocks = BVODE_THREAD_MINBLOCKS;
#else
    struct Hot {
        real yhv[kLmax][N], tauv[kLmax + 1], pv[N][N], elv[kLmax + 1], tqv[6], sdv[SD_COUNT];
        real acsv[N], wjv[N * N];
        int siv[SI_COUNT], ipv[N];
        BV_DEV real &acs(int i) { return acsv[i]; }
        BV_DEV real &wj(int k) { return wjv[k]; }
        BV_DEV real &yh(int j, int i) { return yhv[j][i]; }
        BV_DEV real &tau(int i) { return tauv[i]; }
        BV_DEV real &p(int i, int j) { return pv[i][j]; }
        BV_DEV real &el(int i) { return elv[i]; }
        BV_DEV real &tq(int i) { return tqv[i]; }
        BV_DEV real &sd(int k) { return sdv[k]; }
        BV_DEV int &si(int k) { return siv[k]; }
        BV_DEV int &ipvt(int i) { return ipv[i]; }
        BV_DEV void si_add(int k, int d) { siv[k] = siv[k] + d; }
        BV_DEV void sync() const {}
        BV_DEV void bind(real *, int) {}
    };
    static constexpr size_t kSmemBytes = 0;
    static constexpr int kMinBlocks = 2;
#endif

    struct Lin {};  // P = I - h*rl1*J / its LU factors: hot.p(i, j) (row i, col j), pivot rows
                    // hot.ipvt(k) (0-based), saved Jacobian hot.wj(i*N + j); fast build: the
                    // diagonal holds 1/u_kk after dgefa

    template <class ENG>
    BV_DEV static real acsav_get(ENG &e, int i) { return e.hot.acs(i); }
    template <class ENG>
    BV_DEV static void acsav_set(ENG &e, int i, real x) { e.hot.acs(i) = x; }

    template <class ENG>
    BV_DEV static void lin_init(ENG &e) {
#pragma unroll
        for (int i = 0; i < N; i++) e.hot.ipvt(i) = i;
    }

    // state slots of one system, in visit order: yh, acsav, acor, p, wjs, ipvt, tau, scalars
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) {
#pragma unroll
        for (int i = 0; i < N; i++) v.d(e.hot.acs(i));
    }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) v.d(e.hot.p(i, j));
        if (SAVEJ) {
#pragma unroll
            for (int k = 0; k < N * N; k++) v.d(e.hot.wj(k));
        }
#pragma unroll
        for (int i = 0; i < N; i++) v.i(e.hot.ipvt(i));
    }

    // dewset (M:3238-3273) for one component, or the problem's own (M:431-467)
    BV_DEV static real ewt_of(int i, real yi, real rtol, real atol) {
        if constexpr (HasUserEwt<PROB>::value) return PROB::ewt(i, yi, rtol, atol);
        else return rtol * fabs(yi) + atol;
    }
    // dvnorm_default, M:3295-3311: sequential sum in index order, like the reference
    // (or the problem's own norm, M:468-503).
    BV_DEV static int comp(int i) { return i; }  // the equation that vector element i of this thread holds
    // (nact: equations in the system, M:3295 `n` -- N unless an istate = 3 call reduced neq; the dropped
    // components carry zeros, so only the divisor changes)
    BV_DEV static real norm(const real (&v)[N], const real (&w)[N], int nact = N) {
        if constexpr (HasUserNorm<PROB>::value) return PROB::norm(v, w);
        real sum = BV_R(0.0);
#pragma unroll
        for (int i = 0; i < N; i++) {
            const real p = v[i] * w[i];
            sum = sum + p * p;
        }
#if BVODE_STRICT
        return sqrt(sum / (real)nact);
#else
        return sqrt((nact == N) ? sum * (BV_R(1.0) / (real)N) : sum / (real)nact);
#endif
    }
#if !BVODE_STRICT
    // mean square (norm^2) -- the fast build tests squares and keeps sqrt off the main path
    BV_DEV static real msq(const real (&v)[N], const real (&w)[N], int nact = N) {
        if constexpr (HasUserNorm<PROB>::value) {  // a user norm is only known as a norm
            const real nv = PROB::norm(v, w);
            return nv * nv;
        }
        real sum = BV_R(0.0);
#pragma unroll
        for (int i = 0; i < N; i++) {
            const real p = v[i] * w[i];
            sum = sum + p * p;
        }
        return (nact == N) ? sum * (BV_R(1.0) / (real)N) : sum / (real)nact;
    }
#endif
    BV_DEV static bool any_le_zero(const real (&v)[N]) {
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; i++) bad = bad || (v[i] <= BV_R(0.0));
        return bad;
    }
    // dvhin's upper bound loop, M:1784-1790
    BV_DEV static real hin_upper_bound(real hub, const real (&y0)[N], const real (&ydot)[N],
                                         const Tol &tol) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const real delyi = BV_R(0.1) * fabs(y0[i]) + tol.atol(i);
            const real afi = fabs(ydot[i]);
            if (afi * hub > delyi) hub = delyi / afi;
        }
        return hub;
    }
    BV_DEV static void f(real t, const real (&y)[N], real (&ydot)[N], const Par &par) {
        PROB::f(t, y, ydot, par);
    }
    // compute_imxer, M:1694-1706 (1-based index)
    BV_DEV static int imxer(const real (&acor)[N], const real (&ewt)[N]) {
        real big = BV_R(0.0);
        int im = 1;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const real size = fabs(acor[i] * ewt[i]);
            if (big < size) { big = size; im = i + 1; }
        }
        return im;
    }

    // ---- dgefa, LP:46-120 ----------------------------------------------------------
    BV_DEV static int dgefa(real (&a)[N][N], int (&ipvt)[N]) {
#define a_(i, j) a[i][j]
        int info = 0;
#pragma unroll
        for (int k = 0; k < N - 1; k++) {
            // idamax over rows k..N-1 of column k
            int l = k;
            real dmax = fabs(a_(k, k));
#pragma unroll
            for (int i = k + 1; i < N; i++) {
                const real xmag = fabs(a_(i, k));
                if (xmag > dmax) { l = i; dmax = xmag; }
            }
            ipvt[k] = l;
            real alk = a_(k, k);
#pragma unroll
            for (int i = k + 1; i < N; i++) alk = (l == i) ? a_(i, k) : alk;
            if (alk == BV_R(0.0)) {
                info = k + 1;
            } else {
                // interchange a(l,k) <-> a(k,k)
                const real akk = a_(k, k);
#pragma unroll
                for (int i = k + 1; i < N; i++) a_(i, k) = (l == i) ? akk : a_(i, k);
#if BVODE_STRICT
                a_(k, k) = alk;
                const real t = -BV_R(1.0) / alk;
#else
                // fast build: the diagonal keeps 1/u_kk so that dgesl multiplies; one reciprocal
                // serves both the multipliers and the back substitution
                const real rp = qrcp(alk);
                a_(k, k) = rp;
                const real t = -rp;
#endif
#pragma unroll
                for (int i = k + 1; i < N; i++) a_(i, k) = t * a_(i, k);
#pragma unroll
                for (int j = k + 1; j < N; j++) {
                    real tj = a_(k, j);
#pragma unroll
                    for (int i = k + 1; i < N; i++) tj = (l == i) ? a_(i, j) : tj;
                    const real akj = a_(k, j);
#pragma unroll
                    for (int i = k + 1; i < N; i++) a_(i, j) = (l == i) ? akj : a_(i, j);
                    a_(k, j) = tj;
#pragma unroll
                    for (int i = k + 1; i < N; i++) a_(i, j) = a_(i, j) + tj * a_(i, k);
                }
            }
        }
        ipvt[N - 1] = N - 1;
        if (a_(N - 1, N - 1) == BV_R(0.0)) info = N;
#if !BVODE_STRICT
        a_(N - 1, N - 1) = qrcp(a_(N - 1, N - 1));
#endif
        return info;
#undef a_
    }

    // ---- dgesl job = 0, LP:206-225 -------------------------------------------------
    BV_DEV static void dgesl(const real (&a)[N][N], const int (&ipvt)[N], real (&b)[N]) {
#define a_(i, j) a[i][j]
#pragma unroll
        for (int k = 0; k < N - 1; k++) {
            const int l = ipvt[k];
            real t = b[k];
#pragma unroll
            for (int i = k + 1; i < N; i++) t = (l == i) ? b[i] : t;
            const real bk = b[k];
#pragma unroll
            for (int i = k + 1; i < N; i++) b[i] = (l == i) ? bk : b[i];
            b[k] = t;
#pragma unroll
            for (int i = k + 1; i < N; i++) b[i] = b[i] + t * a_(i, k);
        }
#pragma unroll
        for (int k = N - 1; k >= 0; k--) {
#if BVODE_STRICT
            b[k] = b[k] / a_(k, k);
#else
            b[k] = b[k] * a_(k, k);
#endif
            const real t = -b[k];
#pragma unroll
            for (int i = 0; i < k; i++) b[i] = b[i] + t * a_(i, k);
        }
#undef a_
    }

    // dvsol, M:3139-3191 (returns iersl)
    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[N]) {
        if constexpr (MITER == 3) {
            // diagonal approximation, M:3164-3181: wm(2) = hrl1 of the last update is
            // hot.p(1, 0), the inverted diagonal wm(3:) is hot.p(0, i)
            const real phrl1 = e.hot.p(1, 0);
            const real hrl1 = e.c.h * e.c.rl1;
            e.hot.p(1, 0) = hrl1;
            int iersl = 0;
            if (hrl1 != phrl1) {
                const real r = hrl1 / phrl1;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    if (iersl == 0) {
                        const real di = BV_R(1.0) - r * (BV_R(1.0) - BV_R(1.0) / e.hot.p(0, i));
                        if (fabs(di) == BV_R(0.0)) iersl = 1;
                        else e.hot.p(0, i) = BV_R(1.0) / di;
                    }
                }
            }
            if (iersl == 0) {
#pragma unroll
                for (int i = 0; i < N; i++) x[i] = e.hot.p(0, i) * x[i];
            }
            return iersl;
        } else {
            real a[N][N];
            int ipvt[N];
#pragma unroll
            for (int i = 0; i < N; i++) {
#pragma unroll
                for (int j = 0; j < N; j++) a[i][j] = e.hot.p(i, j);
                ipvt[i] = e.hot.ipvt(i);
            }
            dgesl(a, ipvt, x);
            return 0;
        }
    }

    // ---- dvjac, miter = 3 (M:3004-3028) ------------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup_diag(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        // M:3004-3028: diagonal Jacobian approximation from one extra f evaluation
        e.hot.si_add(SI_NJE, 1);
        c.jcur = 1;
        e.hot.p(1, 0) = hrl1;  // wm(2)
        const real r = c.rl1 * BV_R(0.1);
#pragma unroll
        for (int i = 0; i < N; i++) e.y[i] = e.y[i] + r * (c.h * e.savf[i] - e.hot.yh(1, i));
        real w[N];
        e.feval(c.tn, e.y, w);
        e.hot.si_add(SI_NFE, 1);
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (ierpj == 0 && i < e.nact) {
                const real r0 = c.h * e.savf[i] - e.hot.yh(1, i);
                const real di = BV_R(0.1) * r0 - c.h * (w[i] - e.savf[i]);
                real wi = BV_R(1.0);
                if (fabs(r0) >= kUround / e.ewt[i]) {
                    if (fabs(di) == BV_R(0.0)) ierpj = 1;
                    else wi = BV_R(0.1) * r0 / di;
                }
                if (ierpj == 0) e.hot.p(0, i) = wi;
            }
        }
    }

    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        if constexpr (MITER == 3) jac_setup_diag(e, ierpj);
        else jac_setup_dense(e, ierpj);
    }

    // ---- dvjac, miter 1 / 2 (M:2930-3001) -------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup_dense(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        int jok = SAVEJ ? 1 : -1;
        if (SAVEJ) {
            if (e.hot.si(SI_NST) == 0 || e.hot.si(SI_NST) > e.hot.si(SI_NSLJ) + 50) jok = -1;      // msbj = 50, M:1328
            if (c.icf == 1 && c.drc < BV_R(0.2)) jok = -1;              // ccmxj = 0.2, M:1327
            if (c.icf == 2) jok = -1;
        }
        real P[N][N];
        if (jok == -1) {
            e.hot.si_add(SI_NJE, 1);
            e.hot.si(SI_NSLJ) = e.hot.si(SI_NST);
            c.jcur = 1;
            if (MITER == 1) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) P[i][j] = BV_R(0.0);  // M:2947-2949
                PROB::jac(c.tn, e.y, P, e.par);
                if (e.nact != N) {  // the Jacobian of the first nact equations in their own unknowns
#pragma unroll
                    for (int i = 0; i < N; i++)
#pragma unroll
                        for (int j = 0; j < N; j++)
                            if (i >= e.nact || j >= e.nact) P[i][j] = BV_R(0.0);
                }
            } else {
                // finite differences, M:2959-2976 (ftem = acor)
                real fac = norm(e.savf, e.ewt, e.nact);
                real r0 = BV_R(1000.0) * fabs(c.h) * kUround * (real)e.nact * fac;
                if (r0 == BV_R(0.0)) r0 = BV_R(1.0);
                const real srur = sqrt(kUround);  // wm(1), M:1326
#pragma unroll
                for (int j = 0; j < N; j++) {
                    if (j < e.nact) {
                        const real yj = e.y[j];
                        const real r = dmax2(srur * fabs(yj), qdiv(r0, e.ewt[j]));
                        e.y[j] = e.y[j] + r;
                        fac = qrcp(r);
                        e.feval(c.tn, e.y, e.acor);
#pragma unroll
                        for (int i = 0; i < N; i++) P[i][j] = (e.acor[i] - e.savf[i]) * fac;
                        e.y[j] = yj;
                    } else {
#pragma unroll
                        for (int i = 0; i < N; i++) P[i][j] = BV_R(0.0);
                    }
                }
                e.hot.si_add(SI_NFE, e.nact);
            }
            if (SAVEJ) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) e.hot.wj(i * N + j) = P[i][j];
            }
        } else {
            c.jcur = 0;
            if (SAVEJ) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) P[i][j] = e.hot.wj(i * N + j);
            }
        }
        const real con = -hrl1;
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) P[i][j] = con * P[i][j];
#pragma unroll
        for (int i = 0; i < N; i++) P[i][i] = P[i][i] + BV_R(1.0);
        e.hot.si_add(SI_NLU, 1);
        int ipvt[N];
        const int ier = dgefa(P, ipvt);
        if (ier != 0) ierpj = 1;
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < N; j++) e.hot.p(i, j) = P[i][j];
            e.hot.ipvt(i) = ipvt[i];
        }
    }
};

}  // namespace BVODE_NS
