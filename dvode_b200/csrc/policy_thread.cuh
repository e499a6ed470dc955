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

template <class PROB, int MITER_, bool SAVEJ_>
struct ThreadPolicy {
    using Prob = PROB;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = N;
    static constexpr int MITER = MITER_;     // 1 analytic dense J, 2 finite-difference dense J
    static constexpr bool SAVEJ = SAVEJ_;    // jsv = +1 (mf > 0): keep a copy of J
    static_assert(MITER == 1 || MITER == 2, "thread-per-system policy is dense only");

    // Cold per-system data stays in its HBM state slot instead of registers: it is touched
    // once per Jacobian / LU (every ~5-50 steps), and a warp's access is one coalesced line.
    typedef double Par[PROB::NPAR > 0 ? PROB::NPAR : 1];
    struct GlobalVec {
        double *p;       // slot 0 of this system
        size_t stride;   // distance between slots (= nsys)
        BV_DEV double get(int k) const { return p[(size_t)k * stride]; }
        BV_DEV void set(int k, double v) const { p[(size_t)k * stride] = v; }
    };
    using AcSav = GlobalVec;  // the reference's yh(:,lmax) parking place (M:2259, 2298)
    struct Tol {              // rtol / atol expanded to N entries, in the kernel parameter bank
        const double *r, *a;
        BV_DEV double rtol(int i) const { return r[i]; }
        BV_DEV double atol(int i) const { return a[i]; }
    };

    struct Lin {
        double p[N][N];   // P = I - h*rl1*J, then its LU factors (p[i][j] = row i, col j).
                          // Fast build: the diagonal holds 1/u_kk after dgefa.
        GlobalVec wjs;    // saved Jacobian wm(locjs), element (i,j) at slot i*N+j (mf > 0)
        int ipvt[N];      // 0-based pivot rows
    };

    BV_DEV static void lin_init(Lin &l) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            l.ipvt[i] = i;
#pragma unroll
            for (int j = 0; j < N; j++) l.p[i][j] = 0.0;
        }
    }

    // state slots of one system, in visit order: yh, acsav*, acor, p, wjs*, tau, scalars
    // (* = accessed in place through GlobalVec, skipped by the load / store visitors)
    static constexpr int kSlotAcsav = kLmax * N;
    static constexpr int kSlotWjs = kLmax * N + 2 * N + N * N;

    template <class V>
    BV_DEV static void visit_acsav(AcSav &, V &v) { v.skip_d(N); }
    template <class V>
    BV_DEV static void visit_lin(Lin &l, V &v) {
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) v.d(l.p[i][j]);
        if (SAVEJ) v.skip_d(N * N);
#pragma unroll
        for (int i = 0; i < N; i++) v.i(l.ipvt[i]);
    }

    // dvnorm_default, M:3295-3311: sequential sum in index order, like the reference.
    BV_DEV static double norm(const double (&v)[N], const double (&w)[N]) {
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const double p = v[i] * w[i];
            sum = sum + p * p;
        }
#if BVODE_STRICT
        return sqrt(sum / (double)N);
#else
        return sqrt(sum * (1.0 / (double)N));
#endif
    }
    BV_DEV static bool any_le_zero(const double (&v)[N]) {
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; i++) bad = bad || (v[i] <= 0.0);
        return bad;
    }
    // dvhin's upper bound loop, M:1784-1790
    BV_DEV static double hin_upper_bound(double hub, const double (&y0)[N], const double (&ydot)[N],
                                         const Tol &tol) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const double delyi = 0.1 * fabs(y0[i]) + tol.atol(i);
            const double afi = fabs(ydot[i]);
            if (afi * hub > delyi) hub = delyi / afi;
        }
        return hub;
    }
    BV_DEV static void f(double t, const double (&y)[N], double (&ydot)[N], const Par &par) {
        PROB::f(t, y, ydot, par);
    }
    // compute_imxer, M:1694-1706 (1-based index)
    BV_DEV static int imxer(const double (&acor)[N], const double (&ewt)[N]) {
        double big = 0.0;
        int im = 1;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const double size = fabs(acor[i] * ewt[i]);
            if (big < size) { big = size; im = i + 1; }
        }
        return im;
    }

    // ---- dgefa, LP:46-120 ----------------------------------------------------------
    BV_DEV static int dgefa(double (&a)[N][N], int (&ipvt)[N]) {
        int info = 0;
#pragma unroll
        for (int k = 0; k < N - 1; k++) {
            // idamax over rows k..N-1 of column k
            int l = k;
            double dmax = fabs(a[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; i++) {
                const double xmag = fabs(a[i][k]);
                if (xmag > dmax) { l = i; dmax = xmag; }
            }
            ipvt[k] = l;
            double alk = a[k][k];
#pragma unroll
            for (int i = k + 1; i < N; i++) alk = (l == i) ? a[i][k] : alk;
            if (alk == 0.0) {
                info = k + 1;
            } else {
                // interchange a(l,k) <-> a(k,k)
                const double akk = a[k][k];
#pragma unroll
                for (int i = k + 1; i < N; i++) a[i][k] = (l == i) ? akk : a[i][k];
                a[k][k] = alk;
                const double t = -1.0 / alk;
#pragma unroll
                for (int i = k + 1; i < N; i++) a[i][k] = t * a[i][k];
#pragma unroll
                for (int j = k + 1; j < N; j++) {
                    double tj = a[k][j];
#pragma unroll
                    for (int i = k + 1; i < N; i++) tj = (l == i) ? a[i][j] : tj;
                    const double akj = a[k][j];
#pragma unroll
                    for (int i = k + 1; i < N; i++) a[i][j] = (l == i) ? akj : a[i][j];
                    a[k][j] = tj;
#pragma unroll
                    for (int i = k + 1; i < N; i++) a[i][j] = a[i][j] + tj * a[i][k];
                }
            }
        }
        ipvt[N - 1] = N - 1;
        if (a[N - 1][N - 1] == 0.0) info = N;
#if !BVODE_STRICT
        // fast build: keep 1/u_kk on the diagonal so that dgesl multiplies instead of divides
#pragma unroll
        for (int k = 0; k < N; k++) a[k][k] = 1.0 / a[k][k];
#endif
        return info;
    }

    // ---- dgesl job = 0, LP:206-225 -------------------------------------------------
    BV_DEV static void dgesl(const double (&a)[N][N], const int (&ipvt)[N], double (&b)[N]) {
#pragma unroll
        for (int k = 0; k < N - 1; k++) {
            const int l = ipvt[k];
            double t = b[k];
#pragma unroll
            for (int i = k + 1; i < N; i++) t = (l == i) ? b[i] : t;
            const double bk = b[k];
#pragma unroll
            for (int i = k + 1; i < N; i++) b[i] = (l == i) ? bk : b[i];
            b[k] = t;
#pragma unroll
            for (int i = k + 1; i < N; i++) b[i] = b[i] + t * a[i][k];
        }
#pragma unroll
        for (int k = N - 1; k >= 0; k--) {
#if BVODE_STRICT
            b[k] = b[k] / a[k][k];
#else
            b[k] = b[k] * a[k][k];
#endif
            const double t = -b[k];
#pragma unroll
            for (int i = 0; i < k; i++) b[i] = b[i] + t * a[i][k];
        }
    }

    BV_DEV static void solve(const Lin &l, double (&x)[N]) { dgesl(l.p, l.ipvt, x); }

    // ---- dvjac, miter 1 / 2 (M:2930-3001) -------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        Lin &l = e.lin;
        ierpj = 0;
        const double hrl1 = c.h * c.rl1;
        int jok = SAVEJ ? 1 : -1;
        if (SAVEJ) {
            if (c.nst == 0 || c.nst > c.nslj + 50) jok = -1;      // msbj = 50, M:1328
            if (c.icf == 1 && c.drc < 0.2) jok = -1;              // ccmxj = 0.2, M:1327
            if (c.icf == 2) jok = -1;
        }
        if (jok == -1) {
            c.nje = c.nje + 1;
            c.nslj = c.nst;
            c.jcur = 1;
            if (MITER == 1) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) l.p[i][j] = 0.0;
                PROB::jac(c.tn, e.y, l.p, e.par);
            } else {
                // finite differences, M:2959-2976 (ftem = acor)
                double fac = norm(e.savf, e.ewt);
                double r0 = 1000.0 * fabs(c.h) * kUround * (double)N * fac;
                if (r0 == 0.0) r0 = 1.0;
                const double srur = sqrt(kUround);  // wm(1), M:1326
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const double yj = e.y[j];
                    const double r = dmax2(srur * fabs(yj), r0 / e.ewt[j]);
                    e.y[j] = e.y[j] + r;
                    fac = 1.0 / r;
                    f(c.tn, e.y, e.acor, e.par);
#pragma unroll
                    for (int i = 0; i < N; i++) l.p[i][j] = (e.acor[i] - e.savf[i]) * fac;
                    e.y[j] = yj;
                }
                c.nfe = c.nfe + N;
            }
            if (SAVEJ) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) l.wjs.set(i * N + j, l.p[i][j]);
            }
        } else {
            c.jcur = 0;
            if (SAVEJ) {
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int j = 0; j < N; j++) l.p[i][j] = l.wjs.get(i * N + j);
            }
        }
        const double con = -hrl1;
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) l.p[i][j] = con * l.p[i][j];
#pragma unroll
        for (int i = 0; i < N; i++) l.p[i][i] = l.p[i][i] + 1.0;
        c.nlu = c.nlu + 1;
        const int ier = dgefa(l.p, l.ipvt);
        if (ier != 0) ierpj = 1;
    }
};

}  // namespace BVODE_NS
