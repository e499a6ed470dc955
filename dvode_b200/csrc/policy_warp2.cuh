// policy_warp2.cuh -- one ODE system per warp for 32 < n <= 64: TWO components per lane.
//
// Component c of a vector lives on lane c & 31, slot c >> 5 (NLOC = 2); a lane whose second
// component does not exist (c >= N) carries 0 there.  The scalar control state is replicated on the
// 32 lanes exactly as in policy_warp.cuh (one copy per warp in shared memory), the engine is the
// same (it is generic in NLOC), and the dense Newton matrix is the same shared-memory layout with
// the same implicit pivoting -- each lane now owns two rows, each with its own position.
//
//   WarpDense2Policy : dense finite-difference Jacobian, mf = +-12 / +-22
//                      dvjac miter 2  M:2941-3001, dgefa LP:46-120, dgesl LP:206-225
//   WarpPlain2Policy : functional iteration / diagonal approximation, mf = 10, 13, 20, 23
// This is the first form of the mapping (one step of elimination per pass; two systems per CTA and
// three CTAs per SM because of the 64 x 65 matrix per warp): it is here for coverage of n up to 64.
#pragma once
#include "policy_warp.cuh"

namespace BVODE_NS {

template <class PROB, int METH_>
struct WarpCommon2 {
    static constexpr int N = PROB::N;
    static constexpr int METH = METH_;
    static constexpr int kLmax = kLmaxOf(METH);
    static constexpr int NLOC = 2;
    static_assert(N > 32 && N <= 64, "two components per lane: 32 < n <= 64");
    typedef real Par[PROB::NPAR_LANE > 0 ? PROB::NPAR_LANE : 1];

    BV_DEV static int lane() { return threadIdx.x & 31; }
    BV_DEV static bool valid(int i) { return lane() + 32 * i < N; }

    struct GlobalVec {
        real *p;
        size_t stride;
        BV_DEV real get(int k) const { return p[(size_t)k * stride]; }
        BV_DEV void set(int k, real v) const { p[(size_t)k * stride] = v; }
    };
    static constexpr int kOffTau = 0, kOffEl = kOffTau + kLmax, kOffTq = kOffEl + kLmax,
                         kOffSd = kOffTq + 5, kHotD = kOffSd + SD_COUNT,
                         kWarpHotDoubles = kHotD + (SI_COUNT + 1) / 2;
    struct Hot {
        real yhv[kLmax][2];
        volatile real *b;  // the counters follow the doubles: si(k) is addressed from b as well (one pointer live)
        BV_DEV real &yh(int j, int i) { return yhv[j][i]; }
        BV_DEV volatile real &tau(int i) const { return b[kOffTau + (i - 1)]; }
        BV_DEV volatile real &el(int i) const { return b[kOffEl + (i - 1)]; }
        BV_DEV volatile real &tq(int i) const { return b[kOffTq + (i - 1)]; }
        BV_DEV volatile real &sd(int k) const { return b[kOffSd + k]; }
        BV_DEV volatile int &si(int k) const { return reinterpret_cast<volatile int *>(b + kHotD)[k]; }
        // The control state is ONE copy per warp that all 32 lanes read and write with the same values.  That is
        // only sound while no lane can read a slot after another lane has already written its next value: a
        // read-modify-write separates its read from its write by a warp barrier, and the engine calls sync()
        // where one phase's reads meet the next phase's writes (independent thread scheduling gives no lockstep
        // guarantee after a lane-dependent branch).
        BV_DEV void si_add(int k, int d) const {
            volatile int *p = reinterpret_cast<volatile int *>(b + kHotD) + k;
            const int v = *p;
            __syncwarp();
            *p = v + d;
            __syncwarp();
        }
        BV_DEV void sync() const { __syncwarp(); }
        BV_DEV void bind(real *base, int) {
            b = base;
        }
    };
    struct AcSav {
        real v[2];
    };
    template <class ENG>
    BV_DEV static real acsav_get(ENG &e, int i) { return e.acsav.v[i]; }
    template <class ENG>
    BV_DEV static void acsav_set(ENG &e, int i, real x) { e.acsav.v[i] = x; }
    struct Tol {
        real r[2], a[2];
        BV_DEV real rtol(int i) const { return r[i]; }
        BV_DEV real atol(int i) const { return a[i]; }
    };

    BV_DEV static real ewt_of(int, real yi, real rtol, real atol) { return rtol * fabs(yi) + atol; }
    // dvnorm_default, M:3295-3311
    BV_DEV static real sumsq(const real (&v)[2], const real (&w)[2]) {
        const real p0 = v[0] * w[0], p1 = v[1] * w[1];
        const real q0 = p0 * p0, q1 = p1 * p1;
#if BVODE_STRICT
        real sum = BV_R(0.0);  // in index order, like the reference's loop
#pragma unroll 1
        for (int i = 0; i < 32; i++) sum = sum + shfl_d(q0, i);
#pragma unroll 1
        for (int i = 32; i < N; i++) sum = sum + shfl_d(q1, i - 32);
        return sum;
#else
        real sum = q0 + (valid(1) ? q1 : BV_R(0.0));
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum = sum + shfl_xor_d(sum, m);
        return sum;
#endif
    }
    BV_DEV static int comp(int i) { return lane() + 32 * i; }
    // nact: the equations integrated after an istate = 3 call with a smaller neq (M:1127-1135); the frozen components
    // carry v = 0, so the sum over all N is the reference's sum over the first nact (see ThreadPolicy::norm)
    BV_DEV static real norm(const real (&v)[2], const real (&w)[2], int nact = N) {
#if BVODE_STRICT
        return sqrt(sumsq(v, w) / (real)nact);
#else
        return sqrt((nact == N) ? sumsq(v, w) * (BV_R(1.0) / (real)N) : sumsq(v, w) / (real)nact);
#endif
    }
#if !BVODE_STRICT
    BV_DEV static real msq(const real (&v)[2], const real (&w)[2], int nact = N) {
        return (nact == N) ? sumsq(v, w) * (BV_R(1.0) / (real)N) : sumsq(v, w) / (real)nact;
    }
#endif
    BV_DEV static bool any_le_zero(const real (&v)[2]) {
        return __any_sync(kFull, (v[0] <= BV_R(0.0)) || (valid(1) && v[1] <= BV_R(0.0))) != 0;
    }
    // dvhin's upper bound loop, M:1784-1790: sequential in i because hub feeds the test
    BV_DEV static real hin_upper_bound(real hub, const real (&y0)[2], const real (&ydot)[2], const Tol &tol) {
        const real d0 = BV_R(0.1) * fabs(y0[0]) + tol.a[0], d1 = BV_R(0.1) * fabs(y0[1]) + tol.a[1];
        const real f0 = fabs(ydot[0]), f1 = fabs(ydot[1]);
#pragma unroll 1
        for (int i = 0; i < 32; i++) {
            const real d = shfl_d(d0, i), a = shfl_d(f0, i);
            if (a * hub > d) hub = d / a;
        }
#pragma unroll 1
        for (int i = 32; i < N; i++) {
            const real d = shfl_d(d1, i - 32), a = shfl_d(f1, i - 32);
            if (a * hub > d) hub = d / a;
        }
        return hub;
    }
    BV_DEV static void f(real t, const real (&y)[2], real (&ydot)[2], const Par &par) {
        PROB::f_lane2(t, y, ydot, par, lane());
        if (!valid(1)) ydot[1] = BV_R(0.0);
    }
    // first maximum of |x| over the lanes with `act`, as in WarpCommon
    BV_DEV static unsigned max_abs_ballot(real x, bool act) {
        const unsigned hi = act ? ((unsigned)__double2hiint(x) & 0x7fffffffu) : 0u;
        const unsigned lo = (unsigned)__double2loint(x);
        const unsigned mhi = __reduce_max_sync(kFull, hi);
        const bool c1 = act && (hi == mhi);
        const unsigned mlo = __reduce_max_sync(kFull, c1 ? lo : 0u);
        return __ballot_sync(kFull, c1 && (lo == mlo));
    }
    // compute_imxer, M:1694-1706: first index of the largest |acor*ewt|
    BV_DEV static int imxer(const real (&acor)[2], const real (&ewt)[2]) {
        const real v0 = fabs(acor[0] * ewt[0]), v1 = valid(1) ? fabs(acor[1] * ewt[1]) : -BV_R(1.0);
        const unsigned b0 = max_abs_ballot(v0, true);
        const int l0 = __ffs((int)b0) - 1;
        const real m0 = shfl_d(v0, l0);
        const unsigned b1 = max_abs_ballot(v1, valid(1));
        const int l1 = __ffs((int)b1) - 1;
        const real m1 = shfl_d(v1, l1);
        const real big = (m1 > m0) ? m1 : m0;
        const int idx = (m1 > m0) ? l1 + 32 : l0;
        return (big > BV_R(0.0)) ? idx + 1 : 1;
    }
    BV_DEV static real fd_r0(const Ctl &c, const real (&savf)[2], const real (&ewt)[2], int nact = N) {
        const real fac = norm(savf, ewt, nact);
        real r0 = BV_R(1000.0) * fabs(c.h) * kUround * (real)nact * fac;
        if (r0 == BV_R(0.0)) r0 = BV_R(1.0);
        return r0;
    }
};

template <class PROB, int MITER_, bool SAVEJ_, int METH_ = 2>
struct WarpDense2Policy : WarpCommon2<PROB, METH_> {
    using Base = WarpCommon2<PROB, METH_>;
    static constexpr int METH = METH_;
    static constexpr int kLmax = Base::kLmax;
    using Prob = PROB;
    using typename Base::GlobalVec;
    using typename Base::Hot;
    using Base::lane;
    using Base::valid;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = 2;
    static constexpr int MITER = MITER_;
    static constexpr bool SAVEJ = SAVEJ_;
    static constexpr bool kBand = false;
    static constexpr bool kPlain = false;
    // One system per CTA, 10 CTAs per SM: a system's matrix is N x (N | 1) doubles of shared memory (18.8 KB for
    // n = 48, 33 KB for n = 64: the carve-out then admits 6), the kernel needs ~80 registers, and it is bound by
    // per-warp latency (issue 33 % at 8 warps / SM, ncu r2q) -- measured on diffusion48: 2 systems x 4 CTAs 274 k,
    // 2 x 5 302 k, 1 x 10 and 1 x 11 309 k, 3 x 3 276 k systems/s.  (Reading the pivot lanes of dgesl four steps
    // ahead, to take the load out of the load -> shuffle -> fma chain of a step: 286 k -- withdrawn.)
#ifndef BVODE_DENSE2_TPB
#define BVODE_DENSE2_TPB 32
#endif
#ifndef BVODE_DENSE2_MINBLOCKS
#define BVODE_DENSE2_MINBLOCKS 10
#endif
    static constexpr int kTPB = BVODE_DENSE2_TPB;
    static constexpr int kMinBlocks = BVODE_DENSE2_MINBLOCKS;
    static constexpr int LDP = N | 1;
    static constexpr int kPermOff = N * LDP + ((N * LDP) & 1);
    static constexpr int kBlockDoubles = kPermOff + 32;  // the matrix, then perm[64] (ints)
    static_assert(MITER == 2, "dense finite-difference Jacobian");

    struct Lin {
        real *p;      // shared memory, N rows x LDP
        int *perm;      // shared memory: perm[k] = row (lane + 32*slot) that is the pivot of step k
        GlobalVec wjs;  // saved Jacobian: element j of this lane's row `slot` at slot*N + j (mf > 0)
        int ipv[2];     // perm[c] for this lane's components c
        int pos[2];     // positions of this lane's rows in LINPACK's storage order
    };
    template <class ENG>
    BV_DEV static void lin_init(ENG &e) {
        e.lin.ipv[0] = lane();
        e.lin.ipv[1] = lane() + 32;
        e.lin.pos[0] = lane();
        e.lin.pos[1] = valid(1) ? lane() + 32 : 128;
        e.lin.perm[lane()] = lane();
        e.lin.perm[lane() + 32] = lane() + 32;
        __syncwarp();
    }
    static int smem_doubles(int, int, bool) { return kBlockDoubles; }
    static constexpr int kLaneSlotWjs = 2 * kLmax + 4;  // yh[lmax][2], acsav[2], acor[2], then wjs[2N]
    static constexpr int kLaneSlots = kLaneSlotWjs + (SAVEJ ? 2 * N : 0);
    static constexpr int kLaneIntSlots = 4;
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) {
        v.dv(e.acsav.v[0]);
        v.dv(e.acsav.v[1]);
    }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
        Lin &l = e.lin;
        if (SAVEJ) v.skip_dv(2 * N);
        v.block(l.p, kBlockDoubles);
        v.iv(l.ipv[0]);
        v.iv(l.ipv[1]);
        v.iv(l.pos[0]);
        v.iv(l.pos[1]);
    }

    // idamax over the rows still below the diagonal: the row (lane + 32*slot) of the first maximum
    // in LINPACK's row order
    BV_DEV static int pivot_row(const real (&v)[2], const bool (&act)[2], const int (&pos)[2]) {
        const real a0 = act[0] ? fabs(v[0]) : -BV_R(1.0), a1 = act[1] ? fabs(v[1]) : -BV_R(1.0);
        const bool s1 = (a1 > a0) || (a1 == a0 && act[1] && pos[1] < pos[0]);
        const real vb = s1 ? v[1] : v[0];
        const bool ab = s1 ? act[1] : act[0];
        const int pb = s1 ? pos[1] : pos[0];
        unsigned b = Base::max_abs_ballot(vb, ab);
        if (b & (b - 1)) {  // warp-uniform
            const bool c2 = (b >> lane()) & 1u;
            const unsigned mp = __reduce_min_sync(kFull, c2 ? (unsigned)pb : 0xffffu);
            b = __ballot_sync(kFull, c2 && ((unsigned)pb == mp));
        }
        const int l = __ffs((int)b) - 1;
        return l + 32 * __shfl_sync(kFull, (int)s1, l);
    }

    BV_DEV static int factor(Lin &lin) { return dgefa(lin); }  // dvjac's call of dgefa (M:2998)
    // ---- dgefa, LP:46-120: implicit pivoting as in WarpDensePolicy, two rows per lane ------------
    BV_DEV static int dgefa(Lin &lin) {
        const int ln = lane();
        const bool ok[2] = {true, valid(1)};
        const unsigned pa = sm_addr(lin.p);
        const unsigned rowa[2] = {pa + ln * (LDP * 8), pa + (ok[1] ? ln + 32 : 0) * (LDP * 8)};
        const unsigned perma = pa + (N * LDP + ((N * LDP) & 1)) * 8;
        int pos[2] = {ln, ok[1] ? ln + 32 : 128};
        int info = 0;
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < N - 1; k++) {
            const real v[2] = {lds_d(rowa[0] + k * 8), lds_d(rowa[1] + k * 8)};
            const bool act[2] = {pos[0] >= k, ok[1] && pos[1] >= k};
            const int piv = pivot_row(v, act, pos);
            const int l = piv & 31, sl = piv >> 5;  // warp-uniform
            const real alk = shfl_d(sl ? v[1] : v[0], l);
            const int pl = __shfl_sync(kFull, sl ? pos[1] : pos[0], l);
            if (k < 32) { if (ln == k) lin.ipv[0] = piv; } else { if (ln == k - 32) lin.ipv[1] = piv; }
            sts_i(perma + k * 4, piv);
            if (pos[0] == k) pos[0] = pl;
            if (pos[1] == k) pos[1] = pl;
            if (ln == l) { if (sl) pos[1] = k; else pos[0] = k; }
            if (alk == BV_R(0.0)) {
                info = k + 1;
            } else {
                const bool below[2] = {pos[0] > k, ok[1] && pos[1] > k};
#if BVODE_STRICT
                const real t = -BV_R(1.0) / alk;
                const real m[2] = {t * v[0], t * v[1]};
                if (below[0]) sts_d(rowa[0] + k * 8, m[0]);
                if (below[1]) sts_d(rowa[1] + k * 8, m[1]);
#else
                const real rp = qrcp(alk);
                const real t = -rp;
                const real m[2] = {t * v[0], t * v[1]};
                sts_d(rowa[0] + k * 8, (ln == l && sl == 0) ? rp : m[0]);
                if (ok[1]) sts_d(rowa[1] + k * 8, (ln == l && sl == 1) ? rp : m[1]);
#endif
                const unsigned pr = pa + piv * (LDP * 8);
                unsigned off = (k + 1) * 8;
                int cnt = N - 1 - k;
#pragma unroll 1
                for (; cnt >= 4; cnt -= 4, off += 32) {
                    const unsigned ap = pr + off;
                    const real p0 = lds_d<0>(ap), p1 = lds_d<8>(ap), p2 = lds_d<16>(ap), p3 = lds_d<24>(ap);
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        if (below[i]) {
                            const unsigned ar = rowa[i] + off;
                            const real r0 = lds_d<0>(ar), r1 = lds_d<8>(ar), r2 = lds_d<16>(ar), r3 = lds_d<24>(ar);
                            sts_d<0>(ar, r0 + p0 * m[i]);
                            sts_d<8>(ar, r1 + p1 * m[i]);
                            sts_d<16>(ar, r2 + p2 * m[i]);
                            sts_d<24>(ar, r3 + p3 * m[i]);
                        }
                    }
                }
#pragma unroll 1
                for (; cnt > 0; cnt--, off += 8) {
                    const real p0 = lds_d(pr + off);
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        if (below[i]) sts_d(rowa[i] + off, lds_d(rowa[i] + off) + p0 * m[i]);
                }
            }
            __syncwarp();
        }
        {   // the row that is left is the last pivot
            const unsigned b0 = __ballot_sync(kFull, pos[0] == N - 1);
            const unsigned b1 = __ballot_sync(kFull, ok[1] && pos[1] == N - 1);
            const int piv = b0 ? __ffs((int)b0) - 1 : __ffs((int)b1) - 1 + 32;
            const int l = piv & 31, sl = piv >> 5;
            if (ln == ((N - 1) & 31)) lin.ipv[(N - 1) >> 5] = piv;
            sts_i(perma + (N - 1) * 4, piv);
            const real v[2] = {lds_d<(N - 1) * 8>(rowa[0]), lds_d<(N - 1) * 8>(rowa[1])};
            const real alk = shfl_d(sl ? v[1] : v[0], l);
            if (alk == BV_R(0.0)) info = N;
#if !BVODE_STRICT
            const real rp = qrcp(alk);
            sts_d<(N - 1) * 8>(rowa[0], (ln == l && sl == 0) ? rp : -rp * v[0]);
            if (ok[1]) sts_d<(N - 1) * 8>(rowa[1], (ln == l && sl == 1) ? rp : -rp * v[1]);
#endif
        }
        lin.pos[0] = pos[0];
        lin.pos[1] = pos[1];
        __syncwarp();
        return info;
    }

    // ---- dgesl job = 0, LP:206-225 -------------------------------------------------------
    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[2]) {
        const Lin &lin = e.lin;
        const int ln = lane();
        const bool ok[2] = {true, valid(1)};
        const unsigned pa = sm_addr(lin.p);
        const unsigned rowa[2] = {pa + ln * (LDP * 8), pa + (ok[1] ? ln + 32 : 0) * (LDP * 8)};
        const unsigned perma = pa + (N * LDP + ((N * LDP) & 1)) * 8;
        const int pos[2] = {lin.pos[0], lin.pos[1]};
        real b[2] = {x[0], x[1]};
#pragma unroll 2
        for (int k = 0; k < N - 1; k++) {
            const int piv = lds_i(perma + k * 4);
            const real t = shfl_d((piv >> 5) ? b[1] : b[0], piv & 31);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const real m = lds_d(rowa[i] + k * 8);
                const bool part = ok[i] && pos[i] > k;
#if BVODE_STRICT
                if (part) b[i] = b[i] + t * m;
#else
                b[i] = b[i] + t * (part ? m : BV_R(0.0));
#endif
            }
        }
#if BVODE_STRICT
#pragma unroll 2
        for (int k = N - 1; k >= 0; k--) {
            const int piv = lds_i(perma + k * 4);
            const real m[2] = {lds_d(rowa[0] + k * 8), lds_d(rowa[1] + k * 8)};
            if (pos[0] == k) b[0] = b[0] / m[0];
            if (ok[1] && pos[1] == k) b[1] = b[1] / m[1];
            const real t = -shfl_d((piv >> 5) ? b[1] : b[0], piv & 31);
            if (pos[0] < k) b[0] = b[0] + t * m[0];
            if (ok[1] && pos[1] < k) b[1] = b[1] + t * m[1];
        }
#else
#pragma unroll 2
        for (int k = N - 1; k >= 1; k--) {
            const int piv = lds_i(perma + k * 4);
            const real t = shfl_d((piv >> 5) ? b[1] : b[0], piv & 31);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const real m = lds_d(rowa[i] + k * 8);
                b[i] = b[i] + t * ((ok[i] && pos[i] < k) ? m : BV_R(0.0));
            }
        }
        b[0] = b[0] * lds_d(rowa[0] + pos[0] * 8);  // x_k = b_k / u_kk
        if (ok[1]) b[1] = b[1] * lds_d(rowa[1] + pos[1] * 8);
#endif
        // component c is in the row that was pivot c
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int src = lin.ipv[j];
            const real s0 = shfl_d(b[0], src & 31), s1 = shfl_d(b[1], src & 31);
            x[j] = ok[j] ? ((src >> 5) ? s1 : s0) : BV_R(0.0);
        }
        return 0;
    }

    // ---- dvjac, miter 2 (M:2930-3001) -------------------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        Lin &l = e.lin;
        const int ln = lane();
        const bool ok[2] = {true, valid(1)};
        const unsigned pa = sm_addr(l.p);
        const unsigned rowa[2] = {pa + ln * (LDP * 8), pa + (ok[1] ? ln + 32 : 0) * (LDP * 8)};
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        int jok = SAVEJ ? 1 : -1;
        if (SAVEJ) {
            if (e.hot.si(SI_NST) == 0 || e.hot.si(SI_NST) > e.hot.si(SI_NSLJ) + 50) jok = -1;
            if (c.icf == 1 && c.drc < BV_R(0.2)) jok = -1;
            if (c.icf == 2) jok = -1;
        }
        __syncwarp();
        const real con = -hrl1;
        if (jok == -1) {
            e.hot.si_add(SI_NJE, 1);
            e.hot.si(SI_NSLJ) = e.hot.si(SI_NST);
            c.jcur = 1;
            // finite differences, M:2959-2976: column j for all rows at once
            const real r0 = Base::fd_r0(c, e.savf, e.ewt, e.nact);
            const real srur = sqrt(kUround);
            const real rmine[2] = {dmax2(srur * fabs(e.y[0]), qdiv(r0, e.ewt[0])),
                                     dmax2(srur * fabs(e.y[1]), qdiv(r0, e.ewt[1]))};
            const real ysave[2] = {e.y[0], e.y[1]};
#pragma unroll 1
            for (int j = 0; j < N; j++) {
                real d0 = BV_R(0.0), d1 = BV_R(0.0);  // neq was reduced: nothing in the columns of the dropped unknowns
                if (j < e.nact) {                   // (warp-uniform; the rows of the dropped equations are zero through feval)
                    const int jl = j & 31, js = j >> 5;
                    const real r = shfl_d(js ? rmine[1] : rmine[0], jl);
                    if (ln == jl) { if (js) e.y[1] = ysave[1] + r; else e.y[0] = ysave[0] + r; }
                    const real fac = qrcp(r);
                    e.feval(c.tn, e.y, e.acor);
                    d0 = (e.acor[0] - e.savf[0]) * fac;
                    d1 = (e.acor[1] - e.savf[1]) * fac;
                }
                if (SAVEJ) {
                    l.wjs.set(j, d0);
                    if (ok[1]) l.wjs.set(N + j, d1);
                }
                sts_d(rowa[0] + j * 8, con * d0);
                if (ok[1]) sts_d(rowa[1] + j * 8, con * d1);
                e.y[0] = ysave[0];
                e.y[1] = ysave[1];
            }
            e.hot.si_add(SI_NFE, e.nact);
        } else {
            c.jcur = 0;
            if (SAVEJ) {
#pragma unroll 1
                for (int j = 0; j < N; j++) {
                    sts_d(rowa[0] + j * 8, con * l.wjs.get(j));
                    if (ok[1]) sts_d(rowa[1] + j * 8, con * l.wjs.get(N + j));
                }
            }
        }
        // P = I - h*rl1*J (M:2984-2992)
        sts_d(rowa[0] + ln * 8, lds_d(rowa[0] + ln * 8) + BV_R(1.0));
        if (ok[1]) sts_d(rowa[1] + (ln + 32) * 8, lds_d(rowa[1] + (ln + 32) * 8) + BV_R(1.0));
        e.hot.si_add(SI_NLU, 1);
        const int ier = dgefa(l);
        if (ier != 0) ierpj = 1;
    }
};

// No Newton matrix: functional iteration (miter = 0, M:2822-2837) or the diagonal Jacobian
// approximation (miter = 3: dvjac M:3004-3028, dvsol M:3164-3181), two diagonal entries per lane.
template <class PROB, int MITER_, int METH_ = 2>
struct WarpPlain2Policy : WarpCommon2<PROB, METH_> {
    using Base = WarpCommon2<PROB, METH_>;
    using Prob = PROB;
    using Base::lane;
    using Base::valid;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = 2;
    static constexpr int MITER = MITER_;
    static constexpr int METH = METH_;
    static constexpr int kLmax = Base::kLmax;
    static constexpr bool SAVEJ = false;
    static constexpr bool kBand = false;
    static constexpr bool kPlain = true;
    static constexpr int kMinBlocks = 4;
    static_assert(MITER == 0 || MITER == 3, "plain policy");

    struct Lin {
        real diag[2];  // wm(2 + c): inverted diagonal of P (miter = 3)
        real phrl1;    // wm(2): h*rl1 at the last update (replicated on all lanes)
    };
    template <class ENG>
    BV_DEV static void lin_init(ENG &e) { e.lin.diag[0] = e.lin.diag[1] = BV_R(1.0); e.lin.phrl1 = BV_R(0.0); }
    static int smem_doubles(int, int, bool) { return 0; }
    static constexpr int kLaneSlots = 2 * kLmax + 4 + 3;  // yh, acsav, acor, diag[2], phrl1
    static constexpr int kLaneIntSlots = 1;
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) {
        v.dv(e.acsav.v[0]);
        v.dv(e.acsav.v[1]);
    }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
        v.dv(e.lin.diag[0]);
        v.dv(e.lin.diag[1]);
        v.dv(e.lin.phrl1);
        int dummy = 0;
        v.iv(dummy);
    }
    // index of the first component with `z`, 64 if none (the reference stops at the first zero:
    // entries before it are already updated)
    BV_DEV static int first_of(bool z0, bool z1) {
        const unsigned m0 = __ballot_sync(kFull, z0), m1 = __ballot_sync(kFull, z1);
        return m0 ? (__ffs((int)m0) - 1) : (m1 ? (__ffs((int)m1) - 1 + 32) : 64);
    }

    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[2]) {
        Lin &l = e.lin;
        const real hrl1 = e.c.h * e.c.rl1;
        const real phrl1 = l.phrl1;
        l.phrl1 = hrl1;
        if (hrl1 != phrl1) {
            const real r = hrl1 / phrl1;
            const real di[2] = {BV_R(1.0) - r * (BV_R(1.0) - BV_R(1.0) / l.diag[0]), BV_R(1.0) - r * (BV_R(1.0) - BV_R(1.0) / l.diag[1])};
            const int first = first_of(fabs(di[0]) == BV_R(0.0), valid(1) && fabs(di[1]) == BV_R(0.0));
            if (lane() < first) l.diag[0] = BV_R(1.0) / di[0];
            if (valid(1) && lane() + 32 < first) l.diag[1] = BV_R(1.0) / di[1];
            if (first < 64) return 1;
        }
        x[0] = l.diag[0] * x[0];
        x[1] = valid(1) ? l.diag[1] * x[1] : BV_R(0.0);
        return 0;
    }

    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        Lin &l = e.lin;
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        e.hot.si_add(SI_NJE, 1);
        c.jcur = 1;
        l.phrl1 = hrl1;
        const real r = c.rl1 * BV_R(0.1);
#pragma unroll
        for (int i = 0; i < 2; i++) e.y[i] = e.y[i] + r * (c.h * e.savf[i] - e.hot.yh(1, i));
        real w[2];
        e.feval(c.tn, e.y, w);
        e.hot.si_add(SI_NFE, 1);
        real r0[2], di[2];
        bool big[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            r0[i] = c.h * e.savf[i] - e.hot.yh(1, i);
            di[i] = BV_R(0.1) * r0[i] - c.h * (w[i] - e.savf[i]);
            big[i] = valid(i) && (fabs(r0[i]) >= kUround / e.ewt[i]);
        }
        const int first = first_of(big[0] && fabs(di[0]) == BV_R(0.0), big[1] && fabs(di[1]) == BV_R(0.0));
        if (lane() < first) l.diag[0] = big[0] ? BV_R(0.1) * r0[0] / di[0] : BV_R(1.0);
        if (valid(1) && lane() + 32 < first) l.diag[1] = big[1] ? BV_R(0.1) * r0[1] / di[1] : BV_R(1.0);
        if (first < 64) ierpj = 1;
    }
};

}  // namespace BVODE_NS
