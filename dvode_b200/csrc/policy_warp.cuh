// policy_warp.cuh -- one ODE system per warp (n <= 32): component i lives on lane i.
//
// Every vector of the system is ONE double per lane; the scalar control state (step size,
// order, coefficients, counters) is replicated on all 32 lanes, so the control flow of the
// engine is warp-uniform by construction: no divergence, no masks.  Reductions (norms, pivot
// search) are warp shuffles; pivot-row swaps are lane exchanges.
//
//   WarpDensePolicy : dense Newton matrix, row i of P in lane i's registers (mf = +-21, +-22)
//                     dvjac miter 1/2  M:2941-3001, dgefa LP:46-120, dgesl LP:206-225
//   WarpBandPolicy  : LINPACK band storage in shared memory, one column per lane
//                     (mf = +-24, +-25)  dvjac miter 4/5  M:3031-3104, dacopy M:3112,
//                     dgbfa LP:275-396, dgbsl LP:494-519
// idamax tie rule (BL:334): the FIRST index of the largest magnitude wins.
//
// Strict build: norms are summed in index order (a chain of 32 shuffles) so that the result
// is bit-identical to the oracle's sequential loop; the fast build uses a butterfly.
#pragma once
#include "vode_core.cuh"

namespace BVODE_NS {

constexpr unsigned kFull = 0xffffffffu;

// dginv (the dense inversion) keeps a 64-register row: out of line it gets its own register allocation and the
// step loop around it is not squeezed by it (experiments: -DBVODE_DGINV_NOINLINE=0/1)
#ifndef BVODE_DGINV_NOINLINE
#define BVODE_DGINV_NOINLINE 0
#endif
#if BVODE_DGINV_NOINLINE
#define BVODE_DGINV_INLINE __device__ __noinline__
#else
#define BVODE_DGINV_INLINE BV_DEV
#endif
// knobs of the inversion (measured): buffer loads in flight, the barrier that keeps them there
#ifndef BVODE_DGINV_EARLY_RCP
#define BVODE_DGINV_EARLY_RCP 0
#endif
#ifndef BVODE_DGINV_BATCH
#define BVODE_DGINV_BATCH 4
#endif
#ifndef BVODE_DGINV_FENCE
#define BVODE_DGINV_FENCE 1
#endif

// optional members of a warp problem (bvode_ref_problem.cuh: callbacks in the reference's whole-vector form)
template <class P, class = void>
struct HasJacDenseFull { static constexpr bool value = false; };
template <class P>
struct HasJacDenseFull<P, decltype((void)&P::jac_dense_full)> { static constexpr bool value = true; };
template <class P, class = void>
struct HasJacBandFull { static constexpr bool value = false; };
template <class P>
struct HasJacBandFull<P, decltype((void)&P::jac_band_full)> { static constexpr bool value = true; };
template <class P, class = void>
struct HasJacBandCol { static constexpr bool value = false; };
template <class P>
struct HasJacBandCol<P, decltype((void)&P::jac_band_col)> { static constexpr bool value = true; };
template <class P, class = void>
struct ScratchDoublesOf { static constexpr int value = 0; };  // per-warp shared-memory scratch the problem asks for
template <class P>
struct ScratchDoublesOf<P, decltype((void)P::kScratchDoubles)> { static constexpr int value = P::kScratchDoubles; };

BV_DEV real shfl_d(real v, int src) { return __shfl_sync(kFull, v, src); }
BV_DEV real shfl_xor_d(real v, int m) { return __shfl_xor_sync(kFull, v, m); }

// Shared-memory access through a 32-bit shared address held in ONE register, with compile-time
// byte offsets.  The address is made opaque to the compiler once: under the register cap of the
// warp kernels nvcc otherwise re-derives "lane id * row stride + base" from the special
// registers inside every inner-loop iteration (6 of 16 instructions per solve step, ncu r1l).
BV_DEV unsigned sm_addr(const void *p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
template <int OFF = 0>
BV_DEV real lds_d(unsigned a) {
    real v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF = 0>
BV_DEV void sts_d(unsigned a, real v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
template <int OFF = 0>
BV_DEV int lds_i(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
BV_DEV void lds_i4(unsigned a, int &v0, int &v1, int &v2, int &v3) {  // 16-byte aligned
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a));
}
// b += t*m on the lanes where (unsigned)x < (unsigned)lim: one compare and one predicated fma (the
// compiler's own choice is an unconditional fma plus two selects on the dependent chain).
BV_DEV void fma_if_ltu(real &b, real t, real m, int x, int lim) {
    asm("{\n .reg .pred p;\n setp.lt.u32 p, %3, %4;\n @p fma.rn.f64 %0, %1, %2, %0;\n}"
        : "+d"(b) : "d"(t), "d"(m), "r"(x), "r"(lim));
}
// 16-byte shared accesses (two doubles); the address must be 16-byte aligned
template <int OFF = 0>
BV_DEV void lds_d2(unsigned a, real &v0, real &v1) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(v0), "=d"(v1) : "r"(a), "n"(OFF));
}
template <int OFF = 0>
BV_DEV void sts_d2(unsigned a, real v0, real v1) {
    asm volatile("st.shared.v2.f64 [%0+%1], {%2,%3};" ::"r"(a), "n"(OFF), "d"(v0), "d"(v1) : "memory");
}
// an 8-byte store that ptxas must leave alone: adjacent plain stores get merged into 16-byte ones, and when the two
// values do not sit in an aligned register quad that costs four moves per store
BV_DEV void sts_d_single(unsigned a, real v) {
    asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
BV_DEV void sts_i(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Common pieces of the two warp policies.
template <class PROB, int METH_>
struct WarpCommon {
    static constexpr int N = PROB::N;
    static constexpr int METH = METH_;  // 1 Adams, 2 BDF
    static constexpr int kLmax = kLmaxOf(METH);
    static_assert(N <= 32, "warp-per-system policy holds one component per lane");
    static constexpr int NLOC = 1;
    typedef real Par[PROB::NPAR_LANE > 0 ? PROB::NPAR_LANE : 1];

    BV_DEV static int lane() { return threadIdx.x & 31; }

    // Per-lane vector element kept in its HBM state slot (cold data).
    struct GlobalVec {
        real *p;      // this lane's element of slot 0
        size_t stride;  // distance between slots (= nsys * 32)
        BV_DEV real get(int k) const { return p[(size_t)k * stride]; }
        BV_DEV void set(int k, real v) const { p[(size_t)k * stride] = v; }
    };
    // Hot data.  The Nordsieck column entries of this lane's component live in registers.  The
    // control state that is REPLICATED on the 32 lanes (tau, el, tq, conp / hu / hnew / crate /
    // etamax, counters) is kept once per system in shared memory instead: every lane writes the
    // same value to the same address and reads it back as a broadcast, so no synchronisation is
    // needed, and ~55 registers per lane are free for the LU / solve loops (with the 80-register
    // budget of 24 warps / SM the compiler had been re-deriving the lane id and the shared-memory
    // row address inside every inner loop iteration).  Volatile, so the values are not promoted
    // back into registers across the step loop.
    static constexpr int kOffTau = 0, kOffEl = kOffTau + kLmax, kOffTq = kOffEl + kLmax,
                         kOffSd = kOffTq + 5, kHotD = kOffSd + SD_COUNT,
                         kWarpHotDoubles = kHotD + (SI_COUNT + 1) / 2;
    struct Hot {
        real yhv[kLmax][1];
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
    struct AcSav {  // NLOC = 1: a register
        real v;
        BV_DEV real get(int) const { return v; }
        BV_DEV void set(int, real x) { v = x; }
    };
    template <class ENG>
    BV_DEV static real acsav_get(ENG &e, int) { return e.acsav.v; }
    template <class ENG>
    BV_DEV static void acsav_set(ENG &e, int, real x) { e.acsav.v = x; }
    struct Tol {  // this lane's rtol / atol (0 for lanes >= N)
        real r, a;
        BV_DEV real rtol(int) const { return r; }
        BV_DEV real atol(int) const { return a; }
    };

    // dewset_default for this lane's component, M:3238-3273, or the problem's own (M:431-467)
    BV_DEV static real ewt_of(int, real yi, real rtol, real atol) {
        if constexpr (HasUserEwt<PROB>::value) return PROB::ewt(lane(), yi, rtol, atol);
        else return rtol * fabs(yi) + atol;
    }
    // a user-supplied norm (M:468-503) in its lane-wise form: a fold over the components
    BV_DEV static real user_norm(real v, real w) {
        const real p = PROB::norm_term(v, w);
#if BVODE_STRICT
        real acc = PROB::norm_init;
#pragma unroll 1
        for (int i = 0; i < N; i++) acc = PROB::norm_combine(acc, shfl_d(p, i));
#else
        real acc = (lane() < N) ? PROB::norm_combine(PROB::norm_init, p) : PROB::norm_init;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc = PROB::norm_combine(acc, shfl_xor_d(acc, m));
#endif
        return PROB::norm_final(acc);
    }
    // dvnorm_default, M:3295-3311.  Lanes >= N carry v = 0.
    BV_DEV static int comp(int) { return lane(); }  // the equation this lane holds
    BV_DEV static real norm(const real (&v)[1], const real (&w)[1], int nact = N) {  // nact: see ThreadPolicy::norm
        if constexpr (HasLaneNorm<PROB>::value) return user_norm(v[0], w[0]);
        const real p = v[0] * w[0];
        const real p2 = p * p;
#if BVODE_STRICT
        real sum = BV_R(0.0);
#pragma unroll
        for (int i = 0; i < N; i++) sum = sum + shfl_d(p2, i);
        return sqrt(sum / (real)nact);
#else
        real sum = (lane() < N) ? p2 : BV_R(0.0);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum = sum + shfl_xor_d(sum, m);
        return sqrt((nact == N) ? sum * (BV_R(1.0) / (real)N) : sum / (real)nact);
#endif
    }
#if !BVODE_STRICT
    BV_DEV static real msq(const real (&v)[1], const real (&w)[1], int nact = N) {
        if constexpr (HasLaneNorm<PROB>::value) {  // a user norm is only known as a norm
            const real nv = user_norm(v[0], w[0]);
            return nv * nv;
        }
        const real p = v[0] * w[0];
        real sum = (lane() < N) ? p * p : BV_R(0.0);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum = sum + shfl_xor_d(sum, m);
        return (nact == N) ? sum * (BV_R(1.0) / (real)N) : sum / (real)nact;
    }
#endif
    BV_DEV static bool any_le_zero(const real (&v)[1]) {
        return __any_sync(kFull, (lane() < N) && (v[0] <= BV_R(0.0))) != 0;
    }
    // dvhin's upper bound loop, M:1784-1790: sequential in i because hub feeds the test
    BV_DEV static real hin_upper_bound(real hub, const real (&y0)[1], const real (&ydot)[1],
                                         const Tol &tol) {
        const real delyi = BV_R(0.1) * fabs(y0[0]) + tol.a;
        const real afi = fabs(ydot[0]);
#pragma unroll 1
        for (int i = 0; i < N; i++) {
            const real d = shfl_d(delyi, i), a = shfl_d(afi, i);
            if (a * hub > d) hub = d / a;
        }
        return hub;
    }
    BV_DEV static void f(real t, const real (&y)[1], real (&ydot)[1], const Par &par) {
        ydot[0] = PROB::f_lane(t, y[0], par, lane());
        if (lane() >= N) ydot[0] = BV_R(0.0);
    }
    // compute_imxer, M:1694-1706: first index of the largest |acor*ewt| (strict '<')
    BV_DEV static int imxer(const real (&acor)[1], const real (&ewt)[1]) {
        real v = (lane() < N) ? fabs(acor[0] * ewt[0]) : -BV_R(1.0);
        int idx = lane();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const real ov = shfl_xor_d(v, m);
            const int oi = __shfl_xor_sync(kFull, idx, m);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        return (v > BV_R(0.0)) ? idx + 1 : 1;
    }
    // idamax (BL:311-359) over the lanes with `act`: the lowest such lane holding the largest |x|.
    // |x| of finite doubles order like their bit patterns, so two 32-bit warp reductions and one
    // ballot replace the five-round shuffle tournament of argmax_abs below.
    BV_DEV static unsigned max_abs_ballot(real x, bool act) {
        const unsigned hi = act ? ((unsigned)__double2hiint(x) & 0x7fffffffu) : 0u;
        const unsigned lo = (unsigned)__double2loint(x);
        const unsigned mhi = __reduce_max_sync(kFull, hi);
        const bool c1 = act && (hi == mhi);
        const unsigned mlo = __reduce_max_sync(kFull, c1 ? lo : 0u);
        return __ballot_sync(kFull, c1 && (lo == mlo));
    }
    // idamax over lanes [lo, hi]: returns the lane of the first maximum of |x|
    BV_DEV static int argmax_abs(real x, int lo, int hi) {
        const int ln = lane();
        real v = (ln >= lo && ln <= hi) ? fabs(x) : -BV_R(1.0);
        int idx = ln;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const real ov = shfl_xor_d(v, m);
            const int oi = __shfl_xor_sync(kFull, idx, m);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        return idx;
    }
    // ---- the pivoted inverse of N rows held one per lane (fast builds of the dense and the band policy; the
    // method is described at WarpDensePolicy::dginv) ----------------------------------------------------------
    // pa: the N rows (stride LD doubles, 16-byte aligned); perma: perm[32] (ints); bufa: 32 doubles of scratch;
    // ipv: lane k receives perm[k].  N = PROB::N.
    template <int LD>
    BVODE_DGINV_INLINE static int dginv_rows(const unsigned pa, const unsigned perma, const unsigned bufa, int &ipv) {
        constexpr int N = PROB::N;
        // The lane's row lives in REGISTERS for the whole inversion, rotated by one entry per step so that
        // "column k" is always r[0] and every register index is static inside a run-time loop over k:
        // step k reads r[0] (the pivot candidate), lane l (the pivot row) publishes its row as it stands to a
        // 32-double buffer in shared memory, and every lane forms r'[j] = r[j+1] + t * P[j+1] from 16-byte
        // broadcast loads of that buffer; the entry of column k itself (the inverse's column) goes to the
        // end, r'[N-1].  After N steps the rotation is complete and every row is scaled by 1 / its pivot.
        // Shared-memory traffic per step: one single-lane store and one broadcast load of the pivot row,
        // instead of a load and a store of every lane's own row (the shared-memory data pipe is what bounds
        // the LINPACK form of this kernel, ncu r1o).
        const int ln = lane();
        const bool live = (N == 32) || (ln < N);
        const unsigned rowa = pa + (live ? ln : 0) * (LD * 8);
        constexpr int NE = (N + 1) & ~1;  // entries moved in pairs; r[N] (N odd) is a dummy
        real r[NE];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NE; j += 2) lds_d2(rowa + j * 8, r[j], r[j + 1]);
        bool unused = live;  // this lane's row has not been a pivot row yet
        int info = 0;
        real myrp = BV_R(1.0);  // 1 / (the pivot this lane's row supplied)
#pragma unroll 1
        for (int k = 0; k < N && info == 0; k++) {
            const real v = r[0];
#if BVODE_DGINV_EARLY_RCP
            const real rv = qrcp(v);  // off the ballot -> shuffle chain: every lane inverts its own candidate
#endif
            const unsigned b = max_abs_ballot(v, unused);
            const int l = __ffs((int)b) - 1;  // ties: the lowest row
            const real piv = shfl_d(v, l);
            if (ln == k) ipv = l;
            sts_i(perma + k * 4, l);
            if (piv == BV_R(0.0)) {
                info = k + 1;
            } else {
#if BVODE_DGINV_EARLY_RCP
                const real rp = shfl_d(rv, l);
#else
                const real rp = qrcp(piv);
#endif
                const bool mine = (ln == l);
                // The pivot row is NOT scaled during the elimination: row operations commute with a scaling of
                // the pivot row as long as the multipliers are formed from the entries as they stand
                // (t = -v / pivot), so every row is scaled once, by 1 / its own pivot, after the last step.
                // The single-lane section of a step is then just the publication of the row as it lies in the
                // registers -- N / 2 16-byte stores of aligned pairs, no arithmetic, no moves (150 single-lane
                // instructions per step in r2j, 63 with the row scaled and rotated by the pivot lane).  Every
                // lane forms
                //     r'[j] = r[j+1] + t * P[j+1]  (j < N-1),    r'[N-1] = t        (pivot row: t = 0, r'[N-1] = 1)
                // i.e. the buffer is read one entry ahead of the register it lands in.
                if (mine) {
#pragma unroll
                    for (int j = 0; j < NE; j += 2) sts_d2(bufa + j * 8, r[j], r[j + 1]);
                    unused = false;
                    myrp = rp;
                }
                __syncwarp();
                const real t = mine ? BV_R(0.0) : -(v * rp);
                // kB 16-byte loads of the buffer are in flight before their fmas: with one (the compiler's choice
                // under register pressure: every load lands in the same register quad) a step is N / 2 exposed
                // shared-memory latencies
                constexpr int kB = BVODE_DGINV_BATCH;
#pragma unroll
                for (int j0 = 0; j0 < NE; j0 += 2 * kB) {
                    real p[2 * kB];
#pragma unroll
                    for (int q = 0; q < kB; q++)
                        if (j0 + 2 * q < NE) lds_d2(bufa + (j0 + 2 * q) * 8, p[2 * q], p[2 * q + 1]);
#if BVODE_DGINV_FENCE
                    if (kB > 1) __syncwarp();  // ptxas otherwise sinks each load to its first use again
#endif
#pragma unroll
                    for (int q = 0; q < kB; q++) {
                        const int j = j0 + 2 * q;
                        if (j < NE) {
                            if (j >= 1 && j <= N - 1) r[j - 1] = fma(t, p[2 * q], r[j]);
                            if (j + 1 <= N - 1) r[j] = fma(t, p[2 * q + 1], r[j + 1]);
                        }
                    }
                }
                r[N - 1] = mine ? BV_R(1.0) : t;
                __syncwarp();
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < N; j++) r[j] = r[j] * myrp;
#pragma unroll
            for (int j = 0; j < NE; j += 2) sts_d2(rowa + j * 8, r[j], r[j + 1]);
        }
        __syncwarp();
        return info;
    }
    // x = P^-1 b from the rows dginv_rows left at pa (veca: 32 doubles of scratch; ipv as returned by dginv_rows)
    template <int LD>
    BV_DEV static void solve_inverse_rows(const unsigned pa, const unsigned veca, const int ipv, real (&x)[1]) {
        constexpr int N = PROB::N;
        const bool live = (N == 32) || (lane() < N);
        const unsigned rowa = pa + (live ? lane() : 0) * (LD * 8);
        // b[perm[j]] to lane j, then to shared memory: every lane needs all of them
        const real bp = shfl_d(x[0], ipv);
        __syncwarp();
        sts_d(veca + lane() * 8, live ? bp : BV_R(0.0));
        __syncwarp();
        real s0 = BV_R(0.0), s1 = BV_R(0.0);
        constexpr int NP = N & ~1;
#pragma unroll 8
        for (int j = 0; j < NP; j += 2) {
            real a0, a1, b0, b1;
            lds_d2(rowa + j * 8, a0, a1);
            lds_d2(veca + j * 8, b0, b1);
            s0 = fma(a0, b0, s0);
            s1 = fma(a1, b1, s1);
        }
        if (N & 1) s0 = fma(lds_d(rowa + (N - 1) * 8), lds_d(veca + (N - 1) * 8), s0);
        const real s = live ? s0 + s1 : BV_R(0.0);
        x[0] = shfl_d(s, ipv);  // x_k is the sum of row perm[k]
    }


    // FD increment of component `lane`, M:2966 / M:3065: r = max(srur*|y|, r0/ewt)
    BV_DEV static real fd_r0(const Ctl &c, const real (&savf)[1], const real (&ewt)[1], int nact = N) {
        const real fac = norm(savf, ewt, nact);
        real r0 = BV_R(1000.0) * fabs(c.h) * kUround * (real)nact * fac;
        if (r0 == BV_R(0.0)) r0 = BV_R(1.0);
        return r0;
    }
};

// =========================================================================== dense ====
// Row i of the Newton matrix belongs to lane i for the whole factorisation; the matrix is in
// shared memory, element (i, j) at p[i*LDP + j] with LDP odd, so "every lane reads its own
// row's column j" and "every lane reads element (l, j)" (a broadcast) are both conflict-free.
//
// Pivoting is IMPLICIT: LINPACK's row interchanges (LP:96-110) move the active parts of two rows
// and leave the multipliers where they are; dgesl repeats the same interchanges on b.  Here no
// row ever moves.  Each lane keeps its row and a position `pos` (the index LINPACK's storage
// would give that row); step k picks the pivot lane l = perm[k] and exchanges the POSITIONS of
// the two rows.  A row pivoted at step k has pos = k for good, so "rows below the pivot" is
// pos > k, a row's multipliers stay in its own storage, and b_i travels with row i in dgesl.
// Every product and sum is the one LINPACK forms, on the same operands, in the same order per
// element -- the factors, the solution and the tie rule (first maximum in LINPACK's row order,
// BL:334) are bit-identical; only where a number is stored differs.  This removes the swap
// traffic and the two lane-exchanges per solve step of the first version.
//
// All loops over columns are run-time loops: the very first version kept each row in registers
// with fully unrolled loops (~20k instructions) and was bound by instruction fetch (ncu: 33
// cycles of no-instruction stall per issued instruction,
// profiles/r01_ncu_warp_dense_mech32_v1.txt).
template <class PROB, int MITER_, bool SAVEJ_, int METH_ = 2>
struct WarpDensePolicy : WarpCommon<PROB, METH_> {
    using Base = WarpCommon<PROB, METH_>;
    static constexpr int METH = METH_;
    static constexpr int kLmax = Base::kLmax;
    using Prob = PROB;
    using typename Base::GlobalVec;
    using typename Base::Hot;
    using Base::lane;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = 1;
    static constexpr int MITER = MITER_;
    static constexpr bool SAVEJ = SAVEJ_;
    static constexpr bool kBand = false;
    static constexpr bool kPlain = false;
#ifndef BVODE_DENSE_MINBLOCKS
#define BVODE_DENSE_MINBLOCKS 3
#endif
    static constexpr int kMinBlocks = BVODE_DENSE_MINBLOCKS;  // CTAs per SM (see bvode_kernels.cuh)
    // Fast build: the Newton matrix is INVERTED at dvjac time (Gauss-Jordan with partial pivoting, the
    // pivot sequence of dgefa) and dvsol is one matrix-vector product -- see dginv / solve below.  The
    // strict build keeps LINPACK's factors and solves (bit for bit).
#ifndef BVODE_DENSE_INVERSE
#define BVODE_DENSE_INVERSE 1
#endif
    static constexpr bool kInverse = (BVODE_STRICT == 0) && (BVODE_DENSE_INVERSE != 0);
    // row stride in doubles.  LU: odd (own-row and broadcast 8-byte accesses are conflict-free).  Inverse:
    // = 2 (mod 4), so that rows are 16-byte aligned and the 8 lanes of a quarter-warp hit disjoint banks
    // with 16-byte accesses (the byte stride is 16 * odd).
    static constexpr int LDP = kInverse ? N + ((6 - (N & 3)) & 3) : (N | 1);
    static constexpr int kPermOff = N * LDP + ((N * LDP) & 1);  // perm[32] (ints), 16-byte aligned
    static constexpr int kVecOff = kPermOff + 16;               // inverse: the permuted right-hand side, 32 doubles
    static constexpr int kBlockDoubles = kVecOff + (kInverse ? 32 : 0);
    static_assert(MITER == 1 || MITER == 2, "dense policy");

    struct Lin {
        real *p;      // shared memory, N rows x LDP: P = I - h*rl1*J, then its LU factors.
                        // Fast build: the pivot entry of each row holds 1/u_kk after dgefa.
        int *perm;      // shared memory: perm[k] = lane (row) that is the pivot of step k
        GlobalVec wjs;  // row `lane` of the saved Jacobian, element j at slot j (mf > 0)
        int ipv;        // lane k holds perm[k]
        int pos;        // position of this lane's row in LINPACK's storage order
    };
    BV_DEV static bool real_lane() { return (N == 32) || (lane() < N); }
    template <class ENG>
    BV_DEV static void lin_init(ENG &e) {
        e.lin.ipv = lane();
        e.lin.pos = real_lane() ? lane() : 64;
        e.lin.perm[lane()] = lane();
        __syncwarp();
    }
    static int smem_doubles(int, int, bool) { return kBlockDoubles; }
    // per-lane double slots of one system, in visit order: yh[6], acsav, acor, wjs[N]*
    static constexpr int kLaneSlotWjs = kLmax + 2;
    static constexpr int kLaneSlots = kLmax + 2 + (SAVEJ ? N : 0);
    static constexpr int kLaneIntSlots = 2;
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) { v.dv(e.acsav.v); }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
        Lin &l = e.lin;
        if (SAVEJ) v.skip_dv(N);
        v.block(l.p, kBlockDoubles);
        v.iv(l.ipv);
        v.iv(l.pos);
    }

    // idamax (BL:311-359) over the rows that are still below the diagonal (`act`), magnitudes
    // compared as integers (|v| of a finite double orders like its bit pattern): two warp
    // reductions instead of a five-round shuffle tournament.  Several rows of the same largest
    // magnitude: the first in LINPACK's row order wins.
    BV_DEV static int pivot_lane(real v, bool act, int pos) {
        unsigned b = Base::max_abs_ballot(v, act);
        if (b & (b - 1)) {  // warp-uniform
            const bool c2 = (b >> lane()) & 1u;
            const unsigned mp = __reduce_min_sync(kFull, c2 ? (unsigned)pos : 0xffffu);
            b = __ballot_sync(kFull, c2 && ((unsigned)pos == mp));
        }
        return __ffs((int)b) - 1;
    }

    // ---- dgefa, LP:46-120, rows across lanes ----------------------------------------
    // Two pivot steps per pass over the rows (a rank-2 update): the multipliers of step k are
    // applied to column k+1 alone, the pivot of step k+1 is chosen from it, and every later
    // column j then takes  row(j) += p1(j)*m  and  row(j) += p2'(j)*m1  in one load / store of the
    // row element, where p2'(j) = p2(j) + p1(j)*m(l1) is the second pivot row AFTER step k --
    // formed by every lane from two broadcast loads exactly as the lane that owns that row forms it.
    // Each element sees the same two operations, on the same operands, in the same order as in
    // LINPACK's column-by-column elimination; the shared-memory traffic per element update drops
    // from 5 wavefronts (broadcast + row load + row store) to 3.
    BV_DEV static int dgefa(Lin &lin) {
        const int ln = lane();
        const bool live = real_lane();
        const unsigned pa = sm_addr(lin.p);                 // the matrix
        const unsigned rowa = pa + (live ? ln : 0) * (LDP * 8);  // this lane's row (lanes >= N never write)
        const unsigned perma = pa + kPermOff * 8;
        int pos = live ? ln : 64;
        int info = 0;
        // idamax + the interchange of LP:96-110 (the two rows exchange positions) for column k,
        // v = this lane's entry of that column
        auto pick = [&](int k, real v, int &l, real &alk) {
            l = pivot_lane(v, live && (pos >= k), pos);
            alk = shfl_d(v, l);
            const int pl = __shfl_sync(kFull, pos, l);
            if (ln == k) lin.ipv = l;
            sts_i(perma + k * 4, l);  // the same value from every lane
            if (pos == k) pos = pl;
            if (ln == l) pos = k;
        };
        __syncwarp();
        int k = 0;
#pragma unroll 1
        while (k < N - 1) {
            const unsigned rk = rowa + k * 8;
            const real v = lds_d(rk);
            int l;
            real alk;
            pick(k, v, l, alk);
            int adv = 1;
            if (alk == BV_R(0.0)) {
                info = k + 1;
            } else {
                const bool below = live && (pos > k);
#if BVODE_STRICT
                const real t = -BV_R(1.0) / alk;
                const real m = t * v;  // the multipliers (LP:102-104)
                if (below) sts_d(rk, m);
#else
                // fast build: the pivot row keeps 1/u_kk (dgesl multiplies); the rows ALREADY
                // pivoted scale their entry of column k by -1/u_kk as well, so that back
                // substitution is b_i += b_k * (-u_ik/u_kk) with the un-normalised b_k -- one
                // shuffle and one fma per step on the dependent chain, x_k = b_k/u_kk at the end
                const real rp = qrcp(alk);
                const real t = -rp;
                const real m = t * v;
                if (live) sts_d(rk, (ln == l) ? rp : m);
#endif
                const unsigned pr0 = pa + l * (LDP * 8);  // the pivot row: broadcasts
                bool pair = false;
                if (k + 1 < N - 1) {
                    // ---- column k+1 after step k, and its pivot
                    adv = 2;
                    real w = lds_d<8>(rk);
                    const real p1k = lds_d<8>(pr0 + k * 8);
                    if (below) w = w + p1k * m;
                    int l1;
                    real alk1;
                    pick(k + 1, w, l1, alk1);
                    if (alk1 == BV_R(0.0)) {
                        info = k + 2;  // LP:87-88: nothing is eliminated in column k+1; step k is completed below
                    } else {
                        pair = true;
                        const bool below1 = live && (pos > k + 1);
#if BVODE_STRICT
                        const real t1 = -BV_R(1.0) / alk1;
                        const real m1 = t1 * w;
                        if (below1) sts_d<8>(rk, m1);
                        if (ln == l1) sts_d<8>(rk, w);
#else
                        const real rp1 = qrcp(alk1);
                        const real t1 = -rp1;
                        const real m1 = t1 * w;
                        if (live) sts_d<8>(rk, (ln == l1) ? rp1 : m1);
#endif
                        const real ml1 = shfl_d(m, l1);  // multiplier of the second pivot row in step k
                        const unsigned pr1 = pa + l1 * (LDP * 8);
                        // every lane runs the loop (the __syncwarp between the loads and the stores
                        // orders the other lanes' reads of row l1 before its owner's update of it)
                        unsigned ar = rk + 16, a1 = pr0 + (k + 2) * 8, a2 = pr1 + (k + 2) * 8;
                        int cnt = N - 2 - k;
#pragma unroll 1
                        for (; cnt >= 4; cnt -= 4, ar += 32, a1 += 32, a2 += 32) {
                            const real p0 = lds_d<0>(a1), p1 = lds_d<8>(a1), p2 = lds_d<16>(a1), p3 = lds_d<24>(a1);
                            const real q0 = lds_d<0>(a2), q1 = lds_d<8>(a2), q2 = lds_d<16>(a2), q3 = lds_d<24>(a2);
                            real r0 = lds_d<0>(ar), r1 = lds_d<8>(ar), r2 = lds_d<16>(ar), r3 = lds_d<24>(ar);
                            __syncwarp();
                            r0 = r0 + p0 * m; r1 = r1 + p1 * m; r2 = r2 + p2 * m; r3 = r3 + p3 * m;
                            if (below1) {
                                r0 = r0 + (q0 + p0 * ml1) * m1;
                                r1 = r1 + (q1 + p1 * ml1) * m1;
                                r2 = r2 + (q2 + p2 * ml1) * m1;
                                r3 = r3 + (q3 + p3 * ml1) * m1;
                            }
                            if (below) {
                                sts_d<0>(ar, r0); sts_d<8>(ar, r1); sts_d<16>(ar, r2); sts_d<24>(ar, r3);
                            }
                        }
#pragma unroll 1
                        for (; cnt > 0; cnt--, ar += 8, a1 += 8, a2 += 8) {
                            const real p0 = lds_d(a1), q0 = lds_d(a2);
                            real r0 = lds_d(ar);
                            __syncwarp();
                            r0 = r0 + p0 * m;
                            if (below1) r0 = r0 + (q0 + p0 * ml1) * m1;
                            if (below) sts_d(ar, r0);
                        }
                    }
                }
                if (!pair && below) {
                    // one step: row(j) += pivot_row(j) * m, j > k (LP:105-118); loads first, then stores
                    unsigned ar = rk + 8, ap = pr0 + (k + 1) * 8;
                    int cnt = N - 1 - k;
#pragma unroll 1
                    for (; cnt >= 4; cnt -= 4, ar += 32, ap += 32) {
                        const real p0 = lds_d<0>(ap), p1 = lds_d<8>(ap), p2 = lds_d<16>(ap), p3 = lds_d<24>(ap);
                        const real r0 = lds_d<0>(ar), r1 = lds_d<8>(ar), r2 = lds_d<16>(ar), r3 = lds_d<24>(ar);
                        sts_d<0>(ar, r0 + p0 * m);
                        sts_d<8>(ar, r1 + p1 * m);
                        sts_d<16>(ar, r2 + p2 * m);
                        sts_d<24>(ar, r3 + p3 * m);
                    }
#pragma unroll 1
                    for (; cnt > 0; cnt--, ar += 8, ap += 8) sts_d(ar, lds_d(ar) + lds_d(ap) * m);
                }
            }
            __syncwarp();
            k += adv;
        }
        {   // the row that is left is the last pivot
            const int l = __ffs((int)__ballot_sync(kFull, live && (pos == N - 1))) - 1;
            if (ln == N - 1) lin.ipv = l;
            sts_i(perma + (N - 1) * 4, l);
            const real v = lds_d<(N - 1) * 8>(rowa);
            const real alk = shfl_d(v, l);
            if (alk == BV_R(0.0)) info = N;
#if !BVODE_STRICT
            const real rp = qrcp(alk);
            if (live) sts_d<(N - 1) * 8>(rowa, (ln == l) ? rp : -rp * v);  // column N-1 of U, scaled like the others
#endif
        }
        lin.pos = pos;
        __syncwarp();
        return info;
    }

    // ---- fast build: P^-1 by Gauss-Jordan elimination with partial pivoting, in place ------------------
    // Step k takes the pivot of column k among the rows not used yet -- the same entries dgefa (LP:46-120)
    // looks at, so the pivot sequence and the singularity test (LP:87-88) are dgefa's -- scales that row by
    // 1/pivot and eliminates column k from ALL other rows.  Rows never move (implicit pivoting as above);
    // column k of the storage ends up holding column perm[k] of the accumulated row operations E, where
    // E*P = Pi and Pi has its 1 of row perm[k] in column k.  Hence for P x = b:
    //     x_k = sum_j a[perm[k]][j] * b[perm[j]]                                   (solve below)
    // Why: with one row per lane dgesl is 2N-1 dependent "shuffle -> fma" steps (63 for N = 32) and half of
    // the lanes idle in every one of them; the product is N independent fmas per lane with 16-byte
    // shared-memory loads.  A chord iteration needs its linear solve only to a few digits (the iteration
    // matrix is up to 30 % stale by design, M:2736-2741), and the pivoted inverse delivers
    // cond(P)*uround; the batch statistics stay inside the reference's +-2 % (tests).  Every lane works in
    // every elimination step, so the inversion costs about the instructions of the factorisation it
    // replaces (N^3 lane-fmas at full occupancy against N^3/3 at a third of it).
    BV_DEV static int dginv(Lin &lin) {
        const unsigned pa = sm_addr(lin.p);
        const int info = Base::template dginv_rows<LDP>(pa, pa + kPermOff * 8, pa + kVecOff * 8, lin.ipv);
        lin.pos = 0;
        return info;
    }
    template <class ENG>
    BV_DEV static int solve_inverse(ENG &e, real (&x)[1]) {
        const unsigned pa = sm_addr(e.lin.p);
        Base::template solve_inverse_rows<LDP>(pa, pa + kVecOff * 8, e.lin.ipv, x);
        return 0;
    }

    // ---- dgesl job = 0, LP:206-225 -------------------------------------------------------
    // One step: t = b of the pivot row of column k; rows taking part add t * (their entry of
    // column k).  FWD: rows below the pivot (multipliers); !FWD: rows above it.
    template <bool FWD, int OFF>
    BV_DEV static void solve_step(real &b, unsigned rk, int l, int pos, int k, bool live) {
        const real m = lds_d<OFF * 8>(rk);
        const bool part = live && (FWD ? (pos > k + OFF) : (pos < k + OFF));
#if BVODE_STRICT
        if (!FWD && pos == k + OFF) b = b / m;
        const real t = FWD ? shfl_d(b, l) : -shfl_d(b, l);
        if (part) b = b + t * m;
#else
        const real t = shfl_d(b, l);
        b = b + t * (part ? m : BV_R(0.0));  // the select is off the dependent chain (measured faster
                                       // here than one predicated fma: 162k vs 157k systems/s)
#endif
    }
    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[1]) {
        if constexpr (kInverse) return solve_inverse(e, x);
        const Lin &lin = e.lin;
        const bool live = real_lane();
        const unsigned pa = sm_addr(lin.p);
        const unsigned rowa = pa + (live ? lane() : 0) * (LDP * 8);
        const unsigned perma = pa + kPermOff * 8;
        const int pos = lin.pos;
        real b = x[0];
        // forward elimination, k = 0 .. N-2
        int k = 0;
#pragma unroll 1
        for (; k + 3 < N - 1; k += 4) {
            const unsigned rk = rowa + k * 8;
            int l0, l1, l2, l3;
            lds_i4(perma + k * 4, l0, l1, l2, l3);
            // (a four-step look-ahead -- the next pivot rows' b formed redundantly by every lane
            // from broadcast entries, one shuffle latency per four steps -- was measured SLOWER,
            // 149k vs 165k systems/s: the kernel is bound by instruction and shared-memory
            // throughput, not by this chain)
            solve_step<true, 0>(b, rk, l0, pos, k, live);
            solve_step<true, 1>(b, rk, l1, pos, k, live);
            solve_step<true, 2>(b, rk, l2, pos, k, live);
            solve_step<true, 3>(b, rk, l3, pos, k, live);
        }
#pragma unroll 1
        for (; k < N - 1; k++) solve_step<true, 0>(b, rowa + k * 8, lds_i(perma + k * 4), pos, k, live);
        // back substitution, k = N-1 .. kend (fast build: column 0 has no row above it)
        constexpr int kend = BVODE_STRICT ? 0 : 1;
        k = N - 1;
        // (perm[] is read four entries at a time from 16-byte aligned addresses: single steps down to k = 3 mod 4)
#pragma unroll 1
        for (; (k & 3) != 3 && k >= kend; k--) solve_step<false, 0>(b, rowa + k * 8, lds_i(perma + k * 4), pos, k, live);
#pragma unroll 1
        for (; k - 3 >= kend; k -= 4) {
            const unsigned rk = rowa + k * 8;
            int l0, l1, l2, l3;
            lds_i4(perma + (k - 3) * 4, l3, l2, l1, l0);
            solve_step<false, 0>(b, rk, l0, pos, k, live);
            solve_step<false, -1>(b, rk, l1, pos, k, live);
            solve_step<false, -2>(b, rk, l2, pos, k, live);
            solve_step<false, -3>(b, rk, l3, pos, k, live);
        }
#pragma unroll 1
        for (; k >= kend; k--) solve_step<false, 0>(b, rowa + k * 8, lds_i(perma + k * 4), pos, k, live);
#if !BVODE_STRICT
        if (live) b = b * lds_d(rowa + pos * 8);  // x_k = b_k / u_kk
#endif
        x[0] = shfl_d(b, lin.ipv);  // component k is on the lane that was pivot k
        return 0;
    }

    // ---- dvjac, miter 1 / 2 (M:2930-3001) ---------------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        Lin &l = e.lin;
        const int ln = lane();
        real *row = l.p + ln * LDP;
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        int jok = SAVEJ ? 1 : -1;
        if (SAVEJ) {
            if (e.hot.si(SI_NST) == 0 || e.hot.si(SI_NST) > e.hot.si(SI_NSLJ) + 50) jok = -1;
            if (c.icf == 1 && c.drc < BV_R(0.2)) jok = -1;
            if (c.icf == 2) jok = -1;
        }
        __syncwarp();
        if (jok == -1) {
            e.hot.si_add(SI_NJE, 1);
            e.hot.si(SI_NSLJ) = e.hot.si(SI_NST);
            c.jcur = 1;
            if constexpr (MITER == 1) {
                if (ln < N) {
#pragma unroll 1
                    for (int j = 0; j < N; j++) row[j] = BV_R(0.0);
                }
                if constexpr (HasJacDenseFull<PROB>::value)  // the reference's jac(neq, t, y, ml, mu, pd, nrowpd)
                    PROB::jac_dense_full(c.tn, e.y[0], l.p, LDP, e.par, ln);
                else
                    PROB::jac_row(c.tn, e.y[0], row, e.par, ln);  // every lane calls (it may shuffle); lanes >= N must not write
                if (e.nact != N) {  // neq was reduced: the Jacobian of the first nact equations in their own unknowns
                    __syncwarp();
                    if (ln < N) {
#pragma unroll 1
                        for (int j = 0; j < N; j++)
                            if (ln >= e.nact || j >= e.nact) row[j] = BV_R(0.0);
                    }
                }
            } else {
                // finite differences, M:2959-2976: column j for all rows at once
                const real r0 = Base::fd_r0(c, e.savf, e.ewt, e.nact);
                const real srur = sqrt(kUround);
                const real rmine = dmax2(srur * fabs(e.y[0]), qdiv(r0, e.ewt[0]));
                const real ysave = e.y[0];
#pragma unroll 1
                for (int j = 0; j < N; j++) {
                    if (j < e.nact) {  // (warp-uniform)
                        const real r = shfl_d(rmine, j);
                        if (ln == j) e.y[0] = ysave + r;
                        const real fac = qrcp(r);
                        e.feval(c.tn, e.y, e.acor);
                        if (ln < N) row[j] = (e.acor[0] - e.savf[0]) * fac;
                        e.y[0] = ysave;
                    } else if (ln < N) {
                        row[j] = BV_R(0.0);
                    }
                }
                e.hot.si_add(SI_NFE, e.nact);
            }
            if (SAVEJ && ln < N) {
#pragma unroll 4
                for (int j = 0; j < N; j++) l.wjs.set(j, row[j]);
            }
        } else {
            c.jcur = 0;
            if (SAVEJ && ln < N) {
                // (several loads in flight: the saved copy lives in HBM, one line per column)
#pragma unroll 8
                for (int j = 0; j < N; j++) row[j] = l.wjs.get(j);
            }
        }
        const real con = -hrl1;
        if (ln < N) {
            if constexpr (kInverse) {
                // rows are 16-byte aligned in this layout: pairs, four loads in flight (the element-by-element loop
                // exposed one shared-memory latency per entry: 2.5 % of the samples of the kernel, ncu r2o)
                const unsigned ra = sm_addr(row);
                constexpr int NE = (N + 1) & ~1;  // (N odd: the pad entry of the row is scaled as well)
#pragma unroll 4
                for (int j = 0; j < NE; j += 2) {
                    real a0, a1;
                    lds_d2(ra + j * 8, a0, a1);
                    sts_d2(ra + j * 8, con * a0, con * a1);
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < N; j++) row[j] = con * row[j];
            }
            row[ln] = row[ln] + BV_R(1.0);
        }
        e.hot.si_add(SI_NLU, 1);
        const int ier = factor(l);
        if (ier != 0) ierpj = 1;
    }
    // dvjac's call of dgefa (M:2998): LINPACK's factors (strict build) or the inverse (fast build)
    BV_DEV static int factor(Lin &l) {
        if constexpr (kInverse) return dginv(l);
        else return dgefa(l);
    }
};

// =========================================================================== plain ====
// No Newton matrix: functional iteration (miter = 0, M:2822-2837) or the diagonal Jacobian
// approximation (miter = 3: dvjac M:3004-3028, dvsol M:3164-3181), one diagonal entry per lane.
template <class PROB, int MITER_, int METH_ = 2>
struct WarpPlainPolicy : WarpCommon<PROB, METH_> {
    using Base = WarpCommon<PROB, METH_>;
    using Prob = PROB;
    using Base::lane;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = 1;
    static constexpr int MITER = MITER_;
    static constexpr int METH = METH_;
    static constexpr int kLmax = Base::kLmax;
    static constexpr bool SAVEJ = false;
    static constexpr bool kBand = false;
    static constexpr bool kPlain = true;
    static constexpr int kMinBlocks = 8;
    static_assert(MITER == 0 || MITER == 3, "plain policy");

    struct Lin {
        real diag;   // wm(2 + lane): inverted diagonal of P (miter = 3)
        real phrl1;  // wm(2): h*rl1 at the last update (replicated on all lanes)
    };
    template <class ENG>
    BV_DEV static void lin_init(ENG &e) { e.lin.diag = BV_R(1.0); e.lin.phrl1 = BV_R(0.0); }
    static int smem_doubles(int, int, bool) { return 0; }
    static constexpr int kLaneSlots = kLmax + 2 + 2;  // yh, acsav, acor, diag, phrl1
    static constexpr int kLaneIntSlots = 1;
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) { v.dv(e.acsav.v); }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
        v.dv(e.lin.diag);
        v.dv(e.lin.phrl1);
        int dummy = 0;
        v.iv(dummy);
    }

    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[1]) {
        Lin &l = e.lin;
        const real hrl1 = e.c.h * e.c.rl1;
        const real phrl1 = l.phrl1;
        l.phrl1 = hrl1;
        if (hrl1 != phrl1) {
            const real r = hrl1 / phrl1;
            const real di = BV_R(1.0) - r * (BV_R(1.0) - BV_R(1.0) / l.diag);
            const bool zero = (lane() < N) && (fabs(di) == BV_R(0.0));
            // the reference stops at the first zero: entries before it are already updated
            const unsigned zmask = __ballot_sync(kFull, zero);
            const int first = zmask ? (__ffs((int)zmask) - 1) : 32;
            if (lane() < first) l.diag = BV_R(1.0) / di;
            if (zmask) return 1;
        }
        x[0] = l.diag * x[0];
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
        e.y[0] = e.y[0] + r * (c.h * e.savf[0] - e.hot.yh(1, 0));
        real w[1];
        e.feval(c.tn, e.y, w);
        e.hot.si_add(SI_NFE, 1);
        const real r0 = c.h * e.savf[0] - e.hot.yh(1, 0);
        const real di = BV_R(0.1) * r0 - c.h * (w[0] - e.savf[0]);
        const bool big = (lane() < N) && (fabs(r0) >= kUround / e.ewt[0]);
        const bool zero = big && (fabs(di) == BV_R(0.0));
        const unsigned zmask = __ballot_sync(kFull, zero);
        const int first = zmask ? (__ffs((int)zmask) - 1) : 32;
        if (lane() < first) l.diag = big ? BV_R(0.1) * r0 / di : BV_R(1.0);
        if (zmask) ierpj = 1;
    }
};

// ============================================================================ band ====
// LINPACK band storage (LP:230-260) in shared memory: abd(i, j), i = 1..meband, j = 1..n at
// abd[(j-1)*ld + (i-1)], meband = 2*ml + mu + 1, ld odd (bank-conflict-free column stride).
// The saved Jacobian is the compact mband x n copy (M:3048, dacopy) in shared memory too.
constexpr int kBandMaxRows = 2 * 12 + 12 + 1;  // sized at launch from ml, mu

template <class PROB, int MITER_, bool SAVEJ_, int METH_ = 2>
struct WarpBandPolicy : WarpCommon<PROB, METH_> {
    using Base = WarpCommon<PROB, METH_>;
    static constexpr int METH = METH_;
    static constexpr int kLmax = Base::kLmax;
    using Prob = PROB;
    using Base::lane;
    static constexpr int N = PROB::N;
    static constexpr int NLOC = 1;
    static constexpr int MITER = MITER_;
    static constexpr bool SAVEJ = SAVEJ_;
    static constexpr bool kBand = true;
    static constexpr bool kPlain = false;
#ifndef BVODE_BAND_MINBLOCKS
#define BVODE_BAND_MINBLOCKS 4
#endif
    static constexpr int kMinBlocks = BVODE_BAND_MINBLOCKS;
    static_assert(MITER == 4 || MITER == 5, "band policy");
    // Fast build: like the dense policy, the Newton matrix is INVERTED at dvjac time (WarpCommon::dginv_rows on
    // the dense rows expanded from the band array, one row per lane) and dvsol is one matrix-vector product.
    // dgbsl with b_i on lane i is 2n - 1 dependent "shuffle -> predicated fma" steps whatever the band width
    // (~11 instructions each with run-time ml / mu: 28 % of the samples of this kernel in ncu r2m, dgbfa another
    // 10 %); the product is ~60 instructions.  The band array keeps P = I - h*rl1*J UNFACTORED (it is what is
    // saved with the system's state); the inverse is scratch in shared memory and is rebuilt from it when a
    // system is loaded again.  The strict build keeps LINPACK's dgbfa / dgbsl (bit for bit).
#ifndef BVODE_BAND_INVERSE
#define BVODE_BAND_INVERSE 1
#endif
    static constexpr bool kInverse = (BVODE_STRICT == 0) && (BVODE_BAND_INVERSE != 0);
    static constexpr int LDI = N + ((6 - (N & 3)) & 3);  // row stride of the inverse: = 2 (mod 4) doubles, see WarpDensePolicy::LDP
    static constexpr int kInvPermOff = N * LDI + ((N * LDI) & 1);
    static constexpr int kInvVecOff = kInvPermOff + 16;
    static constexpr int kInvDoubles = kInverse ? kInvVecOff + 32 : 0;

    struct Lin {
        real *abd;   // shared memory, ld x N
        real *wjs;   // shared memory, mband x N (mf > 0)
        real *inv;   // shared memory, N x LDI + perm[32] + 32 doubles (fast build): the pivoted inverse of abd
        int ml, mu, ld;
        int ipv;       // lane k holds ipvt(k+1) as a 0-based row of the full matrix
    };
    template <class ENG>
    BV_DEV static void lin_init(ENG &e) { e.lin.ipv = lane(); }

    __host__ __device__ static int band_ld(int ml, int mu) { return (2 * ml + mu + 1) | 1; }
    // doubles of one system's SAVED image: the band array and the saved Jacobian
    __host__ __device__ static int image_doubles(int ml, int mu, bool savej) {
        const int ld = (2 * ml + mu + 1) | 1;
        return ld * N + (savej ? (ml + mu + 1) * N : 0);
    }
    // doubles of shared memory one system needs: the image (rounded to 16 bytes) + the scratch of the inverse
    __host__ __device__ static int smem_doubles(int ml, int mu, bool savej) {
        const int img = image_doubles(ml, mu, savej);
        return kInverse ? ((img + 1) & ~1) + kInvDoubles : img;
    }
    static constexpr int kLaneSlots = kLmax + 2;  // yh[6], acsav, acor; the band arrays follow
    static constexpr int kLaneIntSlots = 1;
    template <class ENG, class V>
    BV_DEV static void visit_acsav(ENG &e, V &v) { v.dv(e.acsav.v); }
    template <class ENG, class V>
    BV_DEV static void visit_lin(ENG &e, V &v) {
        Lin &l = e.lin;
        // band arrays: the whole shared-memory image of this system, cooperatively
        const int nb = l.ld * N + (SAVEJ ? (l.ml + l.mu + 1) * N : 0);
        v.block(l.abd, nb);
        v.iv(l.ipv);
    }
    // ---- fast build: expand the band array into one dense row per lane and invert (see kInverse above) ----
    // a(i, j) = abd(i - j + m, j), m = ml + mu + 1, for max(1, j - mu) <= i <= min(n, j + ml) (LP:230-260);
    // abd is left as it is.  Returns dgbfa's info (the pivot sequence is dgefa's = dgbfa's: the rows outside the
    // band hold zeros in the pivot column).
    BV_DEV static int invert(Lin &l) {
        const int ln = lane();
        const int ml = l.ml, mu = l.mu, m = ml + mu + 1;
        __syncwarp();
        if (ln < N) {
            const real *col = l.abd + (m - 1 + ln);  // abd[(j)*ld + (i - j + m - 1)], i = ln, 0-based j: advances by ld - 1
            real *row = l.inv + ln * LDI;
            for (int j = 0; j < LDI; j++, col += l.ld - 1)
                row[j] = (j < N && j >= ln - ml && j <= ln + mu) ? *col : BV_R(0.0);
        }
        __syncwarp();
        const unsigned pa = sm_addr(l.inv);
        return Base::template dginv_rows<LDI>(pa, pa + kInvPermOff * 8, pa + kInvVecOff * 8, l.ipv);
    }
    // factor (dvjac, the LU known-answer tests): LINPACK's band factors in place, or the inverse next to the band array
    BV_DEV static int factor(Lin &l) {
        if constexpr (kInverse) return invert(l);
        else return dgbfa(l);
    }

#define ABD(i, j) l.abd[((j)-1) * l.ld + ((i)-1)]
    // ---- dgbfa, LP:275-396 (1-based indices as in the reference) ------------------------
    BV_DEV static int dgbfa(Lin &l) {
        const int ln = lane();
        const int n = N, ml = l.ml, mu = l.mu;
        const int m = ml + mu + 1;
        int info = 0;
        // zero initial fill-in columns, LP:299-308: column jz = lane+1
        const int j0 = mu + 2, j1 = (n < m ? n : m) - 1;
        {
            const int jz = ln + 1;
            if (jz >= j0 && jz <= j1) {
                for (int i = m + 1 - jz; i <= ml; i++) ABD(i, jz) = BV_R(0.0);
            }
        }
        __syncwarp();
        int jz = j1, ju = 0;
        // shared addresses (bytes): entry (i, j) of abd is at base + ((j-1)*ld + (i-1))*8
        const unsigned base = sm_addr(l.abd);
        const int ld8 = l.ld * 8;
        unsigned ck = base + (m - 1 + ln) * 8;  // row m + lane of column k (this lane's pivot candidate)
#pragma unroll 1
        for (int k = 1; k <= n - 1; k++, ck += ld8) {
            jz = jz + 1;
            if (jz <= n && ln < ml) sts_d(base + (jz - 1) * ld8 + ln * 8, BV_R(0.0));  // LP:318-323
            __syncwarp();
            const int lm = (ml < n - k) ? ml : n - k;
            // l = idamax(lm+1, abd(m,k)) + m - 1: lanes 0..lm look at rows m..m+lm of column k
            const bool cact = (ln <= lm);
            real v = BV_R(0.0);
            if (cact) v = lds_d(ck);
            const int lofs = __ffs((int)Base::max_abs_ballot(v, cact)) - 1;  // pivot row - diagonal row
            const int piv = lofs + k;                                        // 1-based row of the matrix
            if (ln == k - 1) l.ipv = piv - 1;
            const real alk = shfl_d(v, lofs);
            if (alk == BV_R(0.0)) {
                info = k;
            } else {
                const bool swap = (lofs != 0);  // warp-uniform
                if (swap) {  // interchange, LP:335-339, in registers
                    const real akk = shfl_d(v, 0);
                    if (ln == lofs) v = akk;
                    if (ln == 0) v = alk;
                }
                const real t = -qrcp(alk);
                if (ln >= 1 && cact) v = t * v;  // dscal, LP:343
                if (cact && (swap || ln >= 1)) sts_d(ck, v);
                const int jun = mu + piv;
                ju = (ju > jun) ? ju : jun;
                ju = (ju < n) ? ju : n;
                __syncwarp();
                // row elimination with column indexing, LP:347-361: column j = k + 1 + lane;
                // mm = m - (lane+1) is the row of abd that holds matrix row k in column j
                if (ln < ju - k) {
                    const unsigned amm = base + (k + ln) * ld8 + (m - ln - 2) * 8;
                    const real tj = lds_d(amm + lofs * 8);
                    if (swap) {
                        const real tmm = lds_d(amm);
                        sts_d(amm + lofs * 8, tmm);
                        sts_d(amm, tj);
                    }
                    // daxpy, LP:359: abd(mm+i, j) += tj * abd(m+i, k), i = 1..lm
                    unsigned aj = amm + 8, ak = base + (k - 1) * ld8 + m * 8;
                    int cnt = lm;
#pragma unroll 1
                    for (; cnt >= 4; cnt -= 4, aj += 32, ak += 32) {
                        const real p0 = lds_d<0>(ak), p1 = lds_d<8>(ak), p2 = lds_d<16>(ak), p3 = lds_d<24>(ak);
                        const real r0 = lds_d<0>(aj), r1 = lds_d<8>(aj), r2 = lds_d<16>(aj), r3 = lds_d<24>(aj);
                        sts_d<0>(aj, r0 + tj * p0);
                        sts_d<8>(aj, r1 + tj * p1);
                        sts_d<16>(aj, r2 + tj * p2);
                        sts_d<24>(aj, r3 + tj * p3);
                    }
#pragma unroll 1
                    for (; cnt > 0; cnt--, aj += 8, ak += 8) sts_d(aj, lds_d(aj) + tj * lds_d(ak));
                }
                __syncwarp();
            }
        }
        if (ln == n - 1) l.ipv = n - 1;
        if (ABD(m, n) == BV_R(0.0)) info = n;
        __syncwarp();
#if !BVODE_STRICT
        // fast build: the diagonal keeps 1/u_kk so that dgbsl multiplies, and the entries of U
        // above it are scaled by -1/u_kk: back substitution is then b_i += b_k * (-u_ik/u_kk)
        // with the un-normalised b_k (one shuffle and one fma per step on the dependent chain)
        if (ln < n) {
            const int j = ln + 1;
            const real rp = qrcp(ABD(m, j));
            ABD(m, j) = rp;
            const int lmj = ((j < m) ? j : m) - 1;
            for (int i = 1; i <= lmj; i++) ABD(m - i, j) = -rp * ABD(m - i, j);
        }
        __syncwarp();
#endif
        return info;
    }

    // ---- dgbsl job = 0, LP:494-519: b_i in lane i-1 ------------------------------------------
    // Entry (row i = lane+1, column k) of the band array is abd[(k-1)*ld + (m - k + lane)]: one
    // shared address per lane that advances by (ld-1) doubles per column, for the multipliers
    // below the diagonal (forward elimination), the diagonal and U above it (back substitution).
    template <class ENG>
    BV_DEV static int solve(ENG &e, real (&x)[1]) {
        const Lin &l = e.lin;
        if constexpr (kInverse) {
            const unsigned pa = sm_addr(l.inv);
            Base::template solve_inverse_rows<LDI>(pa, pa + kInvVecOff * 8, l.ipv, x);
            return 0;
        }
        const int ln = lane();
        const int n = N, ml = l.ml, mu = l.mu, m = mu + ml + 1;
        const bool live = (N == 32) || (ln < n);
        const int step = (l.ld - 1) * 8;
        const unsigned a1 = sm_addr(l.abd) + (m - 1 + (live ? ln : 0)) * 8;  // column 1
        real b = x[0];
        if (ml != 0) {
            // any row interchange at all?  (warp-uniform; none for diagonally dominant P)
            const bool swaps = __any_sync(kFull, live && (l.ipv != ln)) != 0;
            unsigned a = a1;
#pragma unroll 4
            for (int k = 1; k <= n - 1; k++, a += step) {
                real t;
                if (swaps) {
                    const int lp = __shfl_sync(kFull, l.ipv, k - 1);  // 0-based
                    t = shfl_d(b, lp);
                    if (lp != k - 1) {
                        const real bk = shfl_d(b, k - 1);
                        if (ln == lp) b = bk;
                        if (ln == k - 1) b = t;
                    }
                } else {
                    t = shfl_d(b, k - 1);
                }
                // rows k+1 .. min(k+ml, n)
                const bool part = live && ((unsigned)(ln - k) < (unsigned)ml);
                const real mv = lds_d(a);
#if BVODE_STRICT
                if (part) b = b + t * mv;
#else
                (void)part;
                fma_if_ltu(b, t, mv, ln - k, live ? ml : 0);
#endif
            }
        }
        {
            unsigned a = a1 + (n - 1) * step;
#if BVODE_STRICT
#pragma unroll 4
            for (int k = n; k >= 1; k--, a -= step) {
                const real mv = lds_d(a);
                if (ln == k - 1) b = b / mv;
                const real t = -shfl_d(b, k - 1);
                // rows max(k-m+1, 1) .. k-1
                const bool part = live && ((unsigned)(k - 2 - ln) < (unsigned)(m - 1));
                if (part) b = b + t * mv;
            }
#else
#pragma unroll 4
            for (int k = n; k >= 2; k--, a -= step) {
                const real mv = lds_d(a);
                const real t = shfl_d(b, k - 1);
                fma_if_ltu(b, t, mv, k - 2 - ln, live ? m - 1 : 0);
            }
            if (live) b = b * lds_d(a1 + ln * step);  // x_k = b_k / u_kk
#endif
        }
        x[0] = b;
        return 0;
    }

    // ---- dvjac, miter 4 / 5 (M:3031-3104) --------------------------------------------------
    template <class ENG>
    BV_DEV static void jac_setup(ENG &e, int &ierpj) {
        Ctl &c = e.c;
        Lin &l = e.lin;
        const int ln = lane();
        const int n = N, ml = l.ml, mu = l.mu;
        const int mband = ml + mu + 1, meband = mband + ml;
        ierpj = 0;
        const real hrl1 = c.h * c.rl1;
        int jok = SAVEJ ? 1 : -1;
        if (SAVEJ) {
            if (e.hot.si(SI_NST) == 0 || e.hot.si(SI_NST) > e.hot.si(SI_NSLJ) + 50) jok = -1;
            if (c.icf == 1 && c.drc < BV_R(0.2)) jok = -1;
            if (c.icf == 2) jok = -1;
        }
        __syncwarp();
        if (jok == -1) {
            e.hot.si_add(SI_NJE, 1);
            e.hot.si(SI_NSLJ) = e.hot.si(SI_NST);
            c.jcur = 1;
            if constexpr (MITER == 4) {
                // zero wm(3:), then jac fills rows ml+1..meband of each column (M:3044-3047)
                if (ln < n)
                    for (int i = 1; i <= meband; i++) ABD(i, ln + 1) = BV_R(0.0);
                __syncwarp();
                // pd(i - j + mu + 1, j) = df_i/dy_j with pd = abd rows ml+1.. (nrowpd = meband)
                if constexpr (HasJacBandFull<PROB>::value)  // the reference's jac: pd = &abd(ml+1, 1), nrowpd = ld (M:3047)
                    PROB::jac_band_full(c.tn, e.y[0], &ABD(ml + 1, 1), l.ld, ml, mu, e.par, ln);
                else if (ln < n)
                    PROB::jac_band_col(c.tn, e.y[0], &ABD(ml + 1, ln + 1), ml, mu, e.par, ln);
            } else {
                // banded finite differences, M:3056-3081
                const int mba = (mband < e.nact) ? mband : e.nact;
                const real r0 = Base::fd_r0(c, e.savf, e.ewt, e.nact);
                const real srur = sqrt(kUround);
                const real ysave = e.y[0];
                // yh(:,1) is what the reference restores y from (M:3070); identical to ysave
                const real rmine = dmax2(srur * fabs(ysave), qdiv(r0, e.ewt[0]));
#pragma unroll 1
                for (int j = 1; j <= mba; j++) {
                    const bool mine = (ln < e.nact) && (((ln + 1 - j) % mband) == 0) && (ln + 1 >= j);
                    if (mine) e.y[0] = ysave + rmine;
                    e.feval(c.tn, e.y, e.acor);
                    e.y[0] = ysave;
                    // column jj = j, j+mband, ...: rows i1..i2, stored at abd(i - jj + mband, jj)
                    const real diff = e.acor[0] - e.savf[0];  // row ln+1
                    const int i = ln + 1;
                    // the perturbed column that covers row i in this group (bands do not overlap)
#pragma unroll 1
                    for (int jj = j; jj <= n; jj += mband) {
                        const real r = shfl_d(rmine, jj - 1);
                        const real fac = qrcp(r);
                        const int i1 = (jj - mu > 1) ? jj - mu : 1;
                        const int i2 = (jj + ml < n) ? jj + ml : n;
                        if (i >= i1 && i <= i2) ABD(i - jj + mband, jj) = diff * fac;
                    }
                }
                e.hot.si_add(SI_NFE, mba);
            }
            __syncwarp();
            if (e.nact != N) {  // neq was reduced: nothing of the dropped rows / columns (abd(i, j) holds a(i + j - mband, j), LP:230-260)
                if (ln < n)
                    for (int i = ml + 1; i <= meband; i++)
                        if (ln + 1 > e.nact || i + (ln + 1) - mband > e.nact) ABD(i, ln + 1) = BV_R(0.0);
                __syncwarp();
            }
            if (SAVEJ) {  // dacopy(mband, n, wm(ml3), meband, wm(locjs), mband)
                if (ln < n)
                    for (int i = 1; i <= mband; i++) l.wjs[ln * mband + (i - 1)] = ABD(ml + i, ln + 1);
            }
        } else {
            c.jcur = 0;
            if (SAVEJ) {
                if (ln < n)
                    for (int i = 1; i <= mband; i++) ABD(ml + i, ln + 1) = l.wjs[ln * mband + (i - 1)];
            }
        }
        __syncwarp();
        // multiply by -h*rl1 (all meband rows, M:3093) and add the identity (M:3094-3098)
        const real con = -hrl1;
        if (ln < n) {
            for (int i = 1; i <= meband; i++) ABD(i, ln + 1) = con * ABD(i, ln + 1);
            ABD(mband, ln + 1) = ABD(mband, ln + 1) + BV_R(1.0);
        }
        __syncwarp();
        e.hot.si_add(SI_NLU, 1);
        const int ier = factor(l);
        if (ier != 0) ierpj = 1;
    }
    // a system whose state was loaded for a continuation call: the band array holds P; rebuild its inverse
    template <class ENG>
    BV_DEV static void after_load(ENG &e) {
        if constexpr (kInverse) (void)invert(e.lin);
    }
#undef ABD
};

}  // namespace BVODE_NS
