// bvode_ref_problem.cuh -- user callbacks in the REFERENCE's own form for one-system-per-warp kernels.
//
// The reference takes (M:316-334)
//     subroutine f  (me, neq, t, y, ydot)                     y(neq) -> ydot(neq)
//     subroutine jac(me, neq, t, y, ml, mu, pd, nrowpd)       pd(nrowpd, neq), pre-zeroed by the solver
// The warp kernels keep component i of every vector on lane i, and their fast path wants the problem
// written lane-wise (f_lane / jac_row / jac_band_col in problems.cuh: every lane computes its own
// component, neighbours' values by shuffle).  This adapter takes the whole-vector form instead, so that
// any problem with n <= 64 registers without a lane-wise rewrite:
//
//     #include "bvode_user_problem.cuh"
//     namespace BVODE_NS {
//     struct MyProblem {                                            // exactly the reference's arguments
//         static constexpr int N = 25, NPAR = 2;                    // (`me` is the per-system parameter block)
//         BV_DEV static void f  (int neq, double t, const double *y, double *ydot, const double *par);
//         BV_DEV static void jac(int neq, double t, const double *y, int ml, int mu, double *pd, int nrowpd,
//                                const double *par);                // optional (analytic Jacobians, miter 1 / 4)
//     };
//     }
//     BVODE_REGISTER_WARP_REF_BAND_PROBLEM("my_problem", MyProblem)   // or ..._REF_DENSE_..., ..._REF2_DENSE_...
//
// How: y is staged in a per-warp scratch area of shared memory (every lane stores its component), lane 0
// runs the user's f on the staged vector and writes ydot next to it, every lane reads its component back.
// jac writes straight into the kernel's matrix: the band array IS LINPACK's pd(nrowpd, neq) (nrowpd = the
// padded leading dimension, pd offset to row ml + 1 as in M:3047); the dense matrix is stored row-major,
// so a column-major pd is its transpose -- the callback fills it and the warp transposes in place.
// The callbacks run on one lane: this form trades throughput of f for convenience (the integrator around
// it -- predictor, norms, LU, solves -- still runs on all lanes).  Results are those of the lane-wise form
// operation for operation (tests/test_ref_adapter.py: bit for bit against the oracle, strict build).
#pragma once
#include "bvode_math.cuh"

namespace BVODE_NS {

template <class R, class = void>
struct RefHasJac { static constexpr bool value = false; };
template <class R>
struct RefHasJac<R, decltype((void)&R::jac)> { static constexpr bool value = true; };

template <class REF>
struct RefWarpCommon {
    static constexpr int N = REF::N, NPAR = REF::NPAR, NPAR_LANE = 1;
    static constexpr int NP = (N + 1) & ~1;                                  // vectors padded to 16 bytes
    static constexpr int kScratchDoubles = 2 * NP + ((NPAR + 1) & ~1) + 2;   // y, ydot, par per warp
    // this warp's scratch: at the END of the CTA's dynamic shared memory (the kernels lay out their own data
    // from the start; launch_warp adds kScratchDoubles per warp for problems that declare it)
    BV_DEV static real *scratch() {
        extern __shared__ real bv_dyn_smem[];
        unsigned total;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(total));
        const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
        return bv_dyn_smem + (total >> 3) - (nw - w) * kScratchDoubles;
    }
    BV_DEV static real *par_of(real *sc) { return sc + 2 * NP; }
    BV_DEV static void load_par(real (&)[1], const real *params, size_t nsys, size_t sys, int ln) {
        real *ps = par_of(scratch());
        for (int k = ln; k < NPAR; k += 32) ps[k] = params[(size_t)k * nsys + sys];
        __syncwarp();
    }
};

// ---- n <= 32: one component per lane ------------------------------------------------------------
template <class REF, bool HASJAC = RefHasJac<REF>::value>
struct RefWarp : RefWarpCommon<REF> {
    using C = RefWarpCommon<REF>;
    static constexpr int N = REF::N;
    static_assert(N <= 32, "one component per lane: n <= 32 (RefWarp2 for 32 < n <= 64)");
    BV_DEV static void stage(real y, int ln) {
        real *sc = C::scratch();
        __syncwarp();
        if (ln < N) sc[ln] = y;
        __syncwarp();
    }
    BV_DEV static real f_lane(real t, real y, const real (&)[1], int ln) {
        real *sc = C::scratch();
        stage(y, ln);
        if (ln == 0) REF::f(N, t, sc, sc + C::NP, C::par_of(sc));
        __syncwarp();
        return (ln < N) ? sc[C::NP + ln] : BV_R(0.0);
    }
};
template <class REF>
struct RefWarp<REF, true> : RefWarp<REF, false> {
    using B = RefWarp<REF, false>;
    using C = RefWarpCommon<REF>;
    static constexpr int N = REF::N;
    // banded analytic Jacobian (miter 4): pd = &abd(ml + 1, 1), nrowpd = the array's leading dimension (M:3047)
    BV_DEV static void jac_band_full(real t, real y, real *pd, int nrowpd, int ml, int mu, const real (&)[1], int ln) {
        real *sc = C::scratch();
        B::stage(y, ln);
        if (ln == 0) REF::jac(N, t, sc, ml, mu, pd, nrowpd, C::par_of(sc));
        __syncwarp();
    }
    // dense analytic Jacobian (miter 1): the kernel's matrix is p[i*ldp + j]; the callback sees it as the
    // column-major pd(nrowpd = ldp, neq), i.e. transposed, and the warp transposes it in place afterwards
    BV_DEV static void jac_dense_full(real t, real y, real *p, int ldp, const real (&)[1], int ln) {
        real *sc = C::scratch();
        B::stage(y, ln);
        if (ln == 0) REF::jac(N, t, sc, 0, 0, p, ldp, C::par_of(sc));
        __syncwarp();
        for (int i = 0; i < N; i++) {  // lane j > i swaps (i, j) with (j, i)
            if (ln > i && ln < N) {
                const real a = p[i * ldp + ln], b = p[ln * ldp + i];
                p[i * ldp + ln] = b;
                p[ln * ldp + i] = a;
            }
        }
        __syncwarp();
    }
};

// ---- 32 < n <= 64: two components per lane (component c on lane c & 31, slot c >> 5) -----------------
template <class REF>
struct RefWarp2 : RefWarpCommon<REF> {
    using C = RefWarpCommon<REF>;
    static constexpr int N = REF::N;
    static_assert(N > 32 && N <= 64, "two components per lane: 32 < n <= 64");
    BV_DEV static void f_lane2(real t, const real (&y)[2], real (&ydot)[2], const real (&)[1], int ln) {
        real *sc = C::scratch();
        __syncwarp();
        sc[ln] = y[0];
        if (ln + 32 < N) sc[ln + 32] = y[1];
        __syncwarp();
        if (ln == 0) REF::f(N, t, sc, sc + C::NP, C::par_of(sc));
        __syncwarp();
        ydot[0] = sc[C::NP + ln];
        ydot[1] = (ln + 32 < N) ? sc[C::NP + ln + 32] : BV_R(0.0);
    }
};

}  // namespace BVODE_NS
