// problems.cuh -- the user callbacks f / jac (M:316-334) as __device__ functors.
//
// A problem is a struct with
//   static constexpr int N, NPAR;                       equations, per-system parameters
//   static void f  (t, y[N], ydot[N], par[NPAR]);       dy/dt = f(t, y)
//   static void jac(t, y[N], pd[N][N], par[NPAR]);      pd[i][j] = df_i/dy_j, pre-zeroed by the
//                                                       solver exactly as M:2947-2949 does
// compiled into the kernels by template instantiation (a host function pointer cannot be
// called from the device).  Registration with the C ABI is in bvode_registry.cu.
#pragma once
#include "bvode_math.cuh"

namespace BVODE_NS {

// Robertson kinetics, /root/reference/test/example.f90:79-109, with the three rate
// constants as per-system parameters (reference values 0.04, 1e4, 3e7).
struct Robertson {
    static constexpr int N = 3, NPAR = 3;
    BV_DEV static void f(real, const real (&y)[3], real (&ydot)[3], const real (&k)[3]) {
        ydot[0] = -k[0] * y[0] + k[1] * y[1] * y[2];
        ydot[2] = k[2] * y[1] * y[1];
        ydot[1] = -ydot[0] - ydot[2];
    }
    BV_DEV static void jac(real, const real (&y)[3], real (&pd)[3][3], const real (&k)[3]) {
        pd[0][0] = -k[0];
        pd[0][1] = k[1] * y[2];
        pd[0][2] = k[1] * y[1];
        pd[1][0] = k[0];
        pd[1][2] = -pd[0][2];
        pd[2][1] = (BV_R(2.0) * k[2]) * y[1];
        pd[1][1] = -pd[0][1] - pd[2][1];
    }
};

// Van der Pol oscillator, /root/reference/test/dvdemo.f90:207-228; par[0] = 3 there.
struct VanDerPol {
    static constexpr int N = 2, NPAR = 1;
    BV_DEV static void f(real, const real (&y)[2], real (&ydot)[2], const real (&p)[1]) {
        ydot[0] = y[1];
        ydot[1] = p[0] * (BV_R(1.0) - y[0] * y[0]) * y[1] - y[0];
    }
    BV_DEV static void jac(real, const real (&y)[2], real (&pd)[2][2], const real (&p)[1]) {
        pd[0][0] = BV_R(0.0);
        pd[0][1] = BV_R(1.0);
        pd[1][0] = -(BV_R(2.0) * p[0]) * y[0] * y[1] - BV_R(1.0);
        pd[1][1] = p[0] * (BV_R(1.0) - y[0] * y[0]);
    }
};

// Van der Pol with USER-SUPPLIED error weights and norm: the optional `dewset` / `dvnorm`
// procedure arguments of initialize (M:431-503) as two more members of the functor.
//   ewt (i, y_i, rtol_i, atol_i)  replaces dewset_default (M:3238-3273), one component
//   norm(v, w)                    replaces dvnorm_default (M:3295-3311)
// Here: ewt_i = rtol_i*max(|y_i|, 1) + atol_i and the max norm.
struct VanDerPolMaxNorm : VanDerPol {
    BV_DEV static real ewt(int, real yi, real rtol, real atol) {
        real ay = fabs(yi);
        if (ay < BV_R(1.0)) ay = BV_R(1.0);
        return rtol * ay + atol;
    }
    BV_DEV static real norm(const real (&v)[2], const real (&w)[2]) {
        real vm = BV_R(0.0);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const real p = fabs(v[i] * w[i]);
            if (p > vm) vm = p;
        }
        return vm;
    }
};

// Van der Pol with a user dewset (M:431-467) whose first weight turns non-positive once |y_1| < 1/2
// (builder-defined test problem, the oracle's problem 7): it drives the `ewt(i) <= 0` exit of the step
// loop, istate = -6 (M:1516-1521), which the default weights reach only for a component that is exactly 0.
struct VanDerPolEwtDrop : VanDerPol {
    BV_DEV static real ewt(int i, real yi, real rtol, real atol) {
        real e = rtol * fabs(yi) + atol;
        if (i == 0 && fabs(yi) < BV_R(0.5)) e = e - BV_R(1.0);
        return e;
    }
};

}  // namespace BVODE_NS

#if !BVODE_REAL_IS_FLOAT
// ----------------------------------------------------------------------------------------
// Warp-per-system problems: component i on lane i.  f_lane returns ydot of this lane's
// component; neighbours' values come from warp shuffles (all 32 lanes must call).
#include "../../include/bvode_mech32.h"

namespace BVODE_NS {

// dvdemo problem 2: ydot = A*y, 5x5-grid advection, neq = 25, banded with ml = 5, mu = 0
// (/root/reference/test/dvdemo.f90:230-274).  par = (alph1, alph2), reference values 1, 1.
struct Advection25 {
    static constexpr int N = 25, NPAR = 2, NPAR_LANE = 2, NG = 5;
    BV_DEV static void load_par(real (&par)[2], const real *params, size_t nsys, size_t sys, int) {
        par[0] = params[sys];
        par[1] = params[nsys + sys];
    }
    BV_DEV static real f_lane(real, real y, const real (&p)[2], int ln) {
        const real ym1 = __shfl_up_sync(0xffffffffu, y, 1);
        const real ymg = __shfl_up_sync(0xffffffffu, y, NG);
        const int i = ln % NG, j = ln / NG;
        real d = -BV_R(2.0) * y;
        if (i != 0) d = d + ym1 * p[0];
        if (j != 0) d = d + ymg * p[1];
        return d;
    }
    // column j = ln+1 of the band Jacobian; pd points at pd(1, j), pd(i-j+mu+1, j) = df_i/dy_j
    BV_DEV static void jac_band_col(real, real, real *pd, int ml, int mu, const real (&p)[2],
                                    int ln) {
        const int mband = ml + mu + 1;
        pd[mu] = -BV_R(2.0);
        pd[mu + 1] = p[0];
        pd[mband - 1] = p[1];
        if ((ln + 1) % NG == 0) pd[mu + 1] = BV_R(0.0);
    }
};

// dvdemo problem 2 with USER-SUPPLIED error weights and norm (M:431-503) in the lane-wise form of
// the warp kernels: `ewt` as for thread problems (one component), and the norm as a fold over the
// components -- norm_term(v_i, w_i), norm_combine(acc, term) applied in index order (the strict build
// folds sequentially like a Fortran loop; the fast build as a butterfly, so combine should be
// associative up to rounding), norm_init its identity, norm_final(acc).  Here: the weights and the
// max norm of VanDerPolMaxNorm.
struct Advection25MaxNorm : Advection25 {
    BV_DEV static real ewt(int, real yi, real rtol, real atol) {
        real ay = fabs(yi);
        if (ay < BV_R(1.0)) ay = BV_R(1.0);
        return rtol * ay + atol;
    }
    static constexpr real norm_init = BV_R(0.0);
    BV_DEV static real norm_term(real v, real w) { return fabs(v * w); }
    BV_DEV static real norm_combine(real acc, real p) { return (p > acc) ? p : acc; }
    BV_DEV static real norm_final(real acc) { return acc; }
};

// Synthetic 32-species mass-action network (BASELINE.json config 4; builder-defined, NOT in
// the reference).  64 reactions: r < 32 unimolecular A -> B, r >= 32 bimolecular
// A + B -> C + D; tables in include/bvode_mech32.h.  par = 64 rate constants; lane r holds
// k_r and k_{r+32} and evaluates those two reaction rates; species i then sums its entry
// list in reaction order (the same order as the oracle's mech_f).
__device__ const int g_mech_ra[BVODE_MECH_NR] = BVODE_MECH_RA_INIT;
__device__ const int g_mech_rb[BVODE_MECH_NR] = BVODE_MECH_RB_INIT;
__device__ const int g_mech_ent_r[BVODE_MECH_NS][BVODE_MECH_MAXE] = BVODE_MECH_ENT_R_INIT;
__device__ const real g_mech_ent_s[BVODE_MECH_NS][BVODE_MECH_MAXE] = BVODE_MECH_ENT_S_INIT;

struct Mech32 {
    static constexpr int N = 32, NPAR = 64, NPAR_LANE = 2;
    BV_DEV static void load_par(real (&par)[2], const real *params, size_t nsys, size_t sys,
                                int ln) {
        par[0] = params[(size_t)ln * nsys + sys];
        par[1] = params[(size_t)(ln + 32) * nsys + sys];
    }
    BV_DEV static real f_lane(real, real y, const real (&k)[2], int ln) {
        const unsigned full = 0xffffffffu;
        // this lane's two reactions
        const real rate0 = k[0] * __shfl_sync(full, y, g_mech_ra[ln]);  // unimolecular
        real rate1 = k[1] * __shfl_sync(full, y, g_mech_ra[ln + 32]);
        rate1 = rate1 * __shfl_sync(full, y, g_mech_rb[ln + 32]);       // bimolecular
        real d = BV_R(0.0);
#pragma unroll
        for (int e = 0; e < BVODE_MECH_MAXE; e++) {
            const int r = g_mech_ent_r[ln][e];
            const real s = g_mech_ent_s[ln][e];
            const int src = r & 31;
            const real v0 = __shfl_sync(full, rate0, src);
            const real v1 = __shfl_sync(full, rate1, src);
            const real q = (r >= 32) ? v1 : v0;
            d = d + s * q;
        }
        return d;
    }
    // Row `ln` of the Jacobian, accumulated entry by entry in the order of f (all 32 lanes call: the
    // rate constants and the partner concentrations come by shuffle).  d(k y_a y_b)/dy_a = k y_b,
    // d/dy_b = k y_a; unimolecular: k.  `row` is pre-zeroed by the solver (M:2947-2949).
    BV_DEV static void jac_row(real, real y, real *row, const real (&k)[2], int ln) {
        const unsigned full = 0xffffffffu;
#pragma unroll 1
        for (int e = 0; e < BVODE_MECH_MAXE; e++) {
            const int r = g_mech_ent_r[ln][e];  // -1: unused slot (still takes part in the shuffles)
            const real s = g_mech_ent_s[ln][e];
            const int rr = (r < 0) ? 0 : r;
            const int a = g_mech_ra[rr], b = g_mech_rb[rr];  // b = -1: unimolecular
            const real k0 = __shfl_sync(full, k[0], rr & 31), k1 = __shfl_sync(full, k[1], rr & 31);
            const real kr = (rr >= 32) ? k1 : k0;
            const real ya = __shfl_sync(full, y, a), yb = __shfl_sync(full, y, (b < 0) ? 0 : b);
            if (r >= 0) {
                if (b < 0) {
                    row[a] = row[a] + s * kr;
                } else {
                    row[a] = row[a] + s * (kr * yb);
                    row[b] = row[b] + s * (kr * ya);
                }
            }
        }
    }
};

// 1-D reaction-diffusion chain with n = 48 cells and walls at both ends (builder-defined, NOT in the
// reference: it exercises the two-components-per-lane mapping for 32 < n <= 64):
//     y_i' = D (y_{i-1} - 2 y_i + y_{i+1}) - k y_i^2 + s_i,   s_i = 1 in the first four cells.
// par = (D, k).  f_lane2: this lane's components ln (slot 0) and ln + 32 (slot 1).
struct Diffusion48 {
    static constexpr int N = 48, NPAR = 2, NPAR_LANE = 2;
    BV_DEV static void load_par(real (&par)[2], const real *params, size_t nsys, size_t sys, int) {
        par[0] = params[sys];
        par[1] = params[nsys + sys];
    }
    BV_DEV static void f_lane2(real, const real (&y)[2], real (&ydot)[2], const real (&p)[2], int ln) {
        const unsigned full = 0xffffffffu;
        const real up0 = __shfl_up_sync(full, y[0], 1), dn0 = __shfl_down_sync(full, y[0], 1);
        const real up1 = __shfl_up_sync(full, y[1], 1), dn1 = __shfl_down_sync(full, y[1], 1);
        const real y31 = __shfl_sync(full, y[0], 31), y32 = __shfl_sync(full, y[1], 0);
        const real l0 = (ln > 0) ? up0 : BV_R(0.0), r0 = (ln < 31) ? dn0 : y32;
        const real l1 = (ln > 0) ? up1 : y31, r1 = (ln + 33 < N) ? dn1 : BV_R(0.0);
        const real s0 = (ln < 4) ? BV_R(1.0) : BV_R(0.0);
        ydot[0] = p[0] * ((l0 - BV_R(2.0) * y[0]) + r0) - p[1] * y[0] * y[0] + s0;
        ydot[1] = p[0] * ((l1 - BV_R(2.0) * y[1]) + r1) - p[1] * y[1] * y[1] + BV_R(0.0);
    }
};

}  // namespace BVODE_NS
#endif  // !BVODE_REAL_IS_FLOAT
