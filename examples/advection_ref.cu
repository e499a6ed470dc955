// examples/advection_ref.cu -- the reference's OWN callbacks, argument for argument, on the warp kernels.
//
// dvdemo problem 2 (/root/reference/test/dvdemo.f90:230-274: f2, jac2): ydot = A*y on a 5 x 5 grid, banded with
// ml = 5, mu = 0.  f and jac below are f2 / jac2 with `me` replaced by the per-system parameter block
// (alph1, alph2) -- no lane-wise rewrite; csrc/bvode_ref_problem.cuh stages y in shared memory and lets jac
// write straight into the kernel's LINPACK band array (pd(nrowpd, neq), M:3047).  The built-in problem
// "advection" is the lane-wise form of the same equations; tests/test_ref_adapter.py checks that both give the
// same bits as the oracle (strict build).
#include "bvode_user_problem.cuh"

namespace BVODE_NS {
struct AdvectionRef {
    static constexpr int N = 25, NPAR = 2, NG = 5;
    BV_DEV static void f(int neq, double, const double *y, double *ydot, const double *par) {
        (void)neq;
        for (int j = 1; j <= NG; j++) {
            for (int i = 1; i <= NG; i++) {
                const int k = i + (j - 1) * NG;
                double d = -2.0 * y[k - 1];
                if (i != 1) d = d + y[k - 2] * par[0];
                if (j != 1) d = d + y[k - NG - 1] * par[1];
                ydot[k - 1] = d;
            }
        }
    }
    // pd(i - j + mu + 1, j) = df_i/dy_j, column-major with leading dimension nrowpd, pre-zeroed (M:3044-3046)
    BV_DEV static void jac(int neq, double, const double *, int ml, int mu, double *pd, int nrowpd, const double *par) {
        const int mband = ml + mu + 1;
        const int mu1 = mu + 1, mu2 = mu + 2;
        for (int k = 1; k <= neq; k++) {
            pd[(mu1 - 1) + (k - 1) * nrowpd] = -2.0;
            pd[(mu2 - 1) + (k - 1) * nrowpd] = par[0];
            pd[(mband - 1) + (k - 1) * nrowpd] = par[1];
        }
        for (int k = NG; k <= neq; k += NG) pd[(mu2 - 1) + (k - 1) * nrowpd] = 0.0;
    }
};
}  // namespace BVODE_NS

BVODE_REGISTER_WARP_REF_BAND_PROBLEM("advection_ref", AdvectionRef)
