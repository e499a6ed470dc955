// examples/vdp_ref.cu -- dvdemo problem 1 (Van der Pol, /root/reference/test/dvdemo.f90:207-228: f1, jac1) with the
// reference's own argument lists on the dense warp kernels (csrc/bvode_ref_problem.cuh): n = 2 on a warp is
// not how one would map this problem (the built-in "vdp" runs one system per thread) -- it is here to show
// that ANY f / jac pair written for the reference registers as it stands, and to check the dense adapter
// (jac fills a column-major pd, the warp transposes) bit for bit against the oracle.
#include "bvode_user_problem.cuh"

namespace BVODE_NS {
struct VdpRef {
    static constexpr int N = 2, NPAR = 1;
    BV_DEV static void f(int, double, const double *y, double *ydot, const double *par) {
        ydot[0] = y[1];
        ydot[1] = par[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
    }
    BV_DEV static void jac(int, double, const double *y, int, int, double *pd, int nrowpd, const double *par) {
        pd[0 + 0 * nrowpd] = 0.0;
        pd[0 + 1 * nrowpd] = 1.0;
        pd[1 + 0 * nrowpd] = -(2.0 * par[0]) * y[0] * y[1] - 1.0;
        pd[1 + 1 * nrowpd] = par[0] * (1.0 - y[0] * y[0]);
    }
};
}  // namespace BVODE_NS

BVODE_REGISTER_WARP_REF_DENSE_PROBLEM("vdp_ref", VdpRef)
