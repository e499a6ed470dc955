// examples/decay_chain.cu -- a user problem registered at run time.
//
// A -k1-> B,  2B -k2-> C + B',  C -k3-> D : four species, stiff for k2 >> k1, k3, mass conserving
// (y1 + y2 + y3 + y4 = const).  The three rate constants are the per-system parameters.
//     y1' = -k1 y1
//     y2' =  k1 y1 - k2 y2^2
//     y3' =  k2 y2^2 - k3 y3
//     y4' =  k3 y3
// f and jac have the argument meaning of the reference's interfaces (M:316-334): jac fills
// pd(i,j) = df_i/dy_j into an array the solver has zeroed (M:2947-2949).
#include "bvode_user_problem.cuh"

namespace BVODE_NS {
struct DecayChain {
    static constexpr int N = 4, NPAR = 3;
    BV_DEV static void f(double, const double (&y)[4], double (&ydot)[4], const double (&k)[3]) {
        const double r1 = k[0] * y[0], r2 = k[1] * y[1] * y[1], r3 = k[2] * y[2];
        ydot[0] = -r1;
        ydot[1] = r1 - r2;
        ydot[2] = r2 - r3;
        ydot[3] = r3;
    }
    BV_DEV static void jac(double, const double (&y)[4], double (&pd)[4][4], const double (&k)[3]) {
        pd[0][0] = -k[0];
        pd[1][0] = k[0];
        pd[1][1] = -2.0 * k[1] * y[1];
        pd[2][1] = 2.0 * k[1] * y[1];
        pd[2][2] = -k[2];
        pd[3][2] = k[2];
    }
};
}  // namespace BVODE_NS

BVODE_REGISTER_THREAD_PROBLEM("decay_chain", DecayChain)
