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
    BV_DEV static void f(double, const double (&y)[3], double (&ydot)[3], const double (&k)[3]) {
        ydot[0] = -k[0] * y[0] + k[1] * y[1] * y[2];
        ydot[2] = k[2] * y[1] * y[1];
        ydot[1] = -ydot[0] - ydot[2];
    }
    BV_DEV static void jac(double, const double (&y)[3], double (&pd)[3][3], const double (&k)[3]) {
        pd[0][0] = -k[0];
        pd[0][1] = k[1] * y[2];
        pd[0][2] = k[1] * y[1];
        pd[1][0] = k[0];
        pd[1][2] = -pd[0][2];
        pd[2][1] = (2.0 * k[2]) * y[1];
        pd[1][1] = -pd[0][1] - pd[2][1];
    }
};

// Van der Pol oscillator, /root/reference/test/dvdemo.f90:207-228; par[0] = 3 there.
struct VanDerPol {
    static constexpr int N = 2, NPAR = 1;
    BV_DEV static void f(double, const double (&y)[2], double (&ydot)[2], const double (&p)[1]) {
        ydot[0] = y[1];
        ydot[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
    }
    BV_DEV static void jac(double, const double (&y)[2], double (&pd)[2][2], const double (&p)[1]) {
        pd[0][0] = 0.0;
        pd[0][1] = 1.0;
        pd[1][0] = -(2.0 * p[0]) * y[0] * y[1] - 1.0;
        pd[1][1] = p[0] * (1.0 - y[0] * y[0]);
    }
};

}  // namespace BVODE_NS
