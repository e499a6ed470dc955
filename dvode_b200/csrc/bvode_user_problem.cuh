// bvode_user_problem.cuh -- compile YOUR f / jac into a problem library and register it at run time.
//
// The reference takes f and jac as procedure arguments of initialize (M:377-430); a kernel cannot
// call a host function pointer, so here they are a __device__ functor (the forms of problems.cuh)
// instantiated into every kernel of the solver by ONE macro at the end of your .cu file:
//
//     #include "bvode_user_problem.cuh"
//     namespace BVODE_NS {
//     struct MyProblem {
//         static constexpr int N = 4, NPAR = 3;
//         BV_DEV static void f  (double t, const double (&y)[N], double (&ydot)[N], const double (&p)[NPAR]);
//         BV_DEV static void jac(double t, const double (&y)[N], double (&pd)[N][N], const double (&p)[NPAR]);
//     };  // pd[i][j] = df_i/dy_j, pre-zeroed by the solver (M:2947-2949); optional members ewt / norm
//     }
//     BVODE_REGISTER_THREAD_PROBLEM("my_problem", MyProblem)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -I dvode_b200/csrc \
//        -DBVODE_STRICT=0 -c my.cu -o my_fast.o          (and once more with -DBVODE_STRICT=1 -fmad=false)
//   nvcc -shared -o libmy.so my_fast.o my_strict.o
//
// then bvode_register_problem(bvode_user_entry_fast(), bvode_user_entry_strict(), bvode_user_abi_fast())
// (dvode_b200.plugins.build_problem / register_problem do both steps from Python).
#pragma once
#include "bvode_kernels.cuh"
#include "bvode_ref_problem.cuh"

#define BV_UCAT2(a, b) a##b
#define BV_UCAT(a, b) BV_UCAT2(a, b)
#if BVODE_STRICT
#define BVODE_BUILD_TAG strict
#else
#define BVODE_BUILD_TAG fast
#endif
#define BVODE_USER_ENTRY_(KIND, NAME, PROB)                                                          \
    namespace BVODE_NS {                                                                             \
    static const ProblemEntry kUserEntry = KIND(NAME, PROB);                                         \
    }                                                                                                \
    extern "C" __attribute__((visibility("default"))) const void *BV_UCAT(bvode_user_entry_,         \
                                                                           BVODE_BUILD_TAG)(void) { \
        return &BVODE_NS::kUserEntry;                                                                \
    }                                                                                                \
    extern "C" __attribute__((visibility("default"))) int BV_UCAT(bvode_user_abi_,                   \
                                                                  BVODE_BUILD_TAG)(void) {           \
        return BVODE_PLUGIN_ABI;                                                                     \
    }
// one system per thread (tiny n): meth 1 / 2, miter 0 - 3
#define BVODE_REGISTER_THREAD_PROBLEM(NAME, PROB) BVODE_USER_ENTRY_(BVODE_THREAD_PROBLEM, NAME, PROB)
// one system per warp (n <= 32), lane-wise f_lane: banded (miter 4 / 5) or dense FD (miter 2) Jacobian
#define BVODE_REGISTER_WARP_BAND_PROBLEM(NAME, PROB) BVODE_USER_ENTRY_(BVODE_WARP_BAND_PROBLEM, NAME, PROB)
#define BVODE_REGISTER_WARP_DENSE_FD_PROBLEM(NAME, PROB) BVODE_USER_ENTRY_(BVODE_WARP_DENSE_FD_PROBLEM, NAME, PROB)
// one system per warp, two components per lane (32 < n <= 64), f_lane2: dense FD Jacobian (miter 2)
#define BVODE_REGISTER_WARP2_DENSE_FD_PROBLEM(NAME, PROB) BVODE_USER_ENTRY_(BVODE_WARP2_DENSE_FD_PROBLEM, NAME, PROB)

// the same three mappings for callbacks in the REFERENCE's whole-vector form f(neq, t, y, ydot, par) /
// jac(neq, t, y, ml, mu, pd, nrowpd, par) (M:316-334; csrc/bvode_ref_problem.cuh): no lane-wise rewrite
#define BVODE_REGISTER_WARP_REF_BAND_PROBLEM(NAME, REF) BVODE_USER_ENTRY_(BVODE_WARP_BAND_PROBLEM, NAME, RefWarp<REF>)
#define BVODE_REGISTER_WARP_REF_DENSE_PROBLEM(NAME, REF) BVODE_USER_ENTRY_(BVODE_WARP_DENSE_FD_PROBLEM, NAME, RefWarp<REF>)
#define BVODE_REGISTER_WARP_REF2_DENSE_PROBLEM(NAME, REF) BVODE_USER_ENTRY_(BVODE_WARP2_DENSE_FD_PROBLEM, NAME, RefWarp2<REF>)
