// bvode_real.h -- the working precision of the library (the reference's `wp`, src/dvode_kinds_module.F90:29-31).
// REAL64 (default) or REAL32 (-DBVODE_REAL32: libbvode32.so, strict arithmetic, thread-per-system kernels).
// BV_R(x) types a floating literal in the working precision, like the reference's `1.0e-6_wp`.
#pragma once
#ifdef BVODE_REAL32
typedef float real;
#define BVODE_REAL_IS_FLOAT 1
#else
typedef double real;
#define BVODE_REAL_IS_FLOAT 0
#endif
#define BV_R(x) ((real)(x))
