// bvode_registry.cu -- instantiates every kernel of ONE problem for one build and exports its
// table entry.  Compiled once per (build, problem): -DBVODE_STRICT=0 (fast, the product) or
// -DBVODE_STRICT=1 -fmad=false, and -DBVODE_UNIT=0..7.
#include "bvode_kernels.cuh"

#if BVODE_REAL_IS_FLOAT && (BVODE_UNIT == 2 || BVODE_UNIT == 3 || BVODE_UNIT == 5 || BVODE_UNIT == 6)
#error "the REAL32 build carries the thread-per-system problems only (units 0, 1, 4, 7)"
#endif
#ifndef BVODE_UNIT
#error "compile with -DBVODE_UNIT=<problem index>"
#endif

namespace BVODE_NS {
#if BVODE_UNIT == 0
static const ProblemEntry kEntry = BVODE_THREAD_PROBLEM("robertson", Robertson);
#elif BVODE_UNIT == 1
static const ProblemEntry kEntry = BVODE_THREAD_PROBLEM("vdp", VanDerPol);
#elif BVODE_UNIT == 2
static const ProblemEntry kEntry = BVODE_WARP_BAND_PROBLEM("advection", Advection25);
#elif BVODE_UNIT == 3
static const ProblemEntry kEntry = BVODE_WARP_DENSE_FD_PROBLEM("mech32", Mech32);
#elif BVODE_UNIT == 4
static const ProblemEntry kEntry = BVODE_THREAD_PROBLEM("vdp_maxnorm", VanDerPolMaxNorm);
#elif BVODE_UNIT == 5
static const ProblemEntry kEntry = BVODE_WARP_BAND_PROBLEM("advection_maxnorm", Advection25MaxNorm);
#elif BVODE_UNIT == 6
static const ProblemEntry kEntry = BVODE_WARP2_DENSE_FD_PROBLEM("diffusion48", Diffusion48);
#elif BVODE_UNIT == 7
static const ProblemEntry kEntry = BVODE_THREAD_PROBLEM("vdp_ewtdrop", VanDerPolEwtDrop);
#endif
}  // namespace BVODE_NS

#define BV_CAT2(a, b) a##b
#define BV_CAT(a, b) BV_CAT2(a, b)
#if BVODE_STRICT
extern "C" const ProblemEntry *BV_CAT(bvode_entry_strict_, BVODE_UNIT)(void)
#else
extern "C" const ProblemEntry *BV_CAT(bvode_entry_fast_, BVODE_UNIT)(void)
#endif
{
    return &BVODE_NS::kEntry;
}
