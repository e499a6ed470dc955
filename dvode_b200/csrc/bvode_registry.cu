// bvode_registry.cu -- instantiates every kernel of one build and exports its problem table.
// Compiled twice: -DBVODE_STRICT=0 (fast, the product) and -DBVODE_STRICT=1 -fmad=false.
#include "bvode_kernels.cuh"

namespace BVODE_NS {
static const ProblemEntry kEntries[] = {
    BVODE_THREAD_PROBLEM("robertson", Robertson),
    BVODE_THREAD_PROBLEM("vdp", VanDerPol),
    BVODE_WARP_BAND_PROBLEM("advection", Advection25),
    BVODE_WARP_DENSE_FD_PROBLEM("mech32", Mech32),
};
}  // namespace BVODE_NS

#if BVODE_STRICT
extern "C" const ProblemEntry *bvode_registry_strict(int *n)
#else
extern "C" const ProblemEntry *bvode_registry_fast(int *n)
#endif
{
    *n = (int)(sizeof(BVODE_NS::kEntries) / sizeof(BVODE_NS::kEntries[0]));
    return BVODE_NS::kEntries;
}
