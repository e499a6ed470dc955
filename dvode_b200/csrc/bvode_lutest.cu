// bvode_lutest.cu -- the device LINPACK routines in isolation (known-answer tests).
//
// The reference never tests dgefa / dgesl / dgbfa / dgbsl on their own (LP:46-522); the CUDA forms here
// are different enough in shape (implicit pivoting, rows across lanes, rank-2 updates) to deserve it.
// One kernel per mapping factors a batch of matrices and solves one right-hand side each with exactly
// the policy code the integrator kernels run: singular matrices (info = k, LP:87-88), exact pivot ties
// (BL:334: first maximum), band matrices with and without interchanges.  Compiled twice like the
// integrator (fast / strict); reached through bvode_lu_test (include/bvode.h).
#include "bvode_kernels.cuh"

namespace BVODE_NS {

template <int N_>
struct LuProb {  // a "problem" with no f: only N is used by the linear algebra
    static constexpr int N = N_, NPAR = 0, NPAR_LANE = 0;
};
template <class LIN>
struct LuEng {  // the part of Engine<POL> that solve() touches
    LIN lin;
};

// kind 0: one matrix per thread, registers (ThreadPolicy).  a: column-major [n*n]; afac: LINPACK layout
template <int N>
__global__ void k_lu_thread(int nmat, const real *a, const real *b, real *afac, int *ipvt, real *x, int *info) {
    using POL = ThreadPolicy<LuProb<N>, 1, false, 2>;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmat) return;
    real A[N][N], bv[N];
    int ip[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) A[i][j] = a[(size_t)m * N * N + j * N + i];
        bv[i] = b[(size_t)m * N + i];
    }
    const int inf = POL::dgefa(A, ip);
    if (inf == 0) POL::dgesl(A, ip, bv);
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) afac[(size_t)m * N * N + j * N + i] = A[i][j];
        ipvt[(size_t)m * N + i] = ip[i] + 1;
        x[(size_t)m * N + i] = (inf == 0) ? bv[i] : BV_R(0.0);
    }
    info[m] = inf;
}

// kind 1: one matrix per warp, row i on lane i (WarpDensePolicy).  afac: row r of the shared-memory image
// at [r*n + j]; ipvt: perm[k] = the ROW that is the pivot of step k (0-based; rows never move: implicit pivoting)
template <int N>
__global__ void k_lu_dense(int nmat, const real *a, const real *b, real *afac, int *ipvt, real *x, int *info) {
    using POL = WarpDensePolicy<LuProb<N>, 2, false, 2>;
    extern __shared__ real smem[];
    const int m = blockIdx.x, ln = threadIdx.x & 31;
    if (m >= nmat) return;
    LuEng<typename POL::Lin> e;
    e.lin.p = smem;
    e.lin.perm = reinterpret_cast<int *>(smem + POL::kPermOff);
    if (ln < N)
        for (int j = 0; j < N; j++) e.lin.p[ln * POL::LDP + j] = a[(size_t)m * N * N + j * N + ln];
    POL::lin_init(e);
    const int inf = POL::factor(e.lin);
    real xv[1] = {(ln < N) ? b[(size_t)m * N + ln] : BV_R(0.0)};
    if (inf == 0) POL::solve(e, xv);
    __syncwarp();
    if (ln < N) {
        for (int j = 0; j < N; j++) afac[(size_t)m * N * N + ln * N + j] = e.lin.p[ln * POL::LDP + j];
        ipvt[(size_t)m * N + ln] = e.lin.perm[ln];
        x[(size_t)m * N + ln] = (inf == 0) ? xv[0] : BV_R(0.0);
    }
    if (ln == 0) info[m] = inf;
}

// kind 2: one matrix per warp, rows ln and ln + 32 on lane ln (WarpDense2Policy)
template <int N>
__global__ void k_lu_dense2(int nmat, const real *a, const real *b, real *afac, int *ipvt, real *x, int *info) {
    using POL = WarpDense2Policy<LuProb<N>, 2, false, 2>;
    extern __shared__ real smem[];
    const int m = blockIdx.x, ln = threadIdx.x & 31;
    if (m >= nmat) return;
    LuEng<typename POL::Lin> e;
    e.lin.p = smem;
    e.lin.perm = reinterpret_cast<int *>(smem + POL::kPermOff);
    for (int s = 0; s < 2; s++) {
        const int r = ln + 32 * s;
        if (r < N)
            for (int j = 0; j < N; j++) e.lin.p[r * POL::LDP + j] = a[(size_t)m * N * N + j * N + r];
    }
    POL::lin_init(e);
    const int inf = POL::factor(e.lin);
    real xv[2] = {b[(size_t)m * N + ln], (ln + 32 < N) ? b[(size_t)m * N + ln + 32] : BV_R(0.0)};
    if (inf == 0) POL::solve(e, xv);
    __syncwarp();
    for (int s = 0; s < 2; s++) {
        const int r = ln + 32 * s;
        if (r < N) {
            for (int j = 0; j < N; j++) afac[(size_t)m * N * N + r * N + j] = e.lin.p[r * POL::LDP + j];
            ipvt[(size_t)m * N + r] = e.lin.perm[r];
            x[(size_t)m * N + r] = (inf == 0) ? xv[s] : BV_R(0.0);
        }
    }
    if (ln == 0) info[m] = inf;
}

// kind 3: LINPACK band storage abd(lda, n), lda = 2*ml + mu + 1, one matrix per warp (WarpBandPolicy).
// afac / ipvt: LINPACK's own layout (ipvt 1-based).
template <int N>
__global__ void k_lu_band(int nmat, int ml, int mu, const real *a, const real *b, real *afac, int *ipvt, real *x,
                          int *info) {
    using POL = WarpBandPolicy<LuProb<N>, 4, false, 2>;
    extern __shared__ real smem[];
    const int m = blockIdx.x, ln = threadIdx.x & 31;
    if (m >= nmat) return;
    const int lda = 2 * ml + mu + 1;
    LuEng<typename POL::Lin> e;
    e.lin.ml = ml;
    e.lin.mu = mu;
    e.lin.ld = POL::band_ld(ml, mu);
    e.lin.abd = smem;
    e.lin.wjs = nullptr;
    e.lin.inv = smem + ((POL::image_doubles(ml, mu, false) + 1) & ~1);  // fast build: the inverse, next to the band array
    if (ln < N)
        for (int i = 0; i < lda; i++) e.lin.abd[ln * e.lin.ld + i] = a[((size_t)m * N + ln) * lda + i];
    POL::lin_init(e);
    __syncwarp();
    const int inf = POL::factor(e.lin);
    real xv[1] = {(ln < N) ? b[(size_t)m * N + ln] : BV_R(0.0)};
    if (inf == 0) POL::solve(e, xv);
    __syncwarp();
    if (ln < N) {
        for (int i = 0; i < lda; i++) afac[((size_t)m * N + ln) * lda + i] = e.lin.abd[ln * e.lin.ld + i];
        ipvt[(size_t)m * N + ln] = e.lin.ipv + 1;
        x[(size_t)m * N + ln] = (inf == 0) ? xv[0] : BV_R(0.0);
    }
    if (ln == 0) info[m] = inf;
}

template <int N>
static int launch_dense(int nmat, const real *a, const real *b, real *afac, int *ipvt, real *x, int *info) {
    using POL = WarpDensePolicy<LuProb<N>, 2, false, 2>;
    k_lu_dense<N><<<nmat, 32, sizeof(real) * POL::kBlockDoubles>>>(nmat, a, b, afac, ipvt, x, info);
    return (int)cudaGetLastError();
}
template <int N>
static int launch_dense2(int nmat, const real *a, const real *b, real *afac, int *ipvt, real *x, int *info) {
    using POL = WarpDense2Policy<LuProb<N>, 2, false, 2>;
    const size_t smem = sizeof(real) * POL::kBlockDoubles;
    cudaFuncSetAttribute(k_lu_dense2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_lu_dense2<N><<<nmat, 32, smem>>>(nmat, a, b, afac, ipvt, x, info);
    return (int)cudaGetLastError();
}

// all pointers: device.  Returns 0, -1000 (this kind / n is not instantiated) or a cudaError_t.
static int lu_test(int kind, int n, int ml, int mu, int nmat, const real *a, const real *b, real *afac, int *ipvt,
                   real *x, int *info) {
    if (kind == 0) {
        const int grid = (nmat + 63) / 64;
        if (n == 2) k_lu_thread<2><<<grid, 64>>>(nmat, a, b, afac, ipvt, x, info);
        else if (n == 3) k_lu_thread<3><<<grid, 64>>>(nmat, a, b, afac, ipvt, x, info);
        else if (n == 4) k_lu_thread<4><<<grid, 64>>>(nmat, a, b, afac, ipvt, x, info);
        else return -1000;
        return (int)cudaGetLastError();
    }
    if (kind == 1) {
        if (n == 5) return launch_dense<5>(nmat, a, b, afac, ipvt, x, info);
        if (n == 25) return launch_dense<25>(nmat, a, b, afac, ipvt, x, info);
        if (n == 32) return launch_dense<32>(nmat, a, b, afac, ipvt, x, info);
        return -1000;
    }
    if (kind == 2) {
        if (n == 48) return launch_dense2<48>(nmat, a, b, afac, ipvt, x, info);
        if (n == 64) return launch_dense2<64>(nmat, a, b, afac, ipvt, x, info);
        return -1000;
    }
    if (kind == 3) {
        if (n != 25 || ml < 0 || mu < 0 || ml >= n || mu >= n || ml + mu + 1 > 32) return -1000;
        using POL = WarpBandPolicy<LuProb<25>, 4, false, 2>;
        k_lu_band<25><<<nmat, 32, sizeof(real) * POL::smem_doubles(ml, mu, false)>>>(nmat, ml, mu, a, b, afac, ipvt, x, info);
        return (int)cudaGetLastError();
    }
    return -1000;
}

}  // namespace BVODE_NS

#if BVODE_STRICT
extern "C" int bvode_lutest_strict(int kind, int n, int ml, int mu, int nmat, const real *a, const real *b,
                                   real *afac, int *ipvt, real *x, int *info)
#else
extern "C" int bvode_lutest_fast(int kind, int n, int ml, int mu, int nmat, const real *a, const real *b,
                                 real *afac, int *ipvt, real *x, int *info)
#endif
{
    return BVODE_NS::lu_test(kind, n, ml, mu, nmat, a, b, afac, ipvt, x, info);
}
