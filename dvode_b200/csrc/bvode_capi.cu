// bvode_capi.cu -- the C ABI declared in include/bvode.h: handle management, argument
// checking (the reference's blocks A and B, M:1073-1303, done once for the batch on the
// host), host <-> device staging, and kernel dispatch.  No integration arithmetic here.
#include "../../include/bvode.h"
#include "bvode_internal.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

// one entry per (build, problem), each from its own translation unit (bvode_registry.cu)
#include <mutex>
#define BVODE_NPROBLEMS 8
extern "C" {
const ProblemEntry *bvode_entry_fast_0(void);
const ProblemEntry *bvode_entry_fast_1(void);
const ProblemEntry *bvode_entry_fast_2(void);
const ProblemEntry *bvode_entry_fast_3(void);
const ProblemEntry *bvode_entry_fast_4(void);
const ProblemEntry *bvode_entry_fast_5(void);
const ProblemEntry *bvode_entry_fast_6(void);
const ProblemEntry *bvode_entry_fast_7(void);
const ProblemEntry *bvode_entry_strict_0(void);
const ProblemEntry *bvode_entry_strict_1(void);
const ProblemEntry *bvode_entry_strict_2(void);
const ProblemEntry *bvode_entry_strict_3(void);
const ProblemEntry *bvode_entry_strict_4(void);
const ProblemEntry *bvode_entry_strict_5(void);
const ProblemEntry *bvode_entry_strict_6(void);
const ProblemEntry *bvode_entry_strict_7(void);
}
// The problem table: the built-in problems, then whatever bvode_register_problem appended (user
// f / jac functors compiled into a library of their own against csrc/bvode_user_problem.cuh).
// Fixed capacity: handles keep pointers into it.
constexpr int kMaxProblems = 64;
static ProblemEntry g_tab_fast[kMaxProblems], g_tab_strict[kMaxProblems];
static int g_nproblems = 0;
static std::mutex g_tab_mutex;
static void registry_init() {
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (g_nproblems != 0) return;
#if BVODE_REAL_IS_FLOAT
    // REAL32 (libbvode32.so): strict arithmetic, thread-per-system problems; one table serves both flags
    const ProblemEntry *ent[4] = {bvode_entry_strict_0(), bvode_entry_strict_1(), bvode_entry_strict_4(), bvode_entry_strict_7()};
    for (int i = 0; i < 4; i++) g_tab_fast[i] = g_tab_strict[i] = *ent[i];
    g_nproblems = 4;
#else
    g_tab_fast[0] = *bvode_entry_fast_0(); g_tab_fast[1] = *bvode_entry_fast_1();
    g_tab_fast[2] = *bvode_entry_fast_2(); g_tab_fast[3] = *bvode_entry_fast_3();
    g_tab_fast[4] = *bvode_entry_fast_4(); g_tab_fast[5] = *bvode_entry_fast_5();
    g_tab_fast[6] = *bvode_entry_fast_6(); g_tab_fast[7] = *bvode_entry_fast_7();
    g_tab_strict[0] = *bvode_entry_strict_0(); g_tab_strict[1] = *bvode_entry_strict_1();
    g_tab_strict[2] = *bvode_entry_strict_2(); g_tab_strict[3] = *bvode_entry_strict_3();
    g_tab_strict[4] = *bvode_entry_strict_4(); g_tab_strict[5] = *bvode_entry_strict_5();
    g_tab_strict[6] = *bvode_entry_strict_6(); g_tab_strict[7] = *bvode_entry_strict_7();
    g_nproblems = BVODE_NPROBLEMS;
#endif
}
static const ProblemEntry *bvode_registry_fast(int *n) {
    registry_init();
    *n = g_nproblems;
    return g_tab_fast;
}
static const ProblemEntry *bvode_registry_strict(int *n) {
    registry_init();
    *n = g_nproblems;
    return g_tab_strict;
}

static thread_local std::string g_err;
const char *bvode_last_error(void) { return g_err.c_str(); }
void bvode_set_error(const std::string &msg) { g_err = msg; }  // for bvode_multi.cu

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
            return BVODE_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

#include "bvode_handle.h"

// ---------------------------------------------------------------- small kernels ----
// host layout [nsys][n]  <->  device layout [n][nsys]
__global__ void k_aos_to_soa(const real *__restrict__ aos, real *__restrict__ soa, int nsys,
                             int n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    for (int i = 0; i < n; i++) soa[(size_t)i * nsys + s] = aos[(size_t)s * n + i];
}
__global__ void k_soa_to_aos(const real *__restrict__ soa, real *__restrict__ aos, int nsys,
                             int n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    for (int i = 0; i < n; i++) aos[(size_t)s * n + i] = soa[(size_t)i * nsys + s];
}
// the same for the systems [s0, s0 + count) of a batch of nsys, nvec consecutive [n][nsys] arrays
// (grid.y = which array): the chunks of the pipelined host-buffer path
__global__ void k_aos_to_soa_range(const real *__restrict__ aos, real *__restrict__ soa, int nsys,
                                   int n, int s0, int count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const size_t g = (size_t)s0 + s;
    for (int i = 0; i < n; i++) soa[(size_t)i * nsys + g] = aos[g * n + i];
}
__global__ void k_soa_to_aos_range(const real *__restrict__ soa, real *__restrict__ aos, int nsys,
                                   int n, int s0, int count, int nfull = 0) {
    // nfull: rows of one [.][nsys] array on the device when the host layout carries only the first n of them
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const size_t g = (size_t)s0 + s, v = blockIdx.y;
    const real *src = soa + v * (size_t)(nfull > 0 ? nfull : n) * nsys;
    real *dst = aos + v * (size_t)n * nsys;
    for (int i = 0; i < n; i++) dst[g * n + i] = src[(size_t)i * nsys + g];
}
__global__ void k_isoa_to_aos_range(const int *__restrict__ soa, int *__restrict__ aos, int nsys,
                                    int n, int s0, int count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const size_t g = (size_t)s0 + s;
    for (int i = 0; i < n; i++) aos[g * n + i] = soa[(size_t)i * nsys + g];
}
__global__ void k_isoa_to_aos(const int *__restrict__ soa, int *__restrict__ aos, int nsys,
                              int n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    for (int i = 0; i < n; i++) aos[(size_t)s * n + i] = soa[(size_t)i * nsys + s];
}
// REALIFY-OFF
// DFMA peak: 8 independent chains per thread, 1024 threads per block, 4 blocks per SM
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,
           a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// REALIFY-ON
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// -------------------------------------------------------------------- registry ----
static const ProblemEntry *registry(unsigned flags, int *n) {
    return (flags & BVODE_FLAG_STRICT) ? bvode_registry_strict(n) : bvode_registry_fast(n);
}
int bvode_problem_count(void) {
    int n;
    bvode_registry_fast(&n);
    return n;
}
int bvode_problem_lookup(const char *name) {
    int n;
    const ProblemEntry *e = bvode_registry_fast(&n);
    for (int i = 0; i < n; i++)
        if (strcmp(e[i].name, name) == 0) return i;
    return -1;
}
int bvode_problem_info(int problem, int *neq, int *npar, const char **name) {
    int n;
    const ProblemEntry *e = bvode_registry_fast(&n);
    if (problem < 0 || problem >= n) return BVODE_EINVAL;
    if (neq) *neq = e[problem].neq;
    if (npar) *npar = e[problem].npar;
    if (name) *name = e[problem].name;
    return BVODE_OK;
}

int bvode_plugin_abi(void) { return BVODE_PLUGIN_ABI; }
int bvode_register_problem(const void *entry_fast, const void *entry_strict, int abi) {
    if (!entry_fast) return BVODE_EINVAL;
    if (abi != BVODE_PLUGIN_ABI) {
        g_err = "bvode_register_problem: the problem library was compiled against other kernel headers (ABI tag differs)";
        return BVODE_EINVAL;
    }
    registry_init();
    const ProblemEntry *f = static_cast<const ProblemEntry *>(entry_fast);
    const ProblemEntry *st = static_cast<const ProblemEntry *>(entry_strict);
    if (!f->name || !f->launch || !f->state_size || !f->supports || f->neq < 1) return BVODE_EINVAL;
    if (st && (st->neq != f->neq || st->npar != f->npar || st->mode != f->mode)) return BVODE_EINVAL;
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    for (int i = 0; i < g_nproblems; i++)
        if (strcmp(g_tab_fast[i].name, f->name) == 0) {
            g_err = "bvode_register_problem: a problem of this name is registered already";
            return BVODE_EINVAL;
        }
    if (g_nproblems >= kMaxProblems) return BVODE_ENOMEM;
    const int id = g_nproblems;
    g_tab_fast[id] = *f;
    if (st) {
        g_tab_strict[id] = *st;
    } else {  // no strict build supplied: a handle with BVODE_FLAG_STRICT is refused at create
        g_tab_strict[id] = *f;
        g_tab_strict[id].launch = nullptr;
    }
    g_nproblems = id + 1;
    return id;
}

// -------------------------------------------------------------------- lifetime ----
int bvode_create(bvode_solver **out, int problem, int nsys, int device, unsigned flags) {
    if (!out || nsys <= 0) return BVODE_EINVAL;
    int n;
    const ProblemEntry *e = registry(flags, &n);
    if (problem < 0 || problem >= n) return BVODE_EINVAL;
    if (!e[problem].launch) { g_err = "this problem was registered without a strict build"; return BVODE_ENOTIMPL; }
    CUDA_TRY(cudaSetDevice(device));
    bvode_solver *s = new bvode_solver();
    s->problem = problem;
    s->nsys = nsys;
    s->device = device;
    s->flags = flags;
    s->pe = &e[problem];
    s->neq = s->pe->neq;
    s->npar = s->pe->npar;
    const size_t ns = (size_t)nsys;
    cudaError_t err = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_params, sizeof(real) * ns * (s->npar > 0 ? s->npar : 1));
    if (err == cudaSuccess) err = cudaMalloc(&s->d_y, sizeof(real) * ns * s->neq);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_t, sizeof(real) * ns);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_tout_ps, sizeof(real) * ns);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_tol, sizeof(real) * 2 * s->neq);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_rout, sizeof(real) * 4 * ns);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_istate, sizeof(int) * ns);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_iout, sizeof(int) * 12 * ns);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_iaos, sizeof(int) * 12 * ns);
    s->aos_cap = ns * (size_t)(s->neq > s->npar ? s->neq : s->npar);
    if (s->aos_cap < 4 * ns) s->aos_cap = 4 * ns;
    if (err == cudaSuccess) err = cudaMalloc(&s->d_aos, sizeof(real) * s->aos_cap);
    if (err == cudaSuccess) err = cudaMemsetAsync(s->d_rout, 0, sizeof(real) * 4 * ns, s->stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(s->d_iout, 0, sizeof(int) * 12 * ns, s->stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(s->d_params, 0, sizeof(real) * ns * (s->npar > 0 ? s->npar : 1), s->stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(s->d_istate, 0, sizeof(int) * ns, s->stream);
    // the handle's stream does not synchronise with the caller's streams (cudaStreamNonBlocking):
    // the buffers must be ready before a first launch on any of them
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) {
        g_err = std::string("bvode_create: ") + cudaGetErrorString(err);
        bvode_destroy(s);
        return err == cudaErrorMemoryAllocation ? BVODE_ENOMEM : BVODE_ECUDA;
    }
    *out = s;
    return BVODE_OK;
}

void bvode_destroy(bvode_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_state); cudaFree(s->i_state);
    if (!s->params_external) cudaFree(s->d_params);
    cudaFree(s->d_y); cudaFree(s->d_t); cudaFree(s->d_tol); cudaFree(s->d_rout);
    cudaFree(s->d_aos); cudaFree(s->d_yout); cudaFree(s->d_tout_ps); cudaFree(s->d_istate);
    cudaFree(s->d_iout); cudaFree(s->d_iaos); cudaFree(s->d_next); cudaFree(s->d_yout_aos);
    cudaFree(s->d_tol_ps); cudaFree(s->d_optd_ps); cudaFree(s->d_opti_ps); cudaFree(s->d_rout_aos);
    cudaFree(s->d_pack);
    if (s->h_rout) cudaFreeHost(s->h_rout);
    if (s->h_iout) cudaFreeHost(s->h_iout);
    for (cudaEvent_t ev : s->chunk_done) cudaEventDestroy(ev);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    if (s->stream3) cudaStreamDestroy(s->stream3);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int bvode_set_params(bvode_solver *s, const real *params) {
    if (!s || (!params && s->npar > 0)) return BVODE_EINVAL;
    if (s->npar == 0) return BVODE_OK;
    if (s->params_external) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t ns = (size_t)s->nsys;
    CUDA_TRY(cudaMemcpyAsync(s->d_aos, params, sizeof(real) * ns * s->npar, cudaMemcpyHostToDevice, s->stream));
    k_aos_to_soa<<<cdiv(s->nsys, 256), 256, 0, s->stream>>>(s->d_aos, s->d_params, s->nsys, s->npar);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return BVODE_OK;
}

int bvode_set_params_device(bvode_solver *s, const real *params_device) {
    if (!s || !params_device) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaMemcpyAsync(s->d_params, params_device, sizeof(real) * (size_t)s->nsys * s->npar,
                             cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return BVODE_OK;
}

// ------------------------------------------------- argument checks (blocks A, B) ----
struct Checked {
    BvOpts o;
    int lenrw, leniw;
    std::vector<real> tol;  // rtol[neq] then atol[neq], expanded according to itol
};

// Work-array lengths the reference would need, M:1246-1273 (reported in iwork(17:18)).
static void work_sizes(int neq, int mf, int ml, int mu, int maxord, int *lenrw, int *leniw, int nyh = 0) {
    if (nyh <= 0) nyh = neq;  // nyh = the neq of the first call (M:1247): yh keeps its leading dimension
    const int jsv = mf >= 0 ? 1 : -1, mfa = mf < 0 ? -mf : mf, meth = mfa / 10,
              miter = mfa - 10 * meth, jco = jsv > 0 ? 1 : 0;
    int lenwm = 0;
    switch (miter) {
    case 0: lenwm = 0; break;
    case 1: case 2: lenwm = 2 + (1 + jco) * neq * neq; break;
    case 3: lenwm = 2 + neq; break;
    default: lenwm = 2 + (2 * ml + mu + 1) * neq + jco * (ml + mu + 1) * neq; break;
    }
    *lenrw = 20 + nyh * (maxord + 1) + lenwm + 3 * neq;
    *leniw = (miter == 0 || miter == 3) ? 30 : 30 + neq;
}

// The optional inputs of one solver instance (M:729-751), defaulted as in M:1174-1238: rw / iw are that
// instance's rwork(1:20) / iwork(1:30) header.  Returns nullptr, or the text of the input error.
static const char *read_optional_inputs(int meth, int itask, int iopt, const real *rw, const int *iw, BvOpts *o) {
    const int mord = (meth == 1) ? 12 : 5;
    o->tcrit = ((itask == 4 || itask == 5) && rw) ? rw[0] : BV_R(0.0);
    if (iopt == 1) {
        o->maxord = iw[4];
        if (o->maxord < 0) return "maxord < 0";
        if (o->maxord == 0) o->maxord = 100;
        if (o->maxord > mord) o->maxord = mord;
        o->mxstep = iw[5];
        if (o->mxstep < 0) return "mxstep < 0";
        if (o->mxstep == 0) o->mxstep = 500;
        o->mxhnil = iw[6];
        if (o->mxhnil < 0) return "mxhnil < 0";
        if (o->mxhnil == 0) o->mxhnil = 10;
        o->h0 = rw[4];
        const real hmax = rw[5];
        if (hmax < BV_R(0.0)) return "hmax < 0";
        o->hmax_inv = hmax > BV_R(0.0) ? BV_R(1.0) / hmax : BV_R(0.0);
        o->hmin = rw[6];
        if (o->hmin < BV_R(0.0)) return "hmin < 0";
    } else {
        o->maxord = mord;
        o->mxstep = 500;
        o->mxhnil = 10;
        o->h0 = BV_R(0.0);
        o->hmax_inv = BV_R(0.0);
        o->hmin = BV_R(0.0);
    }
    return nullptr;
}

// Returns BVODE_OK, or the error class; mirrors M:1092-1303 for the batch-uniform inputs.
static int check_args(const bvode_solver *s, int neq, int itol, const real *rtol,
                      const real *atol, int itask, int iopt, const real *rwork0,
                      const int *iwork0, int lrw, int liw, bool have_headers, int mf,
                      Checked *c, bool restart3 = false) {
    {   // neq: the problem's, or -- from an istate = 3 call on -- a smaller one (M:1127-1135; never larger again)
        const int cur = s->nact > 0 ? s->nact : s->neq;
        if (neq != cur && !(restart3 && neq >= 1 && neq < cur)) {
            g_err = (neq > cur) ? "neq larger than the problem's (or than the neq of the last istate = 3 call)"
                                : "neq may only be reduced by an istate = 3 call";
            return BVODE_EINVAL;
        }
    }
    if (itask < 1 || itask > 5) { g_err = "itask illegal"; return BVODE_EINVAL; }
    if ((itask == 4 || itask == 5) && !rwork0) { g_err = "itask 4/5 needs rwork(1) = tcrit"; return BVODE_EINVAL; }
    if (itol < 1 || itol > 4) { g_err = "itol illegal"; return BVODE_EINVAL; }
    if (iopt < 0 || iopt > 1) { g_err = "iopt illegal"; return BVODE_EINVAL; }
    if (!rtol || !atol) { g_err = "rtol/atol null"; return BVODE_EINVAL; }
    const int mfa = mf < 0 ? -mf : mf, meth = mfa / 10, miter = mfa - 10 * meth;
    if (meth < 1 || meth > 2 || miter < 0 || miter > 5) { g_err = "mf illegal"; return BVODE_EINVAL; }
    int ml = 0, mu = 0;
    if (miter > 3) {
        if (!iwork0) { g_err = "band method needs iwork(1:2) = ml, mu"; return BVODE_EINVAL; }
        ml = iwork0[0];
        mu = iwork0[1];
        if (ml < 0 || ml >= neq) { g_err = "ml illegal"; return BVODE_EINVAL; }
        if (mu < 0 || mu >= neq) { g_err = "mu illegal"; return BVODE_EINVAL; }
    }
    if (have_headers && (lrw < 20 || liw < 30)) { g_err = "lrw < 20 or liw < 30"; return BVODE_EINVAL; }
    if (!s->pe->supports(mf, ml, mu)) {
        g_err = "this mf (or ml/mu) is not built for this problem";
        return BVODE_ENOTIMPL;
    }
    BvOpts &o = c->o;
    o.itol = itol;
    o.ml = ml;
    o.mu = mu;
    o.itask = itask;
    o.jreset = 0;
    o.stash = 0;
    o.lmaxchg = 0;
    o.nact = 0;
    if (iopt == 1 && (!rwork0 || !iwork0)) { g_err = "iopt = 1 needs rwork and iwork"; return BVODE_EINVAL; }
    if (const char *msg = read_optional_inputs(meth, itask, iopt, rwork0, iwork0, &o)) { g_err = msg; return BVODE_EINVAL; }
    c->tol.resize(2 * (size_t)neq);
    for (int i = 0; i < neq; i++) {
        const real r = (itol >= 3) ? rtol[i] : rtol[0];
        const real a = (itol == 2 || itol == 4) ? atol[i] : atol[0];
        if (r < BV_R(0.0)) { g_err = "rtol < 0"; return BVODE_EINVAL; }
        if (a < BV_R(0.0)) { g_err = "atol < 0"; return BVODE_EINVAL; }
        c->tol[i] = r;
        c->tol[neq + i] = a;
    }
    o.nact = (neq < s->neq) ? neq : 0;
    work_sizes(neq, mf, ml, mu, o.maxord, &c->lenrw, &c->leniw, s->neq);
    return BVODE_OK;
}

// BVODE_FLAG_PER_SYSTEM: every system is its own solver instance as far as rtol / atol and the optional
// inputs go, as in the reference (M:729-751, 928-965): rtol[nsys][nr], atol[nsys][na] with nr, na = 1 or
// neq according to itol, and the optional inputs of system i in row i of rwork / iwork.  A system whose
// row is illegal gets istate = -3 and is skipped (M:1120-1303 does that per instance); with a
// device-resident istate the call is refused instead.  mf, ml / mu, itol, itask, iopt stay batch-uniform.
static int prepare_per_system(bvode_solver *s, const Checked &c, int neq, int itol, const real *rtol,
                              const real *atol, int itask, int iopt, const real *rwork, int lrw,
                              const int *iwork, int liw, int *istate_host, bool restart3, int mf,
                              cudaStream_t st) {
    const size_t ns = (size_t)s->nsys;
    if (iopt == 1 && (!rwork || !iwork || lrw < 20 || liw < 30)) { g_err = "iopt = 1 needs rwork[nsys][lrw >= 20] and iwork[nsys][liw >= 30]"; return BVODE_EINVAL; }
    if ((itask == 4 || itask == 5) && (!rwork || lrw < 1)) { g_err = "itask 4/5 needs rwork(1) = tcrit in every row"; return BVODE_EINVAL; }
    if (!s->d_tol_ps) {
        CUDA_TRY(cudaMalloc(&s->d_tol_ps, sizeof(real) * 2 * neq * ns));
        CUDA_TRY(cudaMalloc(&s->d_optd_ps, sizeof(real) * 4 * ns));
        CUDA_TRY(cudaMalloc(&s->d_opti_ps, sizeof(int) * 3 * ns));
    }
    const int nr = (itol >= 3) ? neq : 1, na = (itol == 2 || itol == 4) ? neq : 1;
    const int meth = abs(mf) / 10;
    std::vector<real> tol(2 * (size_t)neq * ns), od(4 * ns);
    std::vector<int> oi(3 * ns);
    const bool have_prev = s->maxord_ps.size() == ns;
    std::vector<int> maxord_new(ns);
    for (size_t i = 0; i < ns; i++) {
        bool bad = false;
        for (int k = 0; k < neq; k++) {
            const real r = rtol[i * nr + (nr > 1 ? k : 0)], a = atol[i * na + (na > 1 ? k : 0)];
            if (!(r >= BV_R(0.0)) || !(a >= BV_R(0.0))) bad = true;
            tol[(size_t)k * ns + i] = r;
            tol[(size_t)(neq + k) * ns + i] = a;
        }
        BvOpts o = c.o;
        if (read_optional_inputs(meth, itask, iopt, rwork ? rwork + i * (size_t)lrw : nullptr,
                                 iwork ? iwork + i * (size_t)liw : nullptr, &o))
            bad = true;
        if (bad) {
            if (!istate_host) { g_err = "per-system rtol / atol / optional inputs: an illegal row (and istate is on the device)"; return BVODE_EINVAL; }
            istate_host[i] = -3;
        }
        od[i] = o.h0; od[ns + i] = o.hmax_inv; od[2 * ns + i] = o.hmin; od[3 * ns + i] = o.tcrit;
        const int prev = have_prev ? s->maxord_ps[i] : s->maxord;
        oi[i] = o.maxord; oi[ns + i] = o.mxstep;
        oi[2 * ns + i] = ((c.o.jreset || (restart3 && mf > 0 && o.maxord != prev)) ? 1 : 0) |
                         ((restart3 && (c.o.lmaxchg || o.maxord != prev)) ? 2 : 0);
        maxord_new[i] = o.maxord;
    }
    s->maxord_ps.swap(maxord_new);
    CUDA_TRY(cudaMemcpyAsync(s->d_tol_ps, tol.data(), sizeof(real) * tol.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->d_optd_ps, od.data(), sizeof(real) * od.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->d_opti_ps, oi.data(), sizeof(int) * oi.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // the vectors go out of scope
    return BVODE_OK;
}

static int ensure_state(bvode_solver *s, int mf, int ml, int mu) {
    if (s->d_state && s->mf == mf && s->ml == ml && s->mu == mu) return BVODE_OK;
    cudaFree(s->d_state);
    cudaFree(s->i_state);
    s->d_state = nullptr;
    s->i_state = nullptr;
    s->pe->state_size(mf, ml, mu, (size_t)s->nsys, &s->nd, &s->ni);
    CUDA_TRY(cudaMalloc(&s->d_state, sizeof(real) * s->nd));
    CUDA_TRY(cudaMalloc(&s->i_state, sizeof(int) * s->ni));
    CUDA_TRY(cudaMemsetAsync(s->d_state, 0, sizeof(real) * s->nd, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->i_state, 0, sizeof(int) * s->ni, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->mf = mf;
    s->ml = ml;
    s->mu = mu;
    return BVODE_OK;
}

// Launch the integration for touts[0..nout) on device-resident SoA buffers.
static int run_device(bvode_solver *s, const Checked &c, int mf, real *d_y, real *d_t,
                      int *d_istate, int nout, const real *touts, const real *d_tout_ps,
                      real *d_yout, cudaStream_t st, bool general, int sys0 = 0, int count = -1,
                      int lane = 0, bool upload_tol = true) {
    if (count < 0) count = s->nsys;
    if (nout > 1 && !(c.o.itask == 1 || c.o.itask == 4)) {
        g_err = "a tout schedule (nout > 1) needs itask 1 or 4: itask 2, 3, 5 return at mesh points, one call at a time";
        return BVODE_EINVAL;
    }
    int rc = ensure_state(s, mf, c.o.ml, c.o.mu);
    if (rc) return rc;
    s->last_y = d_y;
    s->last_t = d_t;
    s->last_istate = d_istate;
    s->last_stream = st;
    if (upload_tol)
        CUDA_TRY(cudaMemcpyAsync(s->d_tol, c.tol.data(), sizeof(real) * c.tol.size(),
                                 cudaMemcpyHostToDevice, st));
    for (int k0 = 0; k0 < nout; k0 += BVODE_MAX_NOUT) {
        KArgs a;
        memset(&a, 0, sizeof(a));
        a.nsys = s->nsys;
        a.sys0 = sys0;
        a.count = count;
        a.nout = (nout - k0 < BVODE_MAX_NOUT) ? nout - k0 : BVODE_MAX_NOUT;
        for (int k = 0; k < a.nout; k++) a.touts[k] = touts[k0 + k];
        a.tout_ps = d_tout_ps;
        a.dstate = s->d_state;
        a.istate_store = s->i_state;
        a.params = s->d_params;
        a.y = d_y;
        a.t = d_t;
        a.istate = d_istate;
        a.rtol = s->d_tol;
        a.atol = s->d_tol + s->neq;
        for (int i = 0; i < s->neq && i < 8; i++) {
            a.rtolv[i] = c.tol[i];
            a.atolv[i] = c.tol[s->neq + i];
        }
        a.o = c.o;
        a.y_out = d_yout ? d_yout + (size_t)k0 * s->neq * s->nsys : nullptr;
        a.rout = s->d_rout;
        a.iout = s->d_iout;
        a.lenrw = c.lenrw;
        a.leniw = c.leniw;
        a.general = general ? 1 : 0;
        if (s->flags & BVODE_FLAG_PER_SYSTEM) {  // read by the full-driver kernels
            a.tol_ps = s->d_tol_ps;
            a.optd_ps = s->d_optd_ps;
            a.opti_ps = s->d_opti_ps;
            a.general = 1;
        }
        if (s->pe->mode == 0 && (s->flags & BVODE_FLAG_COMPACT) && !a.general) {
            if (!s->d_next) CUDA_TRY(cudaMalloc(&s->d_next, 4 * sizeof(int)));
            CUDA_TRY(cudaMemsetAsync(s->d_next + lane, 0, sizeof(int), st));
            a.next_sys = s->d_next + lane;
        }
        const int e = s->pe->launch(mf, a, st);
        if (e == -1000) { g_err = "mf not built for this problem"; return BVODE_ENOTIMPL; }
        if (e != 0) { g_err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)e); return BVODE_ECUDA; }
        s->last_launches++;
    }
    return BVODE_OK;
}

static int solve_host(bvode_solver *s, int neq, real *y, real *t, int nout,
                      const real *touts, const real *tout_ps, int itol, const real *rtol,
                      const real *atol, int itask, int *istate, int iopt, real *rwork,
                      int lrw, int *iwork, int liw, int mf, real *y_out, long long yout_sys_stride = 0) {
    if (!s || !y || !t || !istate || nout < 1 || (!touts && !tout_ps)) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    const int nsys = s->nsys;
    const size_t ns = (size_t)nsys;
    s->last_launches = 0;
    // block A (M:1073-1109): istate legality, per system
    bool any_first = false;
    int n3 = 0;
    for (int i = 0; i < nsys; i++) {
        if (istate[i] == 1) any_first = true;
        else if (istate[i] == 3) n3++;
        else if (istate[i] < 0) { g_err = "negative istate on input (the reference aborts the run here)"; return BVODE_EISTATE; }
        else if (istate[i] != 2) { for (int j = 0; j < nsys; j++) istate[j] = -3; g_err = "istate illegal"; return BVODE_EINVAL; }
    }
    // istate = 3 (M:1382-1390): the optional inputs are batch-uniform, so a restart with changed
    // parameters applies to the whole batch; mf / ml / mu select the kernel and the state layout
    // and cannot change without a new handle
    if (n3 != 0 && n3 != nsys) { g_err = "istate = 3 must be given for every system of the batch"; return BVODE_EINVAL; }
    if (n3 != 0 && !s->d_state) { g_err = "istate = 3 before a successful first call"; return BVODE_EINVAL; }
    const bool mf_changed = (n3 != 0 && s->mf != mf);
    const bool array_shrinks = mf_changed && (abs(s->mf) / 10 == 1) && (abs(mf) / 10 == 2);  // Adams -> BDF: lmax 13 -> 6
    Checked c;
    const bool have_headers = rwork && iwork;
    int rc = check_args(s, neq, itol, rtol, atol, itask, iopt, rwork, iwork, lrw, liw, have_headers, mf, &c, n3 != 0);
    if (rc != BVODE_OK) {
        if (rc == BVODE_EINVAL) for (int j = 0; j < nsys; j++) istate[j] = -3;
        return rc;
    }
    // ml / mu are re-read on every istate = 3 call (M:1136-1148) and only mean something for the banded method
    // flags (miter 4, 5): another bandwidth is another band image in the state, handled like another mf below
    bool band_changed = false;
    {
        const int mo = abs(s->mf) % 10, mn = abs(mf) % 10;
        const bool band_old = (mo == 4 || mo == 5), band_new = (mn == 4 || mn == 5);
        band_changed = (n3 != 0 && band_old && band_new && (c.o.ml != s->ml || c.o.mu != s->mu));
    }
    if (!(s->flags & BVODE_FLAG_PER_SYSTEM) && n3 != 0 && mf > 0 && c.o.maxord != s->maxord) c.o.jreset = 1;
    if (n3 != 0 && (c.o.maxord != s->maxord || abs(s->mf) / 10 != abs(mf) / 10)) c.o.lmaxchg = 1;
    if (n3 != 0 && c.o.nact != s->nact) c.o.jreset = 1;  // a smaller neq: the saved Jacobian had another leading dimension
    if (mf_changed || band_changed) {
        // M:999-1046 at istate = 3: a new miter / jsv (or, M:1136-1148, another ml / mu).  The state moves into the layout of the new
        // method flag; the Newton matrix and the saved Jacobian are rebuilt at the restart.
        size_t nd = 0, ni = 0;
        s->pe->state_size(mf, c.o.ml, c.o.mu, (size_t)nsys, &nd, &ni);
        real *dn = nullptr;
        int *in = nullptr;
        CUDA_TRY(cudaMalloc(&dn, sizeof(real) * nd));
        CUDA_TRY(cudaMalloc(&in, sizeof(int) * ni));
        CUDA_TRY(cudaMemsetAsync(dn, 0, sizeof(real) * nd, s->stream));
        CUDA_TRY(cudaMemsetAsync(in, 0, sizeof(int) * ni, s->stream));
        const int e = s->pe->state_convert ? s->pe->state_convert(s->mf, mf, c.o.ml, c.o.mu, (size_t)nsys, s->d_state,
                                                                  s->i_state, dn, in, s->stream)
                                           : -1000;
        if (e != 0) {
            cudaFree(dn);
            cudaFree(in);
            g_err = "istate = 3: this change of mf is not built";
            return BVODE_ENOTIMPL;
        }
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        cudaFree(s->d_state);
        cudaFree(s->i_state);
        s->d_state = dn;
        s->i_state = in;
        s->nd = nd;
        s->ni = ni;
        s->mf = mf;
        s->ml = c.o.ml;
        s->mu = c.o.mu;
        c.o.jreset = 1;
        c.o.stash = array_shrinks ? 1 : 0;
    }
    if (!any_first && (s->mf != mf || !s->d_state)) {
        for (int j = 0; j < nsys; j++) istate[j] = -3;  // M:1102-1106
        g_err = "istate = 2 but the solver was not initialised with this mf";
        return BVODE_EINVAL;
    }
    cudaStream_t st = s->stream;
    const int TB = 256;
    const real *d_tout_ps = tout_ps ? s->d_tout_ps : nullptr;
    real *d_yout = nullptr;
    if (y_out) {
        const size_t need = ns * (size_t)s->neq * nout;  // the kernels write all neq rows of every output point
        if (need > s->yout_cap) {
            cudaFree(s->d_yout);
            s->d_yout = nullptr;
            s->yout_cap = 0;
            CUDA_TRY(cudaMalloc(&s->d_yout, sizeof(real) * need));
            s->yout_cap = need;
        }
        if (need > s->yout_aos_cap) {
            cudaFree(s->d_yout_aos);
            s->d_yout_aos = nullptr;
            s->yout_aos_cap = 0;
            CUDA_TRY(cudaMalloc(&s->d_yout_aos, sizeof(real) * need));
            s->yout_aos_cap = need;
        }
        d_yout = s->d_yout;
    }
    rc = ensure_state(s, mf, c.o.ml, c.o.mu);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->d_tol, c.tol.data(), sizeof(real) * c.tol.size(), cudaMemcpyHostToDevice, st));
    if (s->flags & BVODE_FLAG_PER_SYSTEM) {
        rc = prepare_per_system(s, c, neq, itol, rtol, atol, itask, iopt, rwork, lrw, iwork, liw, istate, n3 != 0, mf, st);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    // optional outputs (M:1670-1723) travel with their chunk: SoA -> [nsys][4] / [nsys][12] on the device,
    // one copy into pinned staging, and the host scatters a chunk's rows into the caller's rwork / iwork
    // headers while the GPU works on the next chunks
    if (have_headers) {
        if (!s->d_rout_aos) CUDA_TRY(cudaMalloc(&s->d_rout_aos, sizeof(real) * 4 * ns));
        if (!s->h_rout) CUDA_TRY(cudaMallocHost(&s->h_rout, sizeof(real) * 4 * ns));
        if (!s->h_iout) CUDA_TRY(cudaMallocHost(&s->h_iout, sizeof(int) * 12 * ns));
    }
    // The batch goes through in chunks on three streams: the copies of one chunk (inputs in, every
    // output point out) run under the kernels of its neighbours, and a chunk's kernel fills the
    // SMs the previous one's tail leaves idle.  One chunk for small batches.
    // (measured on 2^20 Robertson systems, e2e ms: 2^15 / 2^16 / 2^17 / 2^18 systems per chunk on 3 streams 115.1 / 115.7 /
    // 117.3 / 123.4, on 2 streams 116.5 (2^16) / 117.9 (2^17); 2^14: 139 -- launch-bound; the kernel alone: 110.4)
    int chunk_log2 = 16, nstreams = 3;
    if (const char *ev = getenv("BVODE_CHUNK_LOG2")) chunk_log2 = atoi(ev);   // experiments
    if (const char *ev = getenv("BVODE_HOST_STREAMS")) nstreams = atoi(ev);
    if (chunk_log2 < 10) chunk_log2 = 10;
    if (chunk_log2 > 24) chunk_log2 = 24;
    if (nstreams < 1) nstreams = 1;
    if (nstreams > 3) nstreams = 3;
    const int kChunk = 1 << chunk_log2;
    const int nchunks = (nsys >= 2 * kChunk && nout <= BVODE_MAX_NOUT) ? cdiv(nsys, kChunk) : 1;
    if (nchunks > 1 && !s->stream2) CUDA_TRY(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    if (nchunks > 1 && !s->stream3) CUDA_TRY(cudaStreamCreateWithFlags(&s->stream3, cudaStreamNonBlocking));
    const real tout1 = BV_R(0.0);
    for (int ch = 0; ch < nchunks; ch++) {
        const int s0 = (int)((long long)nsys * ch / nchunks), s1 = (int)((long long)nsys * (ch + 1) / nchunks);
        const int cnt = s1 - s0;
        const int lane = ch % nstreams;
        cudaStream_t cs = lane == 0 ? s->stream : (lane == 1 ? s->stream2 : s->stream3);
        // host -> device.  On continuation calls y and t are outputs only (M:608-618).
        CUDA_TRY(cudaMemcpyAsync(s->d_istate + s0, istate + s0, sizeof(int) * cnt, cudaMemcpyHostToDevice, cs));
        if (any_first) {
            CUDA_TRY(cudaMemcpyAsync(s->d_aos + (size_t)s0 * neq, y + (size_t)s0 * neq, sizeof(real) * cnt * neq,
                                     cudaMemcpyHostToDevice, cs));
            k_aos_to_soa_range<<<cdiv(cnt, TB), TB, 0, cs>>>(s->d_aos, s->d_y, nsys, neq, s0, cnt);
            s->last_launches++;
            CUDA_TRY(cudaMemcpyAsync(s->d_t + s0, t + s0, sizeof(real) * cnt, cudaMemcpyHostToDevice, cs));
        }
        if (tout_ps)
            CUDA_TRY(cudaMemcpyAsync(s->d_tout_ps + s0, tout_ps + s0, sizeof(real) * cnt, cudaMemcpyHostToDevice, cs));
        rc = run_device(s, c, mf, s->d_y, s->d_t, s->d_istate, nout, touts ? touts : &tout1, d_tout_ps, d_yout, cs,
                        itask != 1 || n3 != 0 || c.o.nact != 0, s0, cnt, lane, false);
        if (rc != BVODE_OK) {  // nothing may still be writing into the caller's buffers when we return
            cudaStreamSynchronize(s->stream);
            if (s->stream2) cudaStreamSynchronize(s->stream2);
            if (s->stream3) cudaStreamSynchronize(s->stream3);
            return rc;
        }
        // device -> host
        k_soa_to_aos_range<<<dim3(cdiv(cnt, TB), 1), TB, 0, cs>>>(s->d_y, s->d_aos, nsys, neq, s0, cnt);
        s->last_launches++;
        CUDA_TRY(cudaMemcpyAsync(y + (size_t)s0 * neq, s->d_aos + (size_t)s0 * neq, sizeof(real) * cnt * neq,
                                 cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(t + s0, s->d_t + s0, sizeof(real) * cnt, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(istate + s0, s->d_istate + s0, sizeof(int) * cnt, cudaMemcpyDeviceToHost, cs));
        if (have_headers) {
            k_soa_to_aos_range<<<dim3(cdiv(cnt, TB), 1), TB, 0, cs>>>(s->d_rout, s->d_rout_aos, nsys, 4, s0, cnt);
            k_isoa_to_aos_range<<<cdiv(cnt, TB), TB, 0, cs>>>(s->d_iout, s->d_iaos, nsys, 12, s0, cnt);
            s->last_launches += 2;
            CUDA_TRY(cudaMemcpyAsync(s->h_rout + (size_t)s0 * 4, s->d_rout_aos + (size_t)s0 * 4, sizeof(real) * 4 * cnt,
                                     cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaMemcpyAsync(s->h_iout + (size_t)s0 * 12, s->d_iaos + (size_t)s0 * 12, sizeof(int) * 12 * cnt,
                                     cudaMemcpyDeviceToHost, cs));
        }
        if (y_out) {
            k_soa_to_aos_range<<<dim3(cdiv(cnt, TB), nout), TB, 0, cs>>>(d_yout, s->d_yout_aos, nsys, neq, s0, cnt, s->neq);
            s->last_launches++;
            const size_t hs = yout_sys_stride > 0 ? (size_t)yout_sys_stride : ns;  // systems per output slab of y_out
            for (int k = 0; k < nout; k++) {
                const size_t off = ((size_t)k * ns + s0) * neq, hoff = ((size_t)k * hs + s0) * neq;
                CUDA_TRY(cudaMemcpyAsync(y_out + hoff, s->d_yout_aos + off, sizeof(real) * cnt * neq,
                                         cudaMemcpyDeviceToHost, cs));
            }
        }
        if (have_headers) {
            while ((int)s->chunk_done.size() <= ch) {
                cudaEvent_t ev;
                CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                s->chunk_done.push_back(ev);
            }
            CUDA_TRY(cudaEventRecord(s->chunk_done[ch], cs));
        }
    }
    s->maxord = c.o.maxord;
    s->mxstep = c.o.mxstep;
    s->nact = c.o.nact;
    if (have_headers) {
        for (int ch = 0; ch < nchunks; ch++) {
            const size_t s0 = (size_t)((long long)nsys * ch / nchunks), s1 = (size_t)((long long)nsys * (ch + 1) / nchunks);
            CUDA_TRY(cudaEventSynchronize(s->chunk_done[ch]));
            for (size_t i = s0; i < s1; i++) {
                memcpy(rwork + i * lrw + 10, s->h_rout + i * 4, sizeof(real) * 4);
                memcpy(iwork + i * liw + 10, s->h_iout + i * 12, sizeof(int) * 12);
            }
        }
    }
    if (s->stream2) CUDA_TRY(cudaStreamSynchronize(s->stream2));
    if (s->stream3) CUDA_TRY(cudaStreamSynchronize(s->stream3));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return BVODE_OK;
}

int bvode_solve(bvode_solver *s, int neq, real *y, real *t, const real *tout,
                int tout_per_system, int itol, const real *rtol, const real *atol, int itask,
                int *istate, int iopt, real *rwork, int lrw, int *iwork, int liw, int mf) {
    if (!tout) return BVODE_EINVAL;
    return solve_host(s, neq, y, t, 1, tout_per_system ? nullptr : tout,
                      tout_per_system ? tout : nullptr, itol, rtol, atol, itask, istate, iopt,
                      rwork, lrw, iwork, liw, mf, nullptr);
}

int bvode_solve_schedule(bvode_solver *s, int neq, real *y, real *t, int nout,
                         const real *touts, int itol, const real *rtol, const real *atol,
                         int itask, int *istate, int iopt, real *rwork, int lrw, int *iwork,
                         int liw, int mf, real *y_out) {
    return solve_host(s, neq, y, t, nout, touts, nullptr, itol, rtol, atol, itask, istate, iopt,
                      rwork, lrw, iwork, liw, mf, y_out);
}

// Device-resident entry points: istate lives in HBM, so the continuation checks of M:1102-1106 cannot be
// made per system here.  A handle whose state exists is therefore bound to its mf / ml / mu: a call
// with other values would silently re-initialise systems that continue with istate = 2, and is refused
// (bvode_reset drops the state when a new problem is meant).
static int device_path_prepare(bvode_solver *s, const Checked &c, int neq, int itol, const real *rtol,
                               const real *atol, int itask, int iopt, const real *rwork_opts,
                               const int *iwork_opts, int mf, cudaStream_t st) {
    if (s->d_state && (s->mf != mf || s->ml != c.o.ml || s->mu != c.o.mu)) {
        g_err = "device-resident solve with another mf / ml / mu than the handle's state was built for "
                "(call bvode_reset first, or use the host-buffer entry points for an istate = 3 change)";
        return BVODE_EINVAL;
    }
    if (s->flags & BVODE_FLAG_PER_SYSTEM)
        return prepare_per_system(s, c, neq, itol, rtol, atol, itask, iopt, rwork_opts, 20, iwork_opts, 30,
                                  nullptr, false, mf, st ? st : s->stream);
    return BVODE_OK;
}

int bvode_reset(bvode_solver *s) {
    if (s) s->nact = 0;
    if (!s) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(s->d_state);
    cudaFree(s->i_state);
    s->d_state = nullptr;
    s->i_state = nullptr;
    s->nd = s->ni = 0;
    s->mf = s->ml = s->mu = s->maxord = 0;
    s->maxord_ps.clear();
    return BVODE_OK;
}

int bvode_solve_schedule_strided(bvode_solver *s, int neq, real *y, real *t, int nout, const real *touts,
                                 const real *tout_ps, int itol, const real *rtol, const real *atol, int itask,
                                 int *istate, int iopt, real *rwork, int lrw, int *iwork, int liw, int mf,
                                 real *y_out, long long yout_sys_stride) {
    return solve_host(s, neq, y, t, nout, touts, tout_ps, itol, rtol, atol, itask, istate, iopt, rwork, lrw, iwork,
                      liw, mf, y_out, yout_sys_stride);
}

int bvode_solve_schedule_device(bvode_solver *s, int neq, real *y_device, real *t_device,
                                int nout, const real *touts, int itol, const real *rtol,
                                const real *atol, int itask, int *istate_device, int iopt,
                                const real *rwork_opts, const int *iwork_opts, int mf,
                                real *y_out_device, void *stream) {
    if (!s || !y_device || !t_device || !istate_device || !touts || nout < 1) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    s->last_launches = 0;
    Checked c;
    int rc = check_args(s, neq, itol, rtol, atol, itask, iopt, rwork_opts, iwork_opts, 20, 30,
                        false, mf, &c);
    if (rc != BVODE_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;  // 0 = the legacy default stream, as documented
    // (device-resident istate: istate = 3 is only accepted through the host-buffer entry points)
    rc = device_path_prepare(s, c, neq, itol, rtol, atol, itask, iopt, rwork_opts, iwork_opts, mf, st);
    if (rc != BVODE_OK) return rc;
    rc = run_device(s, c, mf, y_device, t_device, istate_device, nout, touts, nullptr,
                    y_out_device, st, itask != 1 || c.o.nact != 0);
    if (rc != BVODE_OK) return rc;
    s->maxord = c.o.maxord;
    s->mxstep = c.o.mxstep;
    CUDA_TRY(cudaGetLastError());
    return BVODE_OK;
}

int bvode_solve_device(bvode_solver *s, int neq, real *y_device, real *t_device,
                       const real *tout_device, int itol, const real *rtol, const real *atol,
                       int itask, int *istate_device, int iopt, const real *rwork_opts,
                       const int *iwork_opts, int mf, void *stream) {
    if (!s || !y_device || !t_device || !istate_device || !tout_device) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    s->last_launches = 0;
    Checked c;
    int rc = check_args(s, neq, itol, rtol, atol, itask, iopt, rwork_opts, iwork_opts, 20, 30,
                        false, mf, &c);
    if (rc != BVODE_OK) return rc;
    const real tout1 = BV_R(0.0);  // unused: every system has its own tout
    rc = device_path_prepare(s, c, neq, itol, rtol, atol, itask, iopt, rwork_opts, iwork_opts, mf, (cudaStream_t)stream);
    if (rc != BVODE_OK) return rc;
    rc = run_device(s, c, mf, y_device, t_device, istate_device, 1, &tout1, tout_device, nullptr,
                    (cudaStream_t)stream, itask != 1 || c.o.nact != 0);
    if (rc != BVODE_OK) return rc;
    s->maxord = c.o.maxord;
    s->mxstep = c.o.mxstep;
    CUDA_TRY(cudaGetLastError());
    return BVODE_OK;
}

int bvode_last_launch_count(const bvode_solver *s) { return s ? s->last_launches : 0; }

int bvode_get_stats(bvode_solver *s, real *rout, int *iout) {
    if (!s) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t ns = (size_t)s->nsys;
    const int TB = 256;
    if (rout) {
        k_soa_to_aos<<<cdiv(s->nsys, TB), TB, 0, s->stream>>>(s->d_rout, s->d_aos, s->nsys, 4);
        CUDA_TRY(cudaMemcpyAsync(rout, s->d_aos, sizeof(real) * 4 * ns, cudaMemcpyDeviceToHost, s->stream));
    }
    if (iout) {
        k_isoa_to_aos<<<cdiv(s->nsys, TB), TB, 0, s->stream>>>(s->d_iout, s->d_iaos, s->nsys, 12);
        CUDA_TRY(cudaMemcpyAsync(iout, s->d_iaos, sizeof(int) * 12 * ns, cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return BVODE_OK;
}

int bvode_dvindy(bvode_solver *s, const real *t, int t_per_system, int k, real *dky,
                 int *iflag) {
    if (!s || !t || !dky || !iflag) return BVODE_EINVAL;
    if (!s->d_state) { g_err = "dvindy before any successful solve"; return BVODE_EINVAL; }
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t ns = (size_t)s->nsys;
    const int TB = 256;
    cudaStream_t st = s->stream;
    DkyArgs a;
    memset(&a, 0, sizeof(a));
    a.nsys = s->nsys;
    a.k = k;
    a.ml = s->ml;
    a.mu = s->mu;
    if (t_per_system) {
        CUDA_TRY(cudaMemcpyAsync(s->d_tout_ps, t, sizeof(real) * ns, cudaMemcpyHostToDevice, st));
        a.t_ps = s->d_tout_ps;
    } else {
        a.t = t[0];
    }
    a.dstate = s->d_state;
    a.istate_store = s->i_state;
    // dky goes through d_y-sized scratch: reuse d_yout (allocated on demand)
    const size_t need = ns * s->neq;
    if (need > s->yout_cap) {
        cudaFree(s->d_yout);
        s->d_yout = nullptr;
        s->yout_cap = 0;
        CUDA_TRY(cudaMalloc(&s->d_yout, sizeof(real) * need));
        s->yout_cap = need;
    }
    a.dky = s->d_yout;
    a.iflag = s->d_iaos;
    const int e = s->pe->launch_dky(s->mf, a, st);
    if (e != 0) { g_err = "dvindy kernel launch failed"; return e == -1000 ? BVODE_ENOTIMPL : BVODE_ECUDA; }
    const int nhost = s->nact > 0 ? s->nact : s->neq;  // dky[nsys][neq of the last call]
    k_soa_to_aos<<<cdiv(s->nsys, TB), TB, 0, st>>>(s->d_yout, s->d_aos, s->nsys, nhost);
    CUDA_TRY(cudaMemcpyAsync(dky, s->d_aos, sizeof(real) * ns * nhost, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(iflag, s->d_iaos, sizeof(int) * ns, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return BVODE_OK;
}

// ---- dvsrco (M:3198-3216): the whole batch state as one host blob ------------------
// header, then: real state, integer state, optional outputs (rwork(11:14), iwork(11:22)), and the
// handle's own copy of istate / t / y of the last host-buffer call -- everything a fresh handle needs
// to continue with istate = 2 (or 3) and to answer bvode_istate_summary / bvode_format_message.
struct SavHeader {
    int magic, problem, nsys, mf, ml, mu, maxord, mxstep, neq, reserved;
    unsigned long long nd, ni;
};
static const int kSavMagic = 0x62766f65;  // "bvoe": layout 2 (layout 1 had no istate / t / y / maxord / mxstep)
static size_t sav_bytes(const bvode_solver *s, size_t nd, size_t ni) {
    const size_t ns = (size_t)s->nsys;
    return sizeof(SavHeader) + sizeof(real) * nd + sizeof(int) * ni + sizeof(real) * 4 * ns +
           sizeof(int) * 12 * ns + sizeof(int) * ns + sizeof(real) * ns + sizeof(real) * ns * s->neq;
}
long long bvode_state_bytes(const bvode_solver *s) {
    if (!s || !s->d_state) return 0;
    return (long long)sav_bytes(s, s->nd, s->ni);
}
int bvode_dvsrco_n(bvode_solver *s, void *sav, long long nbytes, int job) {
    if (!s || !sav || (job != 1 && job != 2)) return BVODE_EINVAL;  // M:3212-3213 `error stop`
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->stream2) CUDA_TRY(cudaStreamSynchronize(s->stream2));
    char *p = (char *)sav;
    if (job == 1) {
        if (!s->d_state) { g_err = "dvsrco(job = 1) before any successful solve"; return BVODE_EINVAL; }
        if (nbytes >= 0 && (size_t)nbytes < sav_bytes(s, s->nd, s->ni)) { g_err = "dvsrco: buffer shorter than bvode_state_bytes()"; return BVODE_EINVAL; }
        SavHeader h{kSavMagic + ((s->flags & BVODE_FLAG_STRICT) ? 1 : 0), s->problem, s->nsys, s->mf, s->ml, s->mu, s->maxord, s->mxstep, s->neq, s->nact,
                    (unsigned long long)s->nd, (unsigned long long)s->ni};
        memcpy(p, &h, sizeof(h));
    } else {
        if (nbytes >= 0 && (size_t)nbytes < sizeof(SavHeader)) { g_err = "dvsrco: buffer shorter than its header"; return BVODE_EINVAL; }
        SavHeader h;
        memcpy(&h, p, sizeof(h));
        // the two builds keep different things in the saved matrix image (LINPACK's factors in the strict build; the
        // pivoted inverse / the unfactored band matrix in the fast one), so a state goes back into the build it came from
        if (h.magic == kSavMagic + ((s->flags & BVODE_FLAG_STRICT) ? 0 : 1)) {
            g_err = "dvsrco(job = 2): the buffer was saved by the other build (BVODE_FLAG_STRICT differs)";
            return BVODE_EINVAL;
        }
        if (h.magic != kSavMagic + ((s->flags & BVODE_FLAG_STRICT) ? 1 : 0) || h.problem != s->problem || h.nsys != s->nsys || h.neq != s->neq) {
            g_err = "dvsrco(job = 2): the buffer was not saved from a handle of this problem and batch size";
            return BVODE_EINVAL;
        }
        size_t nd = 0, ni = 0;
        if (!s->pe->supports(h.mf, h.ml, h.mu)) return BVODE_EINVAL;
        s->pe->state_size(h.mf, h.ml, h.mu, (size_t)s->nsys, &nd, &ni);
        if (h.nd != nd || h.ni != ni) { g_err = "dvsrco(job = 2): state sizes in the buffer do not match this build"; return BVODE_EINVAL; }
        if (nbytes >= 0 && (size_t)nbytes < sav_bytes(s, nd, ni)) { g_err = "dvsrco(job = 2): buffer shorter than the state it describes"; return BVODE_EINVAL; }
        int rc = ensure_state(s, h.mf, h.ml, h.mu);
        if (rc) return rc;
        s->maxord = h.maxord;
        s->mxstep = h.mxstep;
        s->nact = h.reserved;  // neq of the last call when an istate = 3 call had reduced it
    }
    p += sizeof(SavHeader);
    const size_t ns = (size_t)s->nsys;
    struct { void *d; size_t n; } parts[7] = {
        {s->d_state, sizeof(real) * s->nd}, {s->i_state, sizeof(int) * s->ni},
        {s->d_rout, sizeof(real) * 4 * ns}, {s->d_iout, sizeof(int) * 12 * ns},
        {s->d_istate, sizeof(int) * ns}, {s->d_t, sizeof(real) * ns}, {s->d_y, sizeof(real) * ns * s->neq}};
    for (auto &q : parts) {
        if (job == 1) CUDA_TRY(cudaMemcpy(p, q.d, q.n, cudaMemcpyDeviceToHost));
        else CUDA_TRY(cudaMemcpy(q.d, p, q.n, cudaMemcpyHostToDevice));
        p += q.n;
    }
    return BVODE_OK;
}
int bvode_dvsrco(bvode_solver *s, void *sav, int job) { return bvode_dvsrco_n(s, sav, -1, job); }

// ---- xerrwd (M:3351-3394) as a host-side message stream --------------------------------
// The kernels never print.  The reference's messages for the per-system outcomes are rebuilt
// here from istate and the optional outputs (tcur, hcur, tolsf, imxer) of the last solve.
int bvode_istate_summary(bvode_solver *s, int counts[8]) {
    if (!s || !counts) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    std::vector<int> ist((size_t)s->nsys);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(ist.data(), s->d_istate, sizeof(int) * ist.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; k++) counts[k] = 0;
    for (int v : ist) {
        if (v >= 1 && v <= 3) counts[0]++;        // 1, 2 (3 never comes back): no error
        else if (v <= -1 && v >= -6) counts[-v]++;
        else counts[7]++;
    }
    return BVODE_OK;
}

// REALIFY-OFF
// Fortran D21.13 edit descriptor: 0.dddddddddddddD+ee
static std::string fmt_d21_13(double x) {
    char buf[64];
    if (x == 0.0) return "   0.0000000000000D+00";
    snprintf(buf, sizeof buf, "%.12E", fabs(x));
    std::string m(buf);
    const size_t e = m.find('E');
    int ex = atoi(m.c_str() + e + 1) + 1;
    std::string digits = m.substr(0, 1) + m.substr(2, e - 2);
    snprintf(buf, sizeof buf, "%s0.%sD%c%02d", x < 0 ? "-" : "", digits.c_str(), ex < 0 ? '-' : '+', abs(ex));
    std::string out(buf);
    while (out.size() < 21) out = " " + out;
    return out;
}

// REALIFY-ON
int bvode_format_message(bvode_solver *s, int sys, char *buf, int buflen) {
    if (!s || !buf || buflen < 1 || sys < 0 || sys >= s->nsys) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const size_t ns = (size_t)s->nsys;
    int ist = 0, imxer = 0;
    real r[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpy(&ist, s->d_istate + sys, sizeof(int), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; k++)
        CUDA_TRY(cudaMemcpy(&r[k], s->d_rout + k * ns + sys, sizeof(real), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&imxer, s->d_iout + 5 * ns + sys, sizeof(int), cudaMemcpyDeviceToHost));
    const real hcur = r[1], tcur = r[2], tolsf = r[3];
    auto i1 = [](int v) { char b[64]; snprintf(b, sizeof b, "      in above message,  i1 =%10d\n", v); return std::string(b); };
    auto r1 = [](real a) { return "      in above message,  r1 =" + fmt_d21_13(a) + "\n"; };
    auto r12 = [](real a, real b) { return "      in above message,  r1 =" + fmt_d21_13(a) + "   r2 =" + fmt_d21_13(b) + "\n"; };
    std::string m;
    switch (ist) {
    case -1:  // M:1506-1509
        m = "dvode--  at current t (=r1), mxstep (=i1) steps\n      taken on this call before reaching tout\n" +
            i1(s->mxstep) + r1(tcur);
        break;
    case -2:  // M:1637-1640
        m = "dvode--  at t (=r1), too much accuracy requested\n      for precision of machine:   see tolsf (=r2)\n" +
            r12(tcur, tolsf);
        break;
    case -3:  // M:1081-1461: the device only sees messages 21, 22, 23-25, 26, 27
        if (tolsf != BV_R(0.0))
            m = "dvode--  at start of problem, too much accuracy\n      requested for precision of machine:   see tolsf (=r1)\n" + r1(tolsf);
        else
            m = "dvode--  illegal input detected on the device (ewt <= 0, tout too close to t to start,\n"
                "      or itask / tout / tcrit inconsistent with the current t: messages 21-25, 27)\n";
        break;
    case -4:  // M:1560-1563
        m = "dvode--  at t(=r1) and step size h(=r2), the error\n      test failed repeatedly or with abs(h) = hmin\n" +
            r12(tcur, hcur);
        break;
    case -5:  // M:1570-1575
        m = "dvode--  at t (=r1) and step size h (=r2), the\n      corrector convergence failed repeatedly\n"
            "      or with abs(h) = hmin\n" + r12(tcur, hcur);
        break;
    case -6:  // M:1689-1690 (the index and value of the offending weight are not kept)
        m = "dvode--  at t (=r1), ewt(i1) has become r2 <= 0.\n" + r1(tcur);
        break;
    default:
        m = "";
    }
    if ((ist == -4 || ist == -5) && imxer > 0) {
        char b[96];
        snprintf(b, sizeof b, "      largest weighted local error: component imxer =%6d\n", imxer);
        m += b;
    }
    snprintf(buf, (size_t)buflen, "%s", m.c_str());
    return (int)m.size();
}

// ---- the device LINPACK routines in isolation (csrc/bvode_lutest.cu) ----------------------------
#if !BVODE_REAL_IS_FLOAT
extern "C" int bvode_lutest_fast(int, int, int, int, int, const real *, const real *, real *, int *, real *, int *);
extern "C" int bvode_lutest_strict(int, int, int, int, int, const real *, const real *, real *, int *, real *, int *);
#endif
int bvode_lu_test(int kind, int n, int ml, int mu, int nmat, const real *a, const real *b, real *afac,
                  int *ipvt, real *x, int *info, int device, unsigned flags) {
    if (!a || !b || !afac || !ipvt || !x || !info || nmat < 1 || n < 1) return BVODE_EINVAL;
#if BVODE_REAL_IS_FLOAT
    g_err = "bvode_lu_test: not part of the REAL32 build";
    return BVODE_ENOTIMPL;
#else
    CUDA_TRY(cudaSetDevice(device));
    const size_t na = (size_t)nmat * (kind == 3 ? (size_t)(2 * ml + mu + 1) * n : (size_t)n * n), nb = (size_t)nmat * n;
    real *da = nullptr, *db = nullptr, *df = nullptr, *dx = nullptr;
    int *dp = nullptr, *di = nullptr;
    cudaError_t e = cudaMalloc(&da, sizeof(real) * na);
    if (e == cudaSuccess) e = cudaMalloc(&df, sizeof(real) * na);
    if (e == cudaSuccess) e = cudaMalloc(&db, sizeof(real) * nb);
    if (e == cudaSuccess) e = cudaMalloc(&dx, sizeof(real) * nb);
    if (e == cudaSuccess) e = cudaMalloc(&dp, sizeof(int) * nb);
    if (e == cudaSuccess) e = cudaMalloc(&di, sizeof(int) * nmat);
    if (e == cudaSuccess) e = cudaMemcpy(da, a, sizeof(real) * na, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, sizeof(real) * nb, cudaMemcpyHostToDevice);
    int rc = BVODE_OK;
    if (e == cudaSuccess) {
        const int r = (flags & BVODE_FLAG_STRICT) ? bvode_lutest_strict(kind, n, ml, mu, nmat, da, db, df, dp, dx, di)
                                                  : bvode_lutest_fast(kind, n, ml, mu, nmat, da, db, df, dp, dx, di);
        if (r == -1000) { g_err = "bvode_lu_test: this kind / n is not instantiated"; rc = BVODE_ENOTIMPL; }
        else if (r != 0) e = (cudaError_t)r;
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && rc == BVODE_OK) {
        e = cudaMemcpy(afac, df, sizeof(real) * na, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(x, dx, sizeof(real) * nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(ipvt, dp, sizeof(int) * nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(info, di, sizeof(int) * nmat, cudaMemcpyDeviceToHost);
    }
    cudaFree(da); cudaFree(db); cudaFree(df); cudaFree(dx); cudaFree(dp); cudaFree(di);
    if (e != cudaSuccess) { g_err = std::string("bvode_lu_test: ") + cudaGetErrorString(e); return BVODE_ECUDA; }
    return rc;
#endif
}

// REALIFY-OFF
int bvode_measure_fp64_peak(int device, double *tflops) {
    if (!tflops) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, tpb = 256, iters = 1 << 16;
    double *out;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * blocks * tpb));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    k_dfma_peak<<<blocks, tpb>>>(out, 1 << 12);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, tpb>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * tpb;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return BVODE_OK;
}
