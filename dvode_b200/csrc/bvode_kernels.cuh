// bvode_kernels.cuh -- kernel entry points, state load/store and the per-build registry.
// Included by bvode_kernels_fast.cu and bvode_kernels_strict.cu (same source, two builds).
#pragma once
#include "bvode_internal.h"
#include "policy_thread.cuh"
#include "problems.cuh"

namespace BVODE_NS {

// ------------------------------------------------------------ state layout ------
// The reference keeps the solver state in the caller's rwork(21:) / iwork(31:) and in
// dvode_data_t.  Here it is structure-of-arrays in HBM: slot k of system s is at
// dstate[k * nsys + s], so a warp's 32 loads of one slot are one coalesced 256-byte line.
// visit_state enumerates the persistent fields in a fixed order; Loader / Storer / Counter
// are the three visitors.
template <class ENG, class V>
BV_DEV void visit_state(ENG &e, V &v) {
    Ctl &c = e.c;
#pragma unroll
    for (int j = 0; j < kLmax; j++)
#pragma unroll
        for (int i = 0; i < ENG::NLOC; i++) v.d(e.yh[j][i]);
    ENG::POLICY::visit_acsav(e.acsav, v);
#pragma unroll
    for (int i = 0; i < ENG::NLOC; i++) v.d(e.acor[i]);
    ENG::POLICY::visit_lin(e.lin, v);
#pragma unroll
    for (int j = 1; j <= kLmax; j++) v.d(c.tau[j]);
    v.d(c.h); v.d(c.hscal); v.d(c.hnew); v.d(c.hu); v.d(c.eta); v.d(c.etamax); v.d(c.tn);
    v.d(c.rc); v.d(c.prl1); v.d(c.crate); v.d(c.conp);
    v.i(c.nq); v.i(c.nqwait); v.i(c.newq); v.i(c.newh); v.i(c.jstart); v.i(c.icf); v.i(c.ipup);
    v.i(c.nslj); v.i(c.nslp); v.i(c.nhnil); v.i(c.init);
    v.i(c.nst); v.i(c.nfe); v.i(c.nje); v.i(c.nlu); v.i(c.nni); v.i(c.ncfn); v.i(c.netf);
    v.i(c.nqu);
}

struct Loader {
    const double *dp;
    const int *ip;
    size_t stride;
    int kd = 0, ki = 0;
    BV_DEV void d(double &x) { x = dp[(size_t)kd * stride]; kd++; }
    BV_DEV void i(int &x) { x = ip[(size_t)ki * stride]; ki++; }
    BV_DEV void skip_d(int n) { kd += n; }
};
struct Storer {
    double *dp;
    int *ip;
    size_t stride;
    int kd = 0, ki = 0;
    BV_DEV void d(double &x) { dp[(size_t)kd * stride] = x; kd++; }
    BV_DEV void i(int &x) { ip[(size_t)ki * stride] = x; ki++; }
    BV_DEV void skip_d(int n) { kd += n; }
};

// ------------------------------------------------------- thread-per-system ------
constexpr int kThreadTPB = 128;
#ifndef BVODE_MINBLOCKS
#define BVODE_MINBLOCKS 2
#endif

template <class POL>
__global__ void __launch_bounds__(kThreadTPB, BVODE_MINBLOCKS) vode_thread_kernel(const __grid_constant__ KArgs a) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= a.nsys) return;
    int ist = a.istate[sys];
    if (ist != 1 && ist != 2) return;  // a failed / illegal system is left untouched
    constexpr int N = POL::N;
    using Prob = typename POL::Prob;

    Engine<POL> e;
    e.o = a.o;
    e.tol.r = a.rtolv;
    e.tol.a = a.atolv;
    e.acsav.p = a.dstate + (size_t)POL::kSlotAcsav * a.nsys + sys;
    e.acsav.stride = (size_t)a.nsys;
    e.lin.wjs.p = a.dstate + (size_t)POL::kSlotWjs * a.nsys + sys;
    e.lin.wjs.stride = (size_t)a.nsys;
#pragma unroll
    for (int k = 0; k < Prob::NPAR; k++) e.par[k] = a.params[(size_t)k * a.nsys + sys];

    double t = a.t[sys];
    double tolsf = 0.0;
    bool first = false;
    const double tout0 = a.tout_ps ? a.tout_ps[sys] : a.touts[0];
    if (ist == 1) {
        if (tout0 == t) return;  // M:1101
#pragma unroll
        for (int i = 0; i < N; i++) e.y[i] = a.y[(size_t)i * a.nsys + sys];
        const int r = e.init_problem(t, tout0);
        if (r < 0) {  // illegal input detected on the device (messages 21, 22): y, t untouched
            a.istate[sys] = r;
            return;
        }
        first = true;
    } else {
        Loader ld{a.dstate + sys, a.istate_store + sys, (size_t)a.nsys};
        visit_state(e, ld);
        e.c.L = e.c.nq + 1;
        e.c.kflag = 0;
        e.c.jcur = 0;
        if (e.c.init != 1) {  // M:1102-1106
            a.istate[sys] = IST_ILLEGAL;
            return;
        }
    }

    int nemit = 0;
    ist = e.run(
        a.nout, [&](int k) { return (k == 0) ? tout0 : a.touts[k]; }, first, t,
        [&](int k, const double (&yk)[N]) {
            nemit = k + 1;
            if (a.y_out) {
#pragma unroll
                for (int i = 0; i < N; i++) a.y_out[((size_t)k * N + i) * a.nsys + sys] = yk[i];
            }
        });
    if (a.y_out) {  // a system that failed keeps its last state in the remaining output slots
        for (int k2 = nemit; k2 < a.nout; k2++) {
#pragma unroll
            for (int i = 0; i < N; i++) a.y_out[((size_t)k2 * N + i) * a.nsys + sys] = e.y[i];
        }
    }
    if (ist == IST_TOLSF || (ist == IST_ILLEGAL && e.c.nst == 0))
        tolsf = 2.0 * kUround * POL::norm(e.yh[0], e.ewt);  // M:1626, 1632, 1641

#pragma unroll
    for (int i = 0; i < N; i++) a.y[(size_t)i * a.nsys + sys] = e.y[i];
    a.t[sys] = t;
    a.istate[sys] = ist;
    {
        Storer st{a.dstate + sys, a.istate_store + sys, (size_t)a.nsys};
        visit_state(e, st);
    }
    // optional outputs, M:1670-1684 (success) / M:1708-1723 (failure)
    const size_t ns = (size_t)a.nsys;
    a.rout[0 * ns + sys] = e.c.hu;
    a.rout[1 * ns + sys] = (ist == IST_OK) ? e.c.hnew : e.c.h;
    a.rout[2 * ns + sys] = e.c.tn;
    a.rout[3 * ns + sys] = tolsf;
    a.iout[0 * ns + sys] = e.c.nst;
    a.iout[1 * ns + sys] = e.c.nfe;
    a.iout[2 * ns + sys] = e.c.nje;
    a.iout[3 * ns + sys] = e.c.nqu;
    a.iout[4 * ns + sys] = (ist == IST_OK) ? e.c.newq : e.c.nq;
    a.iout[5 * ns + sys] =
        (ist == IST_ERRTEST || ist == IST_CONVFAIL) ? POL::imxer(e.acor, e.ewt) : 0;
    a.iout[6 * ns + sys] = a.lenrw;
    a.iout[7 * ns + sys] = a.leniw;
    a.iout[8 * ns + sys] = e.c.nlu;
    a.iout[9 * ns + sys] = e.c.nni;
    a.iout[10 * ns + sys] = e.c.ncfn;
    a.iout[11 * ns + sys] = e.c.netf;
}

// dvindy for any k (M:1878-1959) from the stored state; one thread per system.
template <class POL>
__global__ void __launch_bounds__(kThreadTPB) dvindy_thread_kernel(const __grid_constant__ DkyArgs a) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= a.nsys) return;
    constexpr int N = POL::N;
    Engine<POL> e;
    e.acsav.p = nullptr;
    e.lin.wjs.p = nullptr;
    Loader ld{a.dstate + sys, a.istate_store + sys, (size_t)a.nsys};
    visit_state(e, ld);
    const Ctl &c = e.c;
    const int nq = c.nq, L = nq + 1, k = a.k;
    const double t = a.t_ps ? a.t_ps[sys] : a.t;
    int iflag = 0;
    double dky[N];
#pragma unroll
    for (int i = 0; i < N; i++) dky[i] = 0.0;
    if (k < 0 || k > nq) {
        iflag = -1;
    } else {
        const double tfuzz = 100.0 * kUround * dsign(fabs(c.tn) + fabs(c.hu), c.hu);
        const double tp = c.tn - c.hu - tfuzz;
        const double tn1 = c.tn + tfuzz;
        if ((t - tp) * (t - tn1) > 0.0) {
            iflag = -2;
        } else {
            const double s = (t - c.tn) / c.h;
            int ic = 1;
            if (k != 0)
                for (int jj = L - k; jj <= nq; jj++) ic = ic * jj;
            double cc = (double)ic;
            double col[N];
            e.get_col(L, col);
#pragma unroll
            for (int i = 0; i < N; i++) dky[i] = cc * col[i];
            if (k != nq) {
                for (int jb = 1; jb <= nq - k; jb++) {
                    const int j = nq - jb, jp1 = j + 1;
                    ic = 1;
                    if (k != 0)
                        for (int jj = jp1 - k; jj <= j; jj++) ic = ic * jj;
                    cc = (double)ic;
                    e.get_col(jp1, col);
#pragma unroll
                    for (int i = 0; i < N; i++) dky[i] = cc * col[i] + s * dky[i];
                }
            }
            if (k != 0) {
                const double r = powi(c.h, -k);
#pragma unroll
                for (int i = 0; i < N; i++) dky[i] = r * dky[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) a.dky[(size_t)i * a.nsys + sys] = dky[i];
    a.iflag[sys] = iflag;
}

// --------------------------------------------------------------- launchers ------
struct CountV {
    int kd = 0, ki = 0;
    void d(double &) { kd++; }
    void i(int &) { ki++; }
};

template <class POL>
static int launch_thread(const KArgs &a, cudaStream_t st) {
    const int grid = (a.nsys + kThreadTPB - 1) / kThreadTPB;
    vode_thread_kernel<POL><<<grid, kThreadTPB, 0, st>>>(a);
    return (int)cudaGetLastError();
}
template <class POL>
static int launch_dky_thread(const DkyArgs &a, cudaStream_t st) {
    const int grid = (a.nsys + kThreadTPB - 1) / kThreadTPB;
    dvindy_thread_kernel<POL><<<grid, kThreadTPB, 0, st>>>(a);
    return (int)cudaGetLastError();
}

// thread-per-system problems support mf = +-21, +-22 (dense chord iteration, BDF)
template <class PROB>
static int thread_dispatch(int mf, const KArgs &a, cudaStream_t st) {
    switch (mf) {
    case 21: return launch_thread<ThreadPolicy<PROB, 1, true>>(a, st);
    case 22: return launch_thread<ThreadPolicy<PROB, 2, true>>(a, st);
    case -21: return launch_thread<ThreadPolicy<PROB, 1, false>>(a, st);
    case -22: return launch_thread<ThreadPolicy<PROB, 2, false>>(a, st);
    default: return -1000;
    }
}
template <class PROB>
static int thread_dky_dispatch(int mf, const DkyArgs &a, cudaStream_t st) {
    // the state layout depends on SAVEJ only through the Lin block; MITER does not matter
    if (mf > 0) return launch_dky_thread<ThreadPolicy<PROB, 1, true>>(a, st);
    return launch_dky_thread<ThreadPolicy<PROB, 1, false>>(a, st);
}
template <class PROB>
static void thread_state_size(int mf, int *nd, int *ni) {
    // dense: N*N for P plus N*N for the saved J when mf > 0
    constexpr int N = PROB::N;
    *nd = kLmax * N + 2 * N + N * N + (mf > 0 ? N * N : 0) + kLmax + 11;
    *ni = N + 19;
}
template <class PROB>
static int thread_supports(int mf, int ml, int mu) {
    (void)ml; (void)mu;
    return mf == 21 || mf == 22 || mf == -21 || mf == -22;
}

#define BVODE_THREAD_PROBLEM(NAME, PROB)                                                    \
    ProblemEntry {                                                                          \
        NAME, PROB::N, PROB::NPAR, 0, thread_supports<PROB>, thread_state_size<PROB>,       \
            thread_dispatch<PROB>, thread_dky_dispatch<PROB>                                \
    }

}  // namespace BVODE_NS
