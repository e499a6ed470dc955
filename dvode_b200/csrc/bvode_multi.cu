// bvode_multi.cu -- the multi-GPU part of the C ABI: index sharding of a batch over the
// GPUs of one box and the ONE exchange the path has, the final gather of solutions and statistics
// to a root GPU over NVLink.
//
//   * bvode_shard_range            contiguous blocks [first, last) per rank, sizes differ by at most one
//   * bvode_comm_*                 an NCCL communicator behind plain C types: created from a unique id
//                                  (one process per GPU, the id travels through whatever the application
//                                  uses: MPI_Bcast, a file, torch.distributed's store) or adopted from
//                                  an ncclComm_t the application already has
//   * bvode_gather[_device]        every rank packs the final y / t / istate / rwork(11:14) / iwork(11:22)
//                                  of its handle in host layout (one kernel) and sends the five sections
//                                  straight into their place in the root's arrays (one NCCL group of
//                                  send / recv; nothing is staged or concatenated on the root)
//   * bvode_multi_*                one process, several GPUs: a handle per device, the host-buffer solve
//                                  of every slice on its own thread -- the call a Fortran / C program with no
//                                  MPI makes to use the whole box.  Results land in the caller's host arrays
//                                  through each GPU's own PCIe link; no inter-GPU traffic is needed for that.
//
// NCCL is loaded at run time (dlopen) so that libbvode.so has no link-time dependency on it: the
// single-GPU entry points work on a machine without NCCL, and inside a process that already loaded
// an NCCL (PyTorch's) the same copy is used.
#include "../../include/bvode.h"
#include "bvode_handle.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern void bvode_set_error(const std::string &msg);  // bvode_capi.cu (thread-local text of bvode_last_error)

// ---------------------------------------------------------------- NCCL, loaded at run time ----
// The handful of declarations used here, as in nccl.h 2.x (stable ABI).
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0 };
struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommInitAll)(ncclComm_t *, int, const int *);
    int (*CommDestroy)(ncclComm_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(int);
    int (*GetVersion)(int *);
    bool ok = false;
    std::string why;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static const NcclApi &nccl() {
    std::call_once(g_nccl_once, [] {
        void *h = nullptr;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { g_nccl.why = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed"); return; }
#define BV_SYM(field, sym)                                                              \
        *(void **)(&g_nccl.field) = dlsym(h, sym);                                     \
        if (!g_nccl.field) { g_nccl.why = std::string("NCCL symbol missing: ") + sym; return; }
        BV_SYM(GetUniqueId, "ncclGetUniqueId")
        BV_SYM(CommInitRank, "ncclCommInitRank")
        BV_SYM(CommInitAll, "ncclCommInitAll")
        BV_SYM(CommDestroy, "ncclCommDestroy")
        BV_SYM(GroupStart, "ncclGroupStart")
        BV_SYM(GroupEnd, "ncclGroupEnd")
        BV_SYM(Send, "ncclSend")
        BV_SYM(Recv, "ncclRecv")
        BV_SYM(GetErrorString, "ncclGetErrorString")
        BV_SYM(GetVersion, "ncclGetVersion")
#undef BV_SYM
        g_nccl.ok = true;
    });
    return g_nccl;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            bvode_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
            return BVODE_ECUDA;                                                            \
        }                                                                                  \
    } while (0)
#define NCCL_TRY(expr)                                                                     \
    do {                                                                                   \
        int _e = (expr);                                                                   \
        if (_e != ncclSuccess) {                                                           \
            bvode_set_error(std::string(#expr) + ": " + nccl().GetErrorString(_e));        \
            return BVODE_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

struct bvode_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    bool owned = false;
};

void bvode_shard_range(long long nsys_total, int rank, int world, long long *first, long long *last) {
    if (world < 1) world = 1;
    const long long base = nsys_total / world, rem = nsys_total % world;
    const long long f = rank * base + (rank < rem ? rank : rem);
    if (first) *first = f;
    if (last) *last = f + base + (rank < rem ? 1 : 0);
}

int bvode_nccl_version(void) {
    if (!nccl().ok) return 0;
    int v = 0;
    nccl().GetVersion(&v);
    return v;
}

int bvode_comm_unique_id(char id[128]) {
    if (!id) return BVODE_EINVAL;
    if (!nccl().ok) { bvode_set_error(nccl().why); return BVODE_ENOTIMPL; }
    ncclUniqueId u;
    NCCL_TRY(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return BVODE_OK;
}

int bvode_comm_create(bvode_comm **out, const char id[128], int world, int rank, int device) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return BVODE_EINVAL;
    bvode_comm *c = new bvode_comm();
    c->world = world;
    c->rank = rank;
    c->device = device;
    if (world > 1) {
        if (!nccl().ok) { bvode_set_error(nccl().why); delete c; return BVODE_ENOTIMPL; }
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) { bvode_set_error(cudaGetErrorString(e)); delete c; return BVODE_ECUDA; }
        ncclUniqueId u;
        memcpy(u.internal, id, 128);
        const int r = nccl().CommInitRank(&c->comm, world, u, rank);
        if (r != ncclSuccess) { bvode_set_error(std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); delete c; return BVODE_ECUDA; }
        c->owned = true;
    }
    *out = c;
    return BVODE_OK;
}

int bvode_comm_adopt(bvode_comm **out, void *nccl_comm, int world, int rank, int device) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm)) return BVODE_EINVAL;
    if (world > 1 && !nccl().ok) { bvode_set_error(nccl().why); return BVODE_ENOTIMPL; }
    bvode_comm *c = new bvode_comm();
    c->comm = (ncclComm_t)nccl_comm;
    c->world = world;
    c->rank = rank;
    c->device = device;
    c->owned = false;
    *out = c;
    return BVODE_OK;
}

void bvode_comm_destroy(bvode_comm *c) {
    if (!c) return;
    if (c->owned && c->comm && nccl().ok) {
        cudaSetDevice(c->device);
        nccl().CommDestroy(c->comm);
    }
    delete c;
}

// ------------------------------------------------------------------------------ the gather ----
// Sections of one rank's packed results, host layout: y[n][neq] | t[n] | rout[n][4] | iout[n][12] | istate[n]
struct PackLayout {
    size_t off_y, off_t, off_rout, off_iout, off_ist, total;
};
static PackLayout pack_layout(size_t n, int neq) {
    PackLayout L;
    L.off_y = 0;
    L.off_t = L.off_y + sizeof(real) * n * neq;
    L.off_rout = L.off_t + sizeof(real) * n;
    L.off_iout = L.off_rout + sizeof(real) * 4 * n;
    L.off_ist = L.off_iout + sizeof(int) * 12 * n;
    L.total = L.off_ist + sizeof(int) * n;
    return L;
}
// SoA (the kernels' layout) -> the five host-layout sections; any destination may be null
__global__ void k_pack_results(int n, int neq, const real *__restrict__ y, const real *__restrict__ t,
                               const int *__restrict__ ist, const real *__restrict__ rout,
                               const int *__restrict__ iout, real *__restrict__ py, real *__restrict__ pt,
                               real *__restrict__ prout, int *__restrict__ piout, int *__restrict__ pist) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t ns = (size_t)n;
    if (py) for (int i = 0; i < neq; i++) py[(size_t)s * neq + i] = y[(size_t)i * ns + s];
    if (pt) pt[s] = t[s];
    if (prout) for (int k = 0; k < 4; k++) prout[(size_t)s * 4 + k] = rout[(size_t)k * ns + s];
    if (piout) for (int k = 0; k < 12; k++) piout[(size_t)s * 12 + k] = iout[(size_t)k * ns + s];
    if (pist) pist[s] = ist[s];
}

int bvode_gather_device(bvode_solver *s, bvode_comm *c, int root, long long nsys_total, real *y_all,
                        real *t_all, int *istate_all, real *rout_all, int *iout_all, void *stream) {
    if (!s || !c || root < 0 || root >= c->world) return BVODE_EINVAL;
    long long first, last;
    bvode_shard_range(nsys_total, c->rank, c->world, &first, &last);
    if (last - first != s->nsys) {
        bvode_set_error("bvode_gather: this handle's nsys is not this rank's block of nsys_total (bvode_shard_range)");
        return BVODE_EINVAL;
    }
    if (!s->last_y || !s->last_t || !s->last_istate) { bvode_set_error("bvode_gather before any solve"); return BVODE_EINVAL; }
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;  // as in bvode_solve_*_device: 0 = the legacy default stream
    const int n = s->nsys, neq = s->nact > 0 ? s->nact : s->neq;  // host layout: the neq of the last call
    const bool is_root = (c->rank == root);
    const int TB = 256, grid = (n + TB - 1) / TB;
    if (is_root) {
        // the root's own block goes straight into its place in the gathered arrays
        k_pack_results<<<grid, TB, 0, st>>>(n, neq, s->last_y, s->last_t, s->last_istate, s->d_rout, s->d_iout,
                                            y_all ? y_all + (size_t)first * neq : nullptr, t_all ? t_all + first : nullptr,
                                            rout_all ? rout_all + (size_t)first * 4 : nullptr,
                                            iout_all ? iout_all + (size_t)first * 12 : nullptr,
                                            istate_all ? istate_all + first : nullptr);
        CUDA_TRY(cudaGetLastError());
        if (c->world == 1) return BVODE_OK;
        NCCL_TRY(nccl().GroupStart());
        for (int r = 0; r < c->world; r++) {
            if (r == root) continue;
            long long f, l;
            bvode_shard_range(nsys_total, r, c->world, &f, &l);
            const size_t m = (size_t)(l - f);
            if (m == 0) continue;
            if (y_all) NCCL_TRY(nccl().Recv(y_all + (size_t)f * neq, sizeof(real) * m * neq, ncclInt8, r, c->comm, st));
            if (t_all) NCCL_TRY(nccl().Recv(t_all + f, sizeof(real) * m, ncclInt8, r, c->comm, st));
            if (rout_all) NCCL_TRY(nccl().Recv(rout_all + (size_t)f * 4, sizeof(real) * 4 * m, ncclInt8, r, c->comm, st));
            if (iout_all) NCCL_TRY(nccl().Recv(iout_all + (size_t)f * 12, sizeof(int) * 12 * m, ncclInt8, r, c->comm, st));
            if (istate_all) NCCL_TRY(nccl().Recv(istate_all + f, sizeof(int) * m, ncclInt8, r, c->comm, st));
        }
        NCCL_TRY(nccl().GroupEnd());
        return BVODE_OK;
    }
    // a sending rank: which sections the root wants is part of the call's contract -- every rank passes
    // the same set of non-null `*_all` arguments (on non-root ranks any non-null value means "wanted")
    if (n == 0) return BVODE_OK;
    const PackLayout L = pack_layout((size_t)n, neq);
    if (L.total > s->pack_cap) {
        cudaFree(s->d_pack);
        s->d_pack = nullptr;
        s->pack_cap = 0;
        CUDA_TRY(cudaMalloc(&s->d_pack, L.total));
        s->pack_cap = L.total;
    }
    char *pk = (char *)s->d_pack;
    k_pack_results<<<grid, TB, 0, st>>>(n, neq, s->last_y, s->last_t, s->last_istate, s->d_rout, s->d_iout,
                                        y_all ? (real *)(pk + L.off_y) : nullptr, t_all ? (real *)(pk + L.off_t) : nullptr,
                                        rout_all ? (real *)(pk + L.off_rout) : nullptr,
                                        iout_all ? (int *)(pk + L.off_iout) : nullptr,
                                        istate_all ? (int *)(pk + L.off_ist) : nullptr);
    CUDA_TRY(cudaGetLastError());
    NCCL_TRY(nccl().GroupStart());
    if (y_all) NCCL_TRY(nccl().Send(pk + L.off_y, sizeof(real) * (size_t)n * neq, ncclInt8, root, c->comm, st));
    if (t_all) NCCL_TRY(nccl().Send(pk + L.off_t, sizeof(real) * (size_t)n, ncclInt8, root, c->comm, st));
    if (rout_all) NCCL_TRY(nccl().Send(pk + L.off_rout, sizeof(real) * 4 * (size_t)n, ncclInt8, root, c->comm, st));
    if (iout_all) NCCL_TRY(nccl().Send(pk + L.off_iout, sizeof(int) * 12 * (size_t)n, ncclInt8, root, c->comm, st));
    if (istate_all) NCCL_TRY(nccl().Send(pk + L.off_ist, sizeof(int) * (size_t)n, ncclInt8, root, c->comm, st));
    NCCL_TRY(nccl().GroupEnd());
    return BVODE_OK;
}

// Host destinations on the root: the device gather into scratch on the root's GPU, then one copy per array.
int bvode_gather(bvode_solver *s, bvode_comm *c, int root, long long nsys_total, real *y_all, real *t_all,
                 int *istate_all, real *rout_all, int *iout_all) {
    if (!s || !c || root < 0 || root >= c->world) return BVODE_EINVAL;
    CUDA_TRY(cudaSetDevice(s->device));
    const bool is_root = (c->rank == root);
    const int neq = s->nact > 0 ? s->nact : s->neq;
    char *scratch = nullptr;
    const PackLayout L = pack_layout((size_t)nsys_total, neq);
    real *dy = nullptr, *dt = nullptr, *dr = nullptr;
    int *di = nullptr, *dist = nullptr;
    if (is_root) {
        CUDA_TRY(cudaMalloc(&scratch, L.total));
        dy = (real *)(scratch + L.off_y); dt = (real *)(scratch + L.off_t); dr = (real *)(scratch + L.off_rout);
        di = (int *)(scratch + L.off_iout); dist = (int *)(scratch + L.off_ist);
    }
    // non-root ranks pass "wanted" markers (any non-null pointer)
    real *wy = y_all ? (is_root ? dy : (real *)1) : nullptr, *wt = t_all ? (is_root ? dt : (real *)1) : nullptr;
    real *wr = rout_all ? (is_root ? dr : (real *)1) : nullptr;
    int *wi = iout_all ? (is_root ? di : (int *)1) : nullptr, *wist = istate_all ? (is_root ? dist : (int *)1) : nullptr;
    // the results may still be in flight on the stream of a device-resident solve
    if (s->last_stream != s->stream) CUDA_TRY(cudaStreamSynchronize(s->last_stream));
    int rc = bvode_gather_device(s, c, root, nsys_total, wy, wt, wist, wr, wi, s->stream);
    if (rc == BVODE_OK && is_root) {
        const size_t nt = (size_t)nsys_total;
        cudaError_t e = cudaSuccess;
        if (y_all && e == cudaSuccess) e = cudaMemcpyAsync(y_all, dy, sizeof(real) * nt * neq, cudaMemcpyDeviceToHost, s->stream);
        if (t_all && e == cudaSuccess) e = cudaMemcpyAsync(t_all, dt, sizeof(real) * nt, cudaMemcpyDeviceToHost, s->stream);
        if (rout_all && e == cudaSuccess) e = cudaMemcpyAsync(rout_all, dr, sizeof(real) * 4 * nt, cudaMemcpyDeviceToHost, s->stream);
        if (iout_all && e == cudaSuccess) e = cudaMemcpyAsync(iout_all, di, sizeof(int) * 12 * nt, cudaMemcpyDeviceToHost, s->stream);
        if (istate_all && e == cudaSuccess) e = cudaMemcpyAsync(istate_all, dist, sizeof(int) * nt, cudaMemcpyDeviceToHost, s->stream);
        if (e != cudaSuccess) { bvode_set_error(cudaGetErrorString(e)); rc = BVODE_ECUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess && rc == BVODE_OK) { bvode_set_error(cudaGetErrorString(e)); rc = BVODE_ECUDA; }
    cudaFree(scratch);
    return rc;
}

// --------------------------------------------------------- one process, several GPUs ----
struct bvode_multi {
    int ndev = 0, neq = 0, npar = 0;
    long long nsys_total = 0;
    std::vector<bvode_solver *> h;
    std::vector<long long> first;  // first[d] .. first[d + 1]: the block of device slot d
    std::vector<int> devices;
    std::vector<bvode_comm *> comms;  // created on demand by bvode_multi_gather_device (ncclCommInitAll)
};

int bvode_multi_create(bvode_multi **out, int problem, long long nsys_total, int ndev, const int *devices, unsigned flags) {
    if (!out || ndev < 1 || nsys_total < ndev) return BVODE_EINVAL;
    bvode_multi *m = new bvode_multi();
    m->ndev = ndev;
    m->nsys_total = nsys_total;
    m->first.resize(ndev + 1);
    for (int d = 0; d < ndev; d++) {
        long long f, l;
        bvode_shard_range(nsys_total, d, ndev, &f, &l);
        m->first[d] = f;
        m->first[d + 1] = l;
        m->devices.push_back(devices ? devices[d] : d);
    }
    for (int d = 0; d < ndev; d++) {
        bvode_solver *s = nullptr;
        const int rc = bvode_create(&s, problem, (int)(m->first[d + 1] - m->first[d]), m->devices[d], flags);
        if (rc != BVODE_OK) {
            for (bvode_solver *p : m->h) bvode_destroy(p);
            delete m;
            return rc;
        }
        m->h.push_back(s);
    }
    m->neq = m->h[0]->neq;
    m->npar = m->h[0]->npar;
    *out = m;
    return BVODE_OK;
}

void bvode_multi_destroy(bvode_multi *m) {
    if (!m) return;
    for (bvode_comm *c : m->comms) bvode_comm_destroy(c);
    for (bvode_solver *s : m->h) bvode_destroy(s);
    delete m;
}

int bvode_multi_device_count(const bvode_multi *m) { return m ? m->ndev : 0; }
bvode_solver *bvode_multi_handle(bvode_multi *m, int slot) { return (m && slot >= 0 && slot < m->ndev) ? m->h[slot] : nullptr; }

// Run fn(slot) for every device slot on its own host thread; returns the first non-zero code.
template <class F>
static int for_each_device(bvode_multi *m, F fn) {
    std::vector<int> rc(m->ndev, 0);
    std::vector<std::string> msg(m->ndev);
    std::vector<std::thread> th;
    for (int d = 0; d < m->ndev; d++)
        th.emplace_back([&, d] {
            rc[d] = fn(d);
            if (rc[d] != 0) msg[d] = bvode_last_error();  // thread-local on the worker
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < m->ndev; d++)
        if (rc[d] != 0) { bvode_set_error(msg[d]); return rc[d]; }
    return BVODE_OK;
}

int bvode_multi_set_params(bvode_multi *m, const real *params) {
    if (!m || (!params && m->npar > 0)) return BVODE_EINVAL;
    return for_each_device(m, [&](int d) { return bvode_set_params(m->h[d], params + (size_t)m->first[d] * m->npar); });
}

// internal (bvode_capi.cu): bvode_solve_schedule with the caller's y_out being a slab of a larger
// array -- output point k of this handle's systems starts at y_out + k * yout_sys_stride * neq
extern int bvode_solve_schedule_strided(bvode_solver *s, int neq, real *y, real *t, int nout, const real *touts,
                                        const real *tout_ps, int itol, const real *rtol, const real *atol, int itask,
                                        int *istate, int iopt, real *rwork, int lrw, int *iwork, int liw, int mf,
                                        real *y_out, long long yout_sys_stride);

// solve == dvode (M:604-605) for the whole batch over all devices of the handle; arguments as bvode_solve_schedule
// (tout schedule) with arrays of nsys_total rows.  rtol / atol: batch-uniform, or per system with BVODE_FLAG_PER_SYSTEM.
int bvode_multi_solve_schedule(bvode_multi *m, int neq, real *y, real *t, int nout, const real *touts, int itol,
                               const real *rtol, const real *atol, int itask, int *istate, int iopt, real *rwork,
                               int lrw, int *iwork, int liw, int mf, real *y_out) {
    if (!m || !y || !t || !istate || !touts || nout < 1 || !rtol || !atol) return BVODE_EINVAL;
    const bool ps = (m->h[0]->flags & BVODE_FLAG_PER_SYSTEM) != 0;
    const int nr = ps ? ((itol >= 3) ? neq : 1) : 0, na = ps ? ((itol == 2 || itol == 4) ? neq : 1) : 0;
    return for_each_device(m, [&](int d) {
        const size_t f = (size_t)m->first[d];
        // batch-uniform optional inputs are read from row 0 of each slice: give every slice the batch's row 0
        real *rw = rwork ? rwork + f * lrw : nullptr;
        int *iw = iwork ? iwork + f * liw : nullptr;
        if (!ps && d > 0) {
            if (rw) for (int k = 0; k < 10 && k < lrw; k++) rw[k] = rwork[k];
            if (iw) for (int k = 0; k < 10 && k < liw; k++) iw[k] = iwork[k];
        }
        return bvode_solve_schedule_strided(m->h[d], neq, y + f * neq, t + f, nout, touts, nullptr, itol, rtol + f * nr,
                                            atol + f * na, itask, istate + f, iopt, rw, lrw, iw, liw, mf,
                                            y_out ? y_out + f * neq : nullptr, m->nsys_total);
    });
}

int bvode_multi_solve(bvode_multi *m, int neq, real *y, real *t, const real *tout, int tout_per_system, int itol,
                      const real *rtol, const real *atol, int itask, int *istate, int iopt, real *rwork, int lrw,
                      int *iwork, int liw, int mf) {
    if (!m || !y || !t || !istate || !tout || !rtol || !atol) return BVODE_EINVAL;
    const bool ps = (m->h[0]->flags & BVODE_FLAG_PER_SYSTEM) != 0;
    const int nr = ps ? ((itol >= 3) ? neq : 1) : 0, na = ps ? ((itol == 2 || itol == 4) ? neq : 1) : 0;
    return for_each_device(m, [&](int d) {
        const size_t f = (size_t)m->first[d];
        real *rw = rwork ? rwork + f * lrw : nullptr;
        int *iw = iwork ? iwork + f * liw : nullptr;
        if (!ps && d > 0) {
            if (rw) for (int k = 0; k < 10 && k < lrw; k++) rw[k] = rwork[k];
            if (iw) for (int k = 0; k < 10 && k < liw; k++) iw[k] = iwork[k];
        }
        return bvode_solve_schedule_strided(m->h[d], neq, y + f * neq, t + f, 1, tout_per_system ? nullptr : tout,
                                            tout_per_system ? tout + f : nullptr, itol, rtol + f * nr, atol + f * na, itask,
                                            istate + f, iopt, rw, lrw, iw, liw, mf, nullptr, m->nsys_total);
    });
}

// The final gather into ONE GPU's memory (device slot `root_slot`): y_all[nsys_total][neq] etc. are device
// pointers on that GPU.  One NCCL group over the communicators of ncclCommInitAll.
int bvode_multi_gather_device(bvode_multi *m, int root_slot, real *y_all, real *t_all, int *istate_all,
                              real *rout_all, int *iout_all) {
    if (!m || root_slot < 0 || root_slot >= m->ndev) return BVODE_EINVAL;
    if (m->ndev > 1 && m->comms.empty()) {
        if (!nccl().ok) { bvode_set_error(nccl().why); return BVODE_ENOTIMPL; }
        std::vector<ncclComm_t> cs(m->ndev);
        NCCL_TRY(nccl().CommInitAll(cs.data(), m->ndev, m->devices.data()));
        for (int d = 0; d < m->ndev; d++) {
            bvode_comm *c = new bvode_comm();
            c->comm = cs[d]; c->world = m->ndev; c->rank = d; c->device = m->devices[d]; c->owned = true;
            m->comms.push_back(c);
        }
    }
    if (m->ndev == 1) {
        bvode_comm c1;
        return bvode_gather_device(m->h[0], &c1, 0, m->nsys_total, y_all, t_all, istate_all, rout_all, iout_all, m->h[0]->stream);
    }
    // one thread drives all devices: the sends / receives of every rank inside one group
    NCCL_TRY(nccl().GroupStart());
    int rc = BVODE_OK;
    for (int d = 0; d < m->ndev && rc == BVODE_OK; d++)
        rc = bvode_gather_device(m->h[d], m->comms[d], root_slot, m->nsys_total, y_all, t_all, istate_all, rout_all,
                                 iout_all, m->h[d]->stream);
    NCCL_TRY(nccl().GroupEnd());
    if (rc != BVODE_OK) return rc;
    for (int d = 0; d < m->ndev; d++) {
        CUDA_TRY(cudaSetDevice(m->devices[d]));
        CUDA_TRY(cudaStreamSynchronize(m->h[d]->stream));
    }
    return BVODE_OK;
}
