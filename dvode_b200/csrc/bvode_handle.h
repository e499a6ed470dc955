// bvode_handle.h -- the solver handle behind `bvode_solver*` (private to the C-ABI layer:
// bvode_capi.cu and bvode_multi.cu).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "bvode_internal.h"

struct bvode_solver {
    int problem = 0, nsys = 0, device = 0, neq = 0, npar = 0;
    unsigned flags = 0;
    const ProblemEntry *pe = nullptr;
    cudaStream_t stream = nullptr;
    // state (allocated at the first istate = 1 call, when mf / ml / mu are known)
    int mf = 0, ml = 0, mu = 0;
    int maxord = 0;  // of the last call (istate = 3 may change it)
    int mxstep = 500;
    int nact = 0;    // equations actually integrated after an istate = 3 call with a smaller neq (0 = all neq)
    size_t nd = 0, ni = 0;  // doubles / ints of persistent state for the whole batch
    real *d_state = nullptr;
    int *i_state = nullptr;
    // batch buffers, SoA on the device
    real *d_params = nullptr, *d_y = nullptr, *d_t = nullptr, *d_tol = nullptr,
           *d_rout = nullptr, *d_aos = nullptr, *d_yout = nullptr, *d_tout_ps = nullptr;
    int *d_istate = nullptr, *d_iout = nullptr, *d_iaos = nullptr;
    int *d_next = nullptr;  // hands out the systems beyond the grid (thread-per-system kernels); one per stream
    cudaStream_t stream2 = nullptr;  // second stream of the pipelined host-buffer path
    cudaStream_t stream3 = nullptr;  // third
    real *d_yout_aos = nullptr;    // [nout][nsys][neq] staging of the output points in host layout
    size_t yout_aos_cap = 0;
    size_t aos_cap = 0, yout_cap = 0;
    bool params_external = false;
    int last_launches = 0;
    // ---- per-system tolerances / optional inputs (BVODE_FLAG_PER_SYSTEM)
    real *d_tol_ps = nullptr, *d_optd_ps = nullptr;
    int *d_opti_ps = nullptr;
    std::vector<int> maxord_ps;      // per-system maxord of the last call (istate = 3 compares with it)
    // ---- optional outputs of the host-buffer path: [nsys][4] / [nsys][12] on the device, then pinned
    real *d_rout_aos = nullptr, *h_rout = nullptr;
    int *h_iout = nullptr;
    std::vector<cudaEvent_t> chunk_done;
    // ---- what the last solve call left where (bvode_gather reads these)
    real *last_y = nullptr, *last_t = nullptr;  // SoA [neq][nsys], [nsys] on the device
    int *last_istate = nullptr;
    cudaStream_t last_stream = nullptr;  // the stream the last solve was enqueued on
    real *d_pack = nullptr;        // bvode_gather: this rank's packed final results
    size_t pack_cap = 0;
};
