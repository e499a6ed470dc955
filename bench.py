#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched VODE hot path (BASELINE.json configs[1]).

Workload: 2^20 Robertson systems per GPU with perturbed rate constants, mf = 21 (dense
analytic Jacobian), integrated from t = 0 through the reference's 12-decade output schedule
to t = 4e10 (test/example.f90:56-66), headline tolerances rtol = 1e-6,
atol = (1e-10, 1e-16, 1e-8).  One "step" = one pass of that whole integration over the batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).
  value   systems/s with inputs resident in HBM (CUDA events around the solve launch(es) and, at
          N > 1, the final NCCL gather of solutions and statistics to rank 0: bvode_gather_device)
  e2e     the same through the host-buffer C-ABI call a reference user makes (bvode_solve_schedule
          with rwork / iwork headers): H2D of y0 / t / istate / params, D2H of every output point
          and of the optional outputs; at N > 1 plus the gather to rank 0 and its D2H there
  roofline / other_configs[*].roofline   algorithmic FP64 flops / kernel time vs the
          DFMA peak measured in this run
  strong_scaling   2^20 and 2^24 systems in TOTAL over the N GPUs (device-resident, with the gather)
  strict  the bit-exact build's throughput;  parity  the fast build against the oracle sample
`--impl reference` times the CPU restatement of the reference (oracle/, "port": the image has
no Fortran compiler, so the reference itself cannot be built) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "stiff ODE systems solved/s (Robertson, rtol 1e-6)"
UNIT = "systems/s"
NSYS_PER_GPU = 1 << 20
# measured DRAM traffic of each kernel: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture, per system (profiles/README.md names the capture)
NCU_DRAM_BYTES_PER_SYSTEM = {"robertson": 1137, "advection": None, "mech32": None}
NCU_SOURCE = {"robertson": "profiles/r01_ncu_thread_kernel_r1p.txt (149.0 MB for 131072 systems)",
              "advection": None, "mech32": None}
try:  # refreshed by tools/ncu_traffic.py from this round's captures
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _fh:
        _t = json.load(_fh)
    for _k, _v in _t.items():
        NCU_DRAM_BYTES_PER_SYSTEM[_k] = _v["bytes_per_system"]
        NCU_SOURCE[_k] = _v["source"]
except Exception:
    pass


# ------------------------------------------------------------------ algorithmic work ----
def _S(n, q):
    """flops of one step attempt outside the corrector (ewt + tolsf, predict,
    dvset, yh update, average rescale, acor scaling, eta logic)."""
    return 6 * n + 4 + n * q * (q + 1) / 2 + (q * q + q + 25) + 2 * n * (q + 1) + n * (q + 1) / 2 + n + 12


def _F_LU_dense(n):
    return n * n + n + sum(2 * m * m + m + 1 for m in range(1, n))


def flops_per_system_robertson(nst, nfe, nje, nlu, nni, ncfn, netf):
    """n = 3, mean order q = 4: S = 149.5 per attempt, I(3) = 56 per Newton iteration, F_f = 8,
    F_J = 6 (analytic), F_LU(3) = 27."""
    attempts = nst + netf + ncfn
    return attempts * _S(3, 4) + nni * (2 * 9 + 27 + 11) + nfe * 8.0 + nje * 6.0 + nlu * _F_LU_dense(3)


def flops_per_system_mech32(nst, nfe, nje, nlu, nni, ncfn, netf):
    """Dense n = 32, mf = 22 (finite-difference Jacobian).  F_f = 480: 32 unimolecular rates (1 flop),
    32 bimolecular rates (2 flops), 192 stoichiometric entries (multiply + add).  nfe already counts
    the n f-calls of every FD Jacobian (M:2976), so F_J carries only its own 2n + 4 per column + 3n + 8."""
    n, F_f = 32, 480.0
    attempts = nst + netf + ncfn
    return (attempts * _S(n, 4) + nni * (2 * n * n + 9 * n + 11) + nfe * F_f +
            nje * (n * (2 * n + 4) + 3 * n + 8) + nlu * _F_LU_dense(n))


def flops_per_system_advection(nst, nfe, nje, nlu, nni, ncfn, netf, ml=5, mu=0):
    """Banded n = 25, mf = 24 (analytic band Jacobian): F_f = 100 (4 flops / equation), F_J = n,
    F_bsl = n (1 + 2 ml + 2 (ml + mu)), I = F_bsl + 7 n + 11,
    F_LU = (2 ml + mu + 1) n + n + n (1 + ml + 2 ml (ml + mu))  (upper bounds)."""
    n = 25
    attempts = nst + netf + ncfn
    fbsl = n * (1 + 2 * ml + 2 * (ml + mu))
    flu = (2 * ml + mu + 1) * n + n + n * (1 + ml + 2 * ml * (ml + mu))
    return attempts * _S(n, 4) + nni * (fbsl + 7 * n + 11) + nfe * 100.0 + nje * n + nlu * flu


def bytes_per_system(neq, npar, nout, state_bytes):
    """Algorithmic HBM bytes of one system for one schedule launch: y0 + params + t + istate in;
    nout output points + y + t + istate + optional outputs + persistent state out."""
    return (8 * neq + 8 * npar + 8 + 4) + (nout * 8 * neq + 8 * neq + 8 + 4) + (4 * 8 + 12 * 4) + state_bytes


ROBERTSON_STATE_BYTES = 8 * (6 * 3 + 2 * 3 + 9 + 9 + 6 + 11) + 4 * (3 + 19)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Threads the CPU legs use: every core this process may run on.  torchrun exports
    OMP_NUM_THREADS=1, so the count is passed to the oracle explicitly, never left to OpenMP."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_run(nsample, tol, touts, first=0, nthreads=0, variant=""):
    """Run the CPU restatement (oracle/) on `nsample` systems of the same batch; returns
    (result dict, seconds, threads actually used)."""
    import oracle_lib as O
    from dvode_b200 import batches
    nthreads = nthreads or host_threads()
    par = batches.robertson_batch(nsample, first=first)
    O.batch_solve("robertson", par[:256], batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                  tol["atol"], 21, itol=2, nthreads=nthreads, variant=variant)  # warm the thread pool
    t0 = time.perf_counter()
    r = O.batch_solve("robertson", par, batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                      tol["atol"], 21, itol=2, nthreads=nthreads, variant=variant)
    dt = time.perf_counter() - t0
    assert (r["istate"][:, -1] == 2).all()
    return r, dt, O.threads_used(variant)


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement, kind = "port") on all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dvode_b200 import batches
    tol = batches.ROBERTSON_TOL_HEADLINE
    touts = batches.robertson_touts()
    nthreads = host_threads()
    nsample = 1 << 15
    import oracle_lib as O
    par = batches.robertson_batch(nsample)
    for _ in range(args.warmup):
        O.batch_solve("robertson", par[:2048], batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                      tol["atol"], 21, itol=2, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = O.batch_solve("robertson", par, batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                          tol["atol"], 21, itol=2, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / args.steps
    used = O.threads_used()
    assert (r["istate"][:, -1] == 2).all()
    v = nsample / dt
    sample = "%d of the 2^20 systems per step (same seed, tolerances, tout schedule)" % nsample
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "batched Robertson, perturbed rate constants, mf=21, rtol=1e-6, "
                                   "12 decades to t=4e10", "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": sample, "host_cpus": os.cpu_count(),
                             "note": "C restatement of dvode, gcc -O3 -ffp-contract=off, one solver "
                                     "instance per thread (OpenMP, thread count passed explicitly and "
                                     "read back from omp_get_num_threads); one host whatever --gpus is; "
                                     "no Fortran compiler in the image"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def parity_of_fast_build(got, ref, ref_fma, tol, iout_gpu):
    """north_star's bar -- every state within 10x the rtol-weighted error, batch counters within
    +-2 % -- measured for the fast build against the oracle sample, next to the same figures for
    the reference arithmetic against itself with FMA contraction (two round-off-different VODE runs
    separate: that is the yardstick tests/test_gpu_parity.py holds the fast build to)."""
    wt = tol["rtol"][0] * np.abs(ref["y"]) + np.array(tol["atol"])
    werr = (np.abs(got - ref["y"]) / wt).max(axis=2)
    out = {"sample_systems": int(ref["y"].shape[0]),
           "share_within_10x": float((werr <= 10.0).mean()),
           "share_within_10x_first_two_outputs": float((werr[:, :2] <= 10.0).mean()),
           "max_weighted_error": float(werr.max()),
           "median_weighted_error_by_output": [float(x) for x in np.median(werr, axis=0)]}
    names = {"nst": 0, "nfe": 1, "nje": 2, "nlu": 8, "nni": 9, "ncfn": 10, "netf": 11}
    n = ref["y"].shape[0]
    out["counter_delta_pct"] = {
        k: float(100.0 * (iout_gpu[:n, c].astype(np.int64).sum() - ref["istats"][:, -1, c].astype(np.int64).sum())
                 / max(1, ref["istats"][:, -1, c].astype(np.int64).sum())) for k, c in names.items()}
    if ref_fma is not None:
        m = ref_fma["y"].shape[0]
        wself = (np.abs(ref_fma["y"] - ref["y"][:m]) / wt[:m]).max(axis=2)
        out["reference_vs_reference_fma"] = {
            "sample_systems": int(m), "share_within_10x": float((wself <= 10.0).mean()),
            "max_weighted_error": float(wself.max()),
            "counter_delta_pct": {k: float(100.0 * (ref_fma["istats"][:, -1, c].astype(np.int64).sum() -
                                                      ref["istats"][:m, -1, c].astype(np.int64).sum())
                                           / max(1, ref["istats"][:m, -1, c].astype(np.int64).sum()))
                                  for k, c in names.items()}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nsys", type=int, default=NSYS_PER_GPU, help="systems per GPU")
    ap.add_argument("--tol", default="headline", choices=["headline", "ref"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the banded and the 32-species configurations")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling lines")
    ap.add_argument("--no-strict", action="store_true", help="skip the strict build's throughput")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import dvode_b200 as B
    from dvode_b200 import batches, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = sharding.make_comm(dist, rank, world, local)  # the C ABI's own NCCL communicator
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return float(tt.item())

    class Gathered:
        """rank 0's device arrays of the final gather (y, t, istate, rwork(11:14), iwork(11:22))."""

        def __init__(self, nsys_total, neq):
            self.n = nsys_total
            if rank == 0:
                self.y = torch.empty(nsys_total, neq, dtype=torch.float64, device=dev)
                self.t = torch.empty(nsys_total, dtype=torch.float64, device=dev)
                self.ist = torch.empty(nsys_total, dtype=torch.int32, device=dev)
                self.rout = torch.empty(nsys_total, 4, dtype=torch.float64, device=dev)
                self.iout = torch.empty(nsys_total, 12, dtype=torch.int32, device=dev)
                self.ptrs = [self.y.data_ptr(), self.t.data_ptr(), self.ist.data_ptr(),
                             self.rout.data_ptr(), self.iout.data_ptr()]
            else:
                self.ptrs = [1, 1, 1, 1, 1]  # "wanted"
            self.bytes_from_peers = (nsys_total - nsys_total // world) * (neq * 8 + 8 + 4 + 32 + 48)

        def run(self, solver):
            solver.gather_device(comm, 0, self.n, *self.ptrs, stream=stream.cuda_stream)

    # --------------------------------------------------------------------------------------
    def run_config(name, mf, nsys, par_fn, y0, touts, tolc, iopt, rw_opts, iw_opts, steps, warm,
                   strict=False, timed_gather=True, e2e=True, keep=None):
        """Device-resident timing (+ gather at N > 1) and the host-buffer end-to-end timing of one
        configuration on this rank's block of `nsys` systems.  Returns a dict (rank-local numbers
        already reduced over ranks where the metric needs it)."""
        par_h = par_fn(nsys, first=rank * nsys)
        s = B.dvode_t().initialize(name, nsys, par_h, device=local, strict=strict)
        neq, nout = s.neq, touts.size
        y0_d = torch.tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64)[:, None], (neq, nsys))),
                            dtype=torch.float64, device=dev)
        y_d = torch.empty_like(y0_d)
        t_d = torch.zeros(nsys, dtype=torch.float64, device=dev)
        ist_d = torch.ones(nsys, dtype=torch.int32, device=dev)
        yout_d = torch.empty(nout, neq, nsys, dtype=torch.float64, device=dev)
        gat = Gathered(nsys * world, neq) if (world > 1 and timed_gather) else None

        def device_step(with_gather=True):
            y_d.copy_(y0_d)
            t_d.zero_()
            ist_d.fill_(1)
            s.solve_schedule_device(neq, y_d.data_ptr(), t_d.data_ptr(), touts, tolc["itol"],
                                    tolc["rtol"], tolc["atol"], 1, ist_d.data_ptr(), iopt, rw_opts, iw_opts, mf,
                                    yout_d.data_ptr(), stream.cuda_stream)
            if gat is not None and with_gather:
                gat.run(s)

        for _ in range(warm):
            device_step()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        wall0 = time.perf_counter()
        for k in range(steps):
            flush.fill_(k & 0xFF)  # evict L2 between timed iterations (outside the event pair)
            ev[k][0].record(stream)
            device_step(with_gather=False)
            ev[k][1].record(stream)
            if gat is not None:
                gat.run(s)
            ev[k][2].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        ms_solve = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
        ms_total = float(np.mean([e[0].elapsed_time(e[2]) for e in ev]))
        launches = s.last_launch_count() + 3 + (1 if gat is not None else 0)
        _, iout = s.stats()
        ok = int((ist_d == 2).sum().item())
        cnt = {k: float(iout[:, B.api.ISTAT[k]].astype(np.int64).sum()) for k in
               ("nst", "nfe", "nje", "nlu", "nni", "ncfn", "netf")}
        res = {"nsys": nsys, "neq": neq, "nout": nout, "ms_solve_local": ms_solve,
               "ms_solve": max_over_ranks(ms_solve), "ms": max_over_ranks(ms_total),
               "ok": int(sum_over_ranks(ok)), "counters": cnt, "launches_per_step": launches,
               "wall": wall, "gather_bytes_into_rank0": gat.bytes_from_peers if gat else 0}
        res["value"] = nsys * world / (res["ms"] * 1e-3)
        if gat is not None:
            res["gather_ms"] = max_over_ranks(float(np.mean([e[1].elapsed_time(e[2]) for e in ev])))
        if keep is not None:
            keep["iout_device_run"] = iout

        # ---- end to end through the host-buffer C-ABI call, pinned host memory, rwork / iwork headers
        if e2e:
            y_h = torch.empty(nsys, neq, dtype=torch.float64).pin_memory()
            t_h = torch.empty(nsys, dtype=torch.float64).pin_memory()
            ist_h = torch.empty(nsys, dtype=torch.int32).pin_memory()
            par_p = torch.from_numpy(np.ascontiguousarray(par_h)).pin_memory()
            yout_h = torch.empty(nout, nsys, neq, dtype=torch.float64).pin_memory()
            rw_h = torch.zeros(nsys, 20, dtype=torch.float64).pin_memory()
            iw_h = torch.zeros(nsys, 30, dtype=torch.int32).pin_memory()
            y_n, t_n, ist_n, yout_n, par_n, rw_n, iw_n = (y_h.numpy(), t_h.numpy(), ist_h.numpy(), yout_h.numpy(),
                                                          par_p.numpy(), rw_h.numpy(), iw_h.numpy())
            gh = None
            if gat is not None and rank == 0:  # rank 0's pinned destinations of the gathered results
                gh = [torch.empty_like(x, device="cpu").pin_memory() for x in (gat.y, gat.t, gat.ist, gat.rout, gat.iout)]

            def e2e_step():
                y_n[:] = np.asarray(y0)
                t_n[:] = 0.0
                ist_n[:] = 1
                if rw_opts is not None:
                    rw_n[0, :10] = rw_opts[:10]
                if iw_opts is not None:
                    iw_n[0, :10] = iw_opts[:10]
                barrier()
                t0 = time.perf_counter()
                s.set_params(par_n)
                s.solve_schedule(neq, y_n, t_n, touts, tolc["itol"], tolc["rtol"], tolc["atol"], 1, ist_n, iopt,
                                 rw_n, 20, iw_n, 30, mf, yout_n)
                t1 = time.perf_counter()
                if gat is not None:  # the final gather to rank 0 and its copy to rank 0's host
                    gat.run(s)
                    if rank == 0:
                        for dst, src in zip(gh, (gat.y, gat.t, gat.ist, gat.rout, gat.iout)):
                            dst.copy_(src, non_blocking=True)
                    torch.cuda.synchronize()
                t2 = time.perf_counter()
                return t2 - t0, t2 - t1

            e2e_step()
            tt = [e2e_step() for _ in range(max(2, min(steps, 3)))]
            e2e_ms = max_over_ranks(float(np.mean([a for a, _ in tt])) * 1e3)
            res["e2e"] = {
                "value": nsys * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": nsys * (neq * 8 + 8 + 4 + par_h.shape[1] * 8),
                "d2h_bytes_per_step": nsys * (nout * neq * 8 + neq * 8 + 8 + 4 + 4 * 8 + 12 * 4),
                "all_systems_istate_2": bool(sum_over_ranks(int((ist_n == 2).all())) == world),
                "optional_outputs": "rwork(11:14), iwork(11:22) of every system written into rwork[nsys][20] / iwork[nsys][30]",
                "gpu_launches_per_step": s.last_launch_count() + 1 + (1 if gat is not None else 0)}
            if gat is not None:
                res["e2e"]["gather_and_d2h_on_rank0_ms"] = max_over_ranks(float(np.mean([b for _, b in tt])) * 1e3)
                res["e2e"]["d2h_bytes_per_step_rank0_gathered"] = nsys * world * (neq * 8 + 8 + 4 + 32 + 48)
            if keep is not None:
                keep["yout_host"] = yout_n
                keep["iwork_host"] = iw_n
        s.close()
        return res

    tol = batches.ROBERTSON_TOL_HEADLINE if args.tol == "headline" else batches.ROBERTSON_TOL_REF
    touts = batches.robertson_touts()
    nsys = args.nsys
    warm = max(args.warmup, 3)

    peak_tf = B.fp64_peak_tflops(local)  # measured DFMA peak (roofline denominator)
    hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = json.load(fh)["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        pass

    def roofline(kernel, flops_fn, r, state_bytes, npar, key, extra=None):
        kern_s = r["ms_solve_local"] * 1e-3
        fl = flops_fn(**r["counters"])  # summed over this rank's batch
        by = bytes_per_system(r["neq"], npar, r["nout"], state_bytes)
        tr = NCU_DRAM_BYTES_PER_SYSTEM.get(key)
        out = {"bound": "fp64", "kernel": kernel, "achieved": fl / kern_s / 1e12, "peak": peak_tf,
               "unit": "TFLOP/s", "frac": fl / kern_s / 1e12 / peak_tf,
               "peak_source": "DFMA microbenchmark measured in this run (bvode_measure_fp64_peak); "
                              "MEASURED_PEAKS.json has no FP64 entry",
               "algorithmic_flops_per_system": fl / r["nsys"],
               "kernel_ms": r["ms_solve_local"],
               "traffic": tr * r["nsys"] if tr else None, "traffic_source": NCU_SOURCE.get(key),
               "hbm": {"achieved": by * r["nsys"] / kern_s / 1e9, "peak": hbm_peak, "peak_source": hbm_src,
                       "unit": "GB/s", "frac": by * r["nsys"] / kern_s / 1e9 / hbm_peak,
                       "algorithmic_bytes_per_system": by}}
        if extra:
            out.update(extra)
        return out

    # ---- the headline configuration ---------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    keep = {}
    head = run_config("robertson", 21, nsys, batches.robertson_batch, batches.ROBERTSON_Y0, touts, tol, 0,
                      None, None, args.steps, warm, keep=keep)
    clocks = sampler.stop()

    # ---- the bit-exact (strict) build on the same batch, device-resident ---------------------
    strict = None
    if not args.no_strict:
        r = run_config("robertson", 21, nsys, batches.robertson_batch, batches.ROBERTSON_Y0, touts, tol, 0,
                       None, None, 2, 1, strict=True, e2e=False)
        strict = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms"], "systems_ok": r["ok"],
                  "what": "BVODE_FLAG_STRICT: -fmad=false, IEEE division, portable pow; every output bit-for-bit "
                          "equal to the oracle (tests/test_gpu_parity.py)"}

    # ---- strong scaling: a fixed TOTAL over the N GPUs (device-resident, gather included) -----
    strong = None
    if not args.no_strong:
        strong = []
        for lg in (20, 24):
            per = (1 << lg) // world
            r = run_config("robertson", 21, per, batches.robertson_batch, batches.ROBERTSON_Y0, touts, tol, 0,
                           None, None, 2, 1, e2e=False)
            strong.append({"scaling": "strong", "total_systems": per * world, "systems_per_gpu": per,
                           "value": r["value"], "unit": UNIT, "ms_per_step": r["ms"], "ms_solve": r["ms_solve"],
                           "final_gather_ms": r.get("gather_ms"), "systems_ok": r["ok"]})

    # ---- the other configurations of BASELINE.json (configs[2], configs[3]) --------------------
    other = None
    if not args.no_other_configs:
        other = []
        iw_adv = np.zeros(30, dtype=np.int32)
        iw_adv[0] = 5
        iw_mech = np.zeros(30, dtype=np.int32)
        iw_mech[5] = 100000
        for (name, mf, lg, par_fn, y0c, toutc, tolc, iopt, rw, iw, flops_fn, kernel, npar) in (
                ("advection", 24, 18, batches.advection_batch, batches.advection_y0(), batches.advection_touts(),
                 batches.ADVECTION_TOL, 0, np.zeros(20), iw_adv, flops_per_system_advection,
                 "vode_warp_kernel<WarpBandPolicy<Advection25, miter 4>>", 2),
                ("mech32", 22, 18, batches.mech32_batch, batches.mech32_y0(), batches.mech32_touts(),
                 batches.MECH32_TOL, 1, np.zeros(20), iw_mech, flops_per_system_mech32,
                 "vode_warp_kernel<WarpDensePolicy<Mech32, miter 2>>", 64)):
            nc = 1 << lg
            r = run_config(name, mf, nc, par_fn, y0c, toutc, tolc, iopt, rw, iw, 2, 1)
            neq = r["neq"]
            state = 8 * ((6 + 14) + 32 * 8 + ((2 * 5 + 0 + 1) | 1) * 25 + 6 * 25) if name == "advection" else \
                8 * ((6 + 14) + 32 * (8 + 32) + 32 * 33 + 16)
            other.append({"workload": "%s (BASELINE.json configs[%d]), mf=%d, 2^%d systems per GPU, weak scaling, "
                                      "2 timed steps after a warm-up" % (name, 2 if name == "advection" else 3, mf, lg),
                          "value": r["value"], "unit": UNIT, "ms_per_step": r["ms"], "ms_solve": r["ms_solve"],
                          "final_gather_ms": r.get("gather_ms"), "systems_ok": r["ok"],
                          "counters_per_system": {k: v / nc for k, v in r["counters"].items()},
                          "e2e": r["e2e"], "roofline": roofline(kernel, flops_fn, r, state, npar, name)})

    # ---- CPU baseline (rank 0, N = 1) and the fast build's parity on the same sample ------------
    cpu_base, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = host_threads()
        # ~10-30 s of CPU work: ~4000 systems/s/core
        nsample = min(nsys, 1 << int(np.clip(np.floor(np.log2(15.0 * 4000.0 * nthreads)), 14, 20)))
        ref, dt, used = cpu_port_run(nsample, tol, touts, nthreads=nthreads)
        cpu_base = {"value": nsample / dt, "unit": UNIT, "cores": used, "kind": "port", "host_cpus": os.cpu_count(),
                    "sample": "first %d of the 2^20 systems (%.1f s of CPU work on %d threads)" % (nsample, dt, used),
                    "note": "C restatement of dvode (oracle/), gcc -O3 -ffp-contract=off, one solver "
                            "instance per thread; the image has no Fortran compiler"}
        nfma = min(nsample, 1 << 13)
        ref_fma, _, _ = cpu_port_run(nfma, tol, touts, nthreads=nthreads, variant="fma")
        got = keep["yout_host"][:, :nsample].transpose(1, 0, 2)
        parity = parity_of_fast_build(got, ref, ref_fma, tol, keep["iwork_host"][:, 10:22])
        parity["what"] = ("fast (measured) build, every output point of the first %d systems of this run's end-to-end "
                          "step against the oracle; weighted error = |y - y_ref| / (rtol |y_ref| + atol)" % nsample)

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": head["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched Robertson (BASELINE.json configs[1]): %d systems per GPU, "
                                   "rate constants perturbed +-10%% (splitmix64 seed %d), mf=21, "
                                   "itol=2, rtol=%g, atol=%s, 12 decades to t=4e10, one schedule "
                                   "launch per step" % (nsys, batches.SEED, tol["rtol"][0], tol["atol"]),
                       "systems_per_gpu": nsys, "l2": "flushed between timed iterations (256 MB write)",
                       "parallelism": "index sharding, %d rank(s), no data-path collective; the final NCCL gather "
                                      "of y / t / istate / optional outputs to rank 0 is inside the timed region" % world},
            "systems_ok": head["ok"], "wall_s_timed_region": head["wall"],
            "gpu_launches": head["launches_per_step"] * args.steps,
            "clocks": clocks,
            "e2e": head["e2e"],
            "roofline": roofline("vode_thread_kernel<Robertson, mf=21>", flops_per_system_robertson, head,
                                 ROBERTSON_STATE_BYTES, 3, "robertson"),
            "counters_per_system": {k: v / nsys for k, v in head["counters"].items()},
            "ms_solve": head["ms_solve"],
        }
        if world > 1:
            line["final_gather_ms"] = head["gather_ms"]
            line["final_gather_gbs_into_rank0"] = head["gather_bytes_into_rank0"] / (head["gather_ms"] * 1e-3) / 1e9
        if strict is not None:
            line["strict"] = strict
        if strong is not None:
            line["strong_scaling"] = strong
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if parity is not None:
            line["parity"] = parity
        if other is not None:
            line["other_configs"] = other
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
