#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched VODE hot path (BASELINE.json configs[1]).

Workload: 2^20 Robertson systems per GPU with perturbed rate constants, mf = 21 (dense
analytic Jacobian), integrated from t = 0 through the reference's 12-decade output schedule
to t = 4e10 (test/example.f90:56-66), headline tolerances rtol = 1e-6,
atol = (1e-10, 1e-16, 1e-8).  One "step" = one pass of that whole integration over the batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = systems/s with inputs resident in HBM (CUDA events
around the solve launches); `e2e` = the same through the host-buffer C-ABI call
(bvode_solve_schedule) including H2D of y0/t/istate/params and D2H of every output point.
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
# measured DRAM traffic of the headline kernel (ncu, see roofline.traffic_source)
NCU_DRAM_BYTES_PER_SYSTEM = 1137


def flops_per_system_robertson(nst, nfe, nje, nlu, nni, ncfn, netf):
    """algorithmic FP64 flops of one Robertson system from the solver's
    own counters.  n = 3, mean order q = 4: S = 149.5 per attempt, I(3) = 56 per Newton
    iteration, F_f = 8, F_J = 6, F_LU(3) = 27."""
    attempts = nst + netf + ncfn
    return attempts * 149.5 + nni * 56.0 + nfe * 8.0 + nje * 6.0 + nlu * 27.0


def bytes_per_system_robertson(nout):
    """Algorithmic HBM bytes of one system for one schedule launch: y0 + params + t + istate in;
    nout output points + y + t + istate + optional outputs + persistent state out."""
    state = 8 * (6 * 3 + 2 * 3 + 9 + 9 + 6 + 11) + 4 * (3 + 19)
    return (24 + 24 + 8 + 4) + (nout * 24 + 24 + 8 + 4) + (4 * 8 + 12 * 4) + state


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


def cpu_port_throughput(nsample, tol, touts, first=0, nthreads=0, repeats=1):
    """Time the CPU restatement (oracle/) on `nsample` systems of the same batch."""
    import oracle_lib as O
    from dvode_b200 import batches
    par = batches.robertson_batch(nsample, first=first)
    O.batch_solve("robertson", par[:256], batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                  tol["atol"], 21, itol=2, nthreads=nthreads)  # warm the thread pool
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.batch_solve("robertson", par, batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                          tol["atol"], 21, itol=2, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert (r["istate"][:, -1] == 2).all()
    return nsample / best, best


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement, kind = "port") on all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dvode_b200 import batches
    tol = batches.ROBERTSON_TOL_HEADLINE
    touts = batches.robertson_touts()
    cores = os.cpu_count() or 1
    nsample = 1 << 15
    import oracle_lib as O
    par = batches.robertson_batch(nsample)
    for _ in range(args.warmup):
        O.batch_solve("robertson", par[:2048], batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                      tol["atol"], 21, itol=2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = O.batch_solve("robertson", par, batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"],
                          tol["atol"], 21, itol=2)
    dt = (time.perf_counter() - t0) / args.steps
    assert (r["istate"][:, -1] == 2).all()
    v = nsample / dt
    sample = "%d of the 2^20 systems per step (same seed, tolerances, tout schedule)" % nsample
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "batched Robertson, perturbed rate constants, mf=21, rtol=1e-6, "
                                   "12 decades to t=4e10", "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample,
                             "note": "C restatement of dvode, gcc -O3 -ffp-contract=off, one solver "
                                     "instance per core (OpenMP); no Fortran compiler in the image"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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
                    help="skip the device-resident numbers of the banded and the 32-species configurations")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import dvode_b200 as B
    from dvode_b200 import batches

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nsys = args.nsys
    tol = batches.ROBERTSON_TOL_HEADLINE if args.tol == "headline" else batches.ROBERTSON_TOL_REF
    touts = batches.robertson_touts()
    nout = touts.size
    neq = 3
    # this rank's contiguous slice of the global batch (index sharding)
    par_h = batches.robertson_batch(nsys, first=rank * nsys)

    s = B.dvode_t().initialize("robertson", nsys, par_h, device=local)
    # pristine inputs resident in HBM, SoA [neq][nsys]
    y0_d = torch.zeros(neq, nsys, dtype=torch.float64, device=dev)
    y0_d[0].fill_(1.0)
    y_d = torch.empty_like(y0_d)
    t_d = torch.zeros(nsys, dtype=torch.float64, device=dev)
    ist_d = torch.ones(nsys, dtype=torch.int32, device=dev)
    yout_d = torch.empty(nout, neq, nsys, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    stream = torch.cuda.current_stream()

    def device_step():
        y_d.copy_(y0_d)
        t_d.zero_()
        ist_d.fill_(1)
        s.solve_schedule_device(neq, y_d.data_ptr(), t_d.data_ptr(), touts, tol["itol"],
                                tol["rtol"], tol["atol"], 1, ist_d.data_ptr(), 0, None, None, 21,
                                yout_d.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    peak_tf = B.fp64_peak_tflops(local)  # measured DFMA peak (roofline denominator)

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        device_step()
        ev[k][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop()
    ms_step = float(np.mean(ms))
    launches_per_step = s.last_launch_count() + 3  # solve kernel(s) + the 3 input-reset kernels

    # batch counters for the roofline (from the device-resident optional outputs)
    _, iout = s.stats()
    ok = int((ist_d == 2).sum().item())
    c = {k: float(iout[:, B.api.ISTAT[k]].astype(np.int64).sum()) for k in
         ("nst", "nfe", "nje", "nlu", "nni", "ncfn", "netf")}
    flops_launch = flops_per_system_robertson(**c)  # already summed over the batch
    bytes_launch = bytes_per_system_robertson(nout) * nsys

    t_ms = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    tot_ok = torch.tensor([ok], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_ok, op=dist.ReduceOp.SUM)
    ms_max = float(t_ms.item())
    value = nsys * world / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI call, pinned host memory ----------------
    y_h = torch.empty(nsys, neq, dtype=torch.float64).pin_memory()
    t_h = torch.empty(nsys, dtype=torch.float64).pin_memory()
    ist_h = torch.empty(nsys, dtype=torch.int32).pin_memory()
    par_p = torch.from_numpy(par_h).pin_memory()
    yout_h = torch.empty(nout, nsys, neq, dtype=torch.float64).pin_memory()
    y_n, t_n, ist_n, yout_n, par_n = (y_h.numpy(), t_h.numpy(), ist_h.numpy(), yout_h.numpy(),
                                      par_p.numpy())

    def e2e_step():
        y_n[:] = batches.ROBERTSON_Y0
        t_n[:] = 0.0
        ist_n[:] = 1
        t0 = time.perf_counter()
        s.set_params(par_n)
        s.solve_schedule(neq, y_n, t_n, touts, tol["itol"], tol["rtol"], tol["atol"], 1, ist_n, 0,
                         None, 0, None, 0, 21, yout_n)
        return time.perf_counter() - t0

    e2e_step()
    barrier()
    e2e_t = [e2e_step() for _ in range(max(2, min(args.steps, 3)))]
    e2e_launches = s.last_launch_count() + 1
    e2e_ms = torch.tensor([float(np.mean(e2e_t)) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = nsys * world / (float(e2e_ms.item()) * 1e-3)
    e2e_ok = bool((ist_n == 2).all())
    h2d = nsys * (neq * 8 + 8 + 4 + 3 * 8)
    d2h = nsys * (nout * neq * 8 + neq * 8 + 8 + 4)

    # ---- multi-GPU: the only exchange is the final gather of solutions and statistics -------
    gather_ms = None
    if world > 1:
        from dvode_b200 import sharding
        final = yout_d[-1].t().contiguous()                      # [nsys, neq]
        stats_d = torch.from_numpy(iout.astype(np.int32)).to(dev)  # [nsys, 12]
        sharding.gather_to_rank0(final, dist, rank, world)        # warm-up (NCCL init)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        yall = sharding.gather_to_rank0(final, dist, rank, world)
        sall = sharding.gather_to_rank0(stats_d, dist, rank, world)
        g1.record()
        torch.cuda.synchronize()
        if rank == 0:
            assert yall.shape[0] == nsys * world and sall.shape[0] == nsys * world
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = float(gt.item())

    # ---- the other single-GPU configurations of BASELINE.json, device-resident, for the record ----
    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other = []
        for name, mf, lg, par_fn, y0c, toutc, tolc, iopt, ml, mxs in (
                ("advection", 24, 18, batches.advection_batch, batches.advection_y0(), batches.advection_touts(),
                 batches.ADVECTION_TOL, 0, 5, 0),
                ("mech32", 22, 16, batches.mech32_batch, batches.mech32_y0(), batches.mech32_touts(),
                 batches.MECH32_TOL, 1, 0, 100000)):
            nc = 1 << lg
            sc = B.dvode_t().initialize(name, nc, par_fn(nc))
            y0d = torch.tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(y0c)[:, None], (sc.neq, nc))),
                               dtype=torch.float64, device=dev)
            ydc, tdc = torch.empty_like(y0d), torch.zeros(nc, dtype=torch.float64, device=dev)
            isc = torch.ones(nc, dtype=torch.int32, device=dev)
            iwc = np.zeros(30, dtype=np.int32)
            iwc[0], iwc[5] = ml, mxs
            best = None
            for rep in range(3):
                ydc.copy_(y0d); tdc.zero_(); isc.fill_(1)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                sc.solve_schedule_device(sc.neq, ydc.data_ptr(), tdc.data_ptr(), toutc, tolc["itol"], tolc["rtol"],
                                         tolc["atol"], 1, isc.data_ptr(), iopt, np.zeros(20), iwc, mf, 0,
                                         stream.cuda_stream)
                c1.record(stream)
                torch.cuda.synchronize()
                if rep > 0:
                    best = c0.elapsed_time(c1) if best is None else min(best, c0.elapsed_time(c1))
            other.append({"workload": "%s, mf=%d, 2^%d systems, device-resident, best of 2 after a warm-up" % (name, mf, lg),
                          "value": nc / best * 1e3, "unit": UNIT, "ms": best,
                          "systems_ok": int((isc == 2).sum().item())})
            sc.close()
            del y0d, ydc, tdc, isc

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ~10-30 s of CPU work: ~4000 systems/s/core -> 2^20 systems on a 16-core box
        cores = os.cpu_count() or 1
        nsample = 1 << int(np.clip(np.floor(np.log2(15.0 * 4000.0 * cores)), 14, 20))
        v, dt = cpu_port_throughput(nsample, tol, touts)
        cpu_base = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": "first %d of the 2^20 systems (%.1f s of CPU work on all cores)" % (nsample, dt),
                    "note": "C restatement of dvode (oracle/), gcc -O3 -ffp-contract=off, one solver "
                            "instance per core; the image has no Fortran compiler"}

    if rank == 0:
        kern_s = ms_step * 1e-3
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = json.load(fh)["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched Robertson (BASELINE.json configs[1]): %d systems per GPU, "
                                   "rate constants perturbed +-10%% (splitmix64 seed %d), mf=21, "
                                   "itol=2, rtol=%g, atol=%s, 12 decades to t=4e10, one schedule "
                                   "launch per step" % (nsys, batches.SEED, tol["rtol"][0], tol["atol"]),
                       "systems_per_gpu": nsys, "l2": "flushed between timed iterations (256 MB write)",
                       "parallelism": "index sharding, %d rank(s), no data-path collective" % world},
            "systems_ok": int(tot_ok.item()), "wall_s_timed_region": wall,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": float(e2e_ms.item()),
                    "all_systems_istate_2": e2e_ok, "gpu_launches_per_step": e2e_launches},
            "roofline": {"bound": "fp64", "kernel": "vode_thread_kernel<Robertson, mf=21>",
                         "achieved": flops_launch / kern_s / 1e12, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": flops_launch / kern_s / 1e12 / peak_tf,
                         "peak_source": "DFMA microbenchmark measured in this run (bvode_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic_flops_per_system": flops_launch / nsys,
                         "traffic": NCU_DRAM_BYTES_PER_SYSTEM * nsys,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                           "capture (profiles/r01_ncu_thread_kernel_r1p.txt: 149.0 MB for 131072 "
                                           "systems = 1137 B/system vs 1024 algorithmic), scaled to this launch",
                         "fp64_pipe_active_ncu": 0.301, "issue_active_ncu": 0.483,
                         "hbm": {"achieved": bytes_launch / kern_s / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": bytes_launch / kern_s / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_system": bytes_per_system_robertson(nout)}},
            "counters_per_system": {k: v / nsys for k, v in c.items()},
        }
        if gather_ms is not None:
            line["final_gather_ms"] = gather_ms
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if other is not None:
            line["other_configs"] = other
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
