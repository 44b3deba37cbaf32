#!/usr/bin/env python
"""bench.py -- conditioning sets/sec of createU (U_NZentries) on B200, per the round contract.

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N ...            CPU arm: the restated reference
                                                           (oracle/, OpenMP + LAPACK) on host cores

Workload (BASELINE.json configs[1]): n = 1e6 uniform 2-D locations per GPU (weak scaling: n =
N * 1e6, rows sharded by contiguous range, every rank holds all locations), m = 30, Matern
nu = 1.5 closed form, standard Vecchia ('z') conditioning, per-location nuggets.  One step = one
U_NZentries pass over the rank's rows.  `value` times the device-resident call (gpv_u_dev);
`e2e` times the reference-facing call with host buffers (gpv_u_values_packed: nuggets H2D,
packed U values D2H) -- the call createU() makes.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIG2 = 1.0


def flops_per_set(p, d, cov):
    """every +,-,*,/,sqrt,exp,pow,K_nu counted as 1."""
    c_cov = {"nu0.5": 3, "nu1.5": 6, "nu2.5": 9, "esqe": 8, "general": 4}[cov]
    P = p * (p - 1) // 2
    return p ** 3 / 3 + p ** 2 / 2 + p / 6 + p ** 2 + P * 3 * d + P * c_cov


def bytes_per_set(p, d):
    return p * 4 + p * d * 8 + 8 + p * 8 + p * 8


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML through
    pynvml (~1 ms per sample), falling back to the nvidia-smi query of the profiling recipe."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.reasons, self.max_sm, self.power = [], set(), None, []
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[0]))
        self.max_sm = max(self.max_sm or 0.0, float(r[1]))
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.002 if self._nvml is not None else 0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None, sm_max_mhz=self.max_sm,
                    reasons=sorted(self.reasons), samples=len(self.sm),
                    power_w_max=max(self.power) if self.power else None,
                    source="nvml" if self._nvml is not None else "nvidia-smi")


WORKLOADS = {
    "cfg1": dict(n=10_000, scaling="strong", d=2, m=20, covType="matern", nu=1.5, tag="nu1.5", layout="zy", n_pred=0,
                 doc="BASELINE configs[0]: n=1e4 2-D, m=20, Matern nu=1.5, response-first (zy) layout"),
    "cfg2": dict(n=1_000_000, scaling="weak", d=2, m=30, covType="matern", nu=1.5, tag="nu1.5", layout="z", n_pred=0,
                 doc="BASELINE configs[1]: createU, 2-D, m=30, Matern nu=1.5 closed form, 1e6 rows per GPU"),
    "cfg3": dict(n=10_000_000, scaling="strong", d=2, m=30, covType="matern", nu=0.8, tag="general", layout="z", n_pred=0,
                 doc="BASELINE configs[2]: n=1e7 total, general-nu Matern (nu=0.8), rows sharded over the GPUs"),
    "cfg4": dict(n=4_000_000, scaling="strong", d=3, m=40, covType="esqe", nu=None, tag="esqe", layout="z", n_pred=0,
                 doc="BASELINE configs[3]: n=4e6 total 3-D, m=40 (one set per warp), esqe covariance"),
    "cfg5": dict(n=2_000_000, scaling="strong", d=2, m=30, covType="matern", nu=1.5, tag="nu1.5", layout="zy", n_pred=500_000,
                 doc="BASELINE configs[4]: obs+pred joint ordering, n_obs=2e6 + n_pred=5e5, m=30, zy layout (4.5e6 rows, 2.5e6 full)"),
}


def make_inputs(wl, n_total, world, rank, device, use_gpu_nn=True):
    """Synthetic problem of the workload; returns a dict with the rank's rows of revNN/revCond and
    the replicated arrays.  n_total = number of observed locations."""
    from gpvecchia_b200 import harness as H
    from gpvecchia_b200 import shard
    d, m = wl["d"], wl["m"]
    rng_ = H.default_range(n_total, d)
    covparms = np.array([SIG2, rng_, wl["nu"]]) if wl["covType"] == "matern" else np.array([1.0, rng_, 0.5, rng_])
    locs_obs = H.make_locs(n_total, d, stream=2)
    tau = H.make_nuggets(n_total, stream=2)
    z = H.make_data(n_total, stream=2)
    if wl["layout"] == "z":
        # 'z' conditioning: neighbours on the response, self on the latent (vecchia_specify.R:189-190)
        # equal expected kernel time per rank (late rows gather from more than the L2 holds); the closed
        # forms feel it twice as much, relatively, as the slower general-nu kernel
        cuts = shard.locality_cuts(n_total, world, d, penalty_scale=1.0 if wl["tag"] != "general" else 0.5)
        rb, re_ = int(cuts[rank]), int(cuts[rank + 1])
        if use_gpu_nn:
            revNN = H.ordered_nn_gpu(locs_obs, m, rb, re_, device=device)
        else:
            revNN = H.rev(H.ordered_nn_kdtree(locs_obs, m, rb, re_)).astype(np.int32)
        revCond = np.zeros(revNN.shape, dtype=np.int32)
        revCond[revNN == 0] = np.iinfo(np.int32).min
        revCond[:, -1] = 1
        nfull_total = n_total - 1
        return dict(locs=locs_obs, revNN=revNN, revCond=revCond, obs=np.ones(n_total, dtype=np.int32),
                    nug_all=tau, nug_obs=tau, z=z, covparms=covparms, rb=rb, re=re_, N=n_total,
                    skip_rows=0, nfull_total=nfull_total, nfull_rank=int((revNN != 0).sum(axis=1).__ge__(2).sum()))
    # response-first zy layout, optionally with prediction locations (vecchia_specify.R:191-224)
    n_p = wl["n_pred"]
    if n_p > 0:
        locs_pred = H.make_locs(n_p, d, stream=3)
        locs2, NN, Cond, obs = H.layout_zy_pred(locs_obs, locs_pred, m, device=device, use_gpu=use_gpu_nn)
    else:
        locs2, NN, Cond, obs = H.layout_zy(locs_obs, m, n_total)
    N = locs2.shape[0]
    n0 = (NN != 0).sum(axis=1)
    cuts = shard.row_cuts(n0, world)              # balance sum n0^3: the dummy rows cost nothing
    rb, re_ = int(cuts[rank]), int(cuts[rank + 1])
    revNN = H.rev(NN[rb:re_]).astype(np.int32)
    revCond = H.rev(Cond[rb:re_]).astype(np.int32)
    revCond[revCond < 0] = np.iinfo(np.int32).min
    nug_all = np.concatenate([tau, np.zeros(N - n_total)])      # createU.R:75-77 with ord = identity
    return dict(locs=locs2, revNN=revNN, revCond=revCond, obs=obs.astype(np.int32), nug_all=nug_all,
                nug_obs=tau, z=z, covparms=covparms, rb=rb, re=re_, N=N, skip_rows=n_total,
                nfull_total=int((n0 >= 2).sum()), nfull_rank=int((n0[rb:re_] >= 2).sum()))


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so
    omp_get_max_threads() would understate the box; the oracle takes the team size as an argument
    (num_threads(Ncores), like src/U_NZentries.cpp:37)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_cpus(cuda_index):
    """N > 1 only: run this rank on the CPUs NVML reports as local to its GPU before any pinned host buffer is
    allocated, so that the ranks' 264 MB device-to-host copies land in the memory of the socket their GPU hangs
    on instead of all in one (profiles/r01_bench_n8.json: the end-to-end value fell from 3.6e8 at N = 4 to
    2.2e8 at N = 8).  GPV_BENCH_NUMA=0 turns it off.  Returns a description for `config`."""
    if os.environ.get("GPV_BENCH_NUMA", "1") != "1":
        return "off (GPV_BENCH_NUMA=0)"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(cuda_index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            hdl = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            hdl = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            bind_to_gpu_cpus.unbound = allowed      # restored before the CPU baseline, which uses every host core
            os.sched_setaffinity(0, cpus)
            return f"rank runs on the {len(cpus)} CPUs local to its GPU (of {len(allowed)})"
        return "one affinity domain: nothing to bind"
    except Exception as e:     # no NVML, no permission: measure unbound
        return f"unavailable ({type(e).__name__})"


def cpu_reference_rate(locs, revNN_rows, revCond_rows, row_begin, nuggets, covparms, target_s=12.0, threads=None,
                       covType="matern"):
    """Times the restated reference (oracle/: OpenMP schedule(static) + LAPACK dpotrf/dtrtrs) on a
    bounded sample of the same workload's rows; returns (sets/s, threads, sample description)."""
    import oracle as O
    threads = threads or host_threads()
    n_total = locs.shape[0]
    nr = revNN_rows.shape[0]

    def timed(nrows_s):
        pr = O.RowsProblem(locs, revNN_rows[nr - nrows_s:], revCond_rows[nr - nrows_s:], row_begin + nr - nrows_s,
                           nuggets, covType, covparms)
        t0 = time.perf_counter()
        pr.run(threads)
        return time.perf_counter() - t0
    pilot = min(20000, nr)
    timed(min(2000, nr))                      # thread-pool / page-fault warm-up
    t_p = timed(pilot)
    nrows_s = int(min(nr, max(pilot, pilot / t_p * target_s)))
    # about target_s seconds of CPU work in total: repeat the pass when the rank has too few rows
    passes = int(min(10, max(1, round(target_s / max(nrows_s * t_p / pilot, 1e-3)))))
    t_s = sum(timed(nrows_s) for _ in range(passes))
    lo = row_begin + nr - nrows_s
    return (nrows_s * passes / t_s, threads,
            f"rows [{lo},{lo + nrows_s}) of the n={n_total} workload x {passes} passes, {t_s:.1f} s of CPU time")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the workload's n (per GPU if weak, total if strong)")
    ap.add_argument("--host-nn", action="store_true", help="build neighbour arrays with cKDTree on the host")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.n:
        wl["n"] = args.n
    n_total = wl["n"] * world if wl["scaling"] == "weak" else wl["n"]
    m, d = wl["m"], wl["d"]
    p = m + 1
    cov_desc = f"Matern nu={wl['nu']}" if wl["covType"] == "matern" else "esqe"
    workload = (f"{args.workload}: createU/U_NZentries, n={n_total} uniform {d}-D locs"
                f"{' (' + str(wl['n']) + '/GPU on average)' if wl['scaling'] == 'weak' else ''}"
                f"{' + ' + str(wl['n_pred']) + ' prediction locs' if wl['n_pred'] else ''}, m={m}, {cov_desc}, "
                f"'{wl['layout']}' conditioning")
    per_row_mb = (p * 4 + p * 8) / 1e6
    config = dict(workload=workload, n=n_total, m=m, d=d, covmodel=wl["covType"], nu=wl["nu"], cond_yz=wl["layout"],
                  sharding=(f"rows by contiguous range over {world} rank(s), cut for equal expected kernel time "
                            "(late rows gather from more than the L2 holds: gpvecchia_b200/shard.py); locs and nuggets replicated"),
                  l2=(f"touched per step: {per_row_mb * n_total / world:.0f} MB of ids + U values per rank "
                      + ("(larger than the 126 MB L2)" if per_row_mb * n_total / world > 126 else
                         "(fits the 126 MB L2: this workload is launch-latency bound, not a bandwidth case)")))

    if args.impl == "reference":
        # ---- CPU arm: the restated reference (oracle/: OpenMP + LAPACK) on this box's host cores -----
        if rank != 0:
            return
        import oracle as O
        from gpvecchia_b200 import harness as H
        try:
            import gpvecchia_b200 as G
            have_gpu = G.lib.gpv_device_count() > 0
        except Exception:
            have_gpu = False
        use_gpu_nn = have_gpu and not args.host_nn
        if wl["layout"] == "z":
            n_s = min(n_total, 400_000)      # neighbour arrays for a bounded sample of the workload's rows
            locs = H.make_locs(n_total, d, stream=2)
            rb, re_ = n_total - n_s, n_total
            if use_gpu_nn:
                revNN = H.ordered_nn_gpu(locs, m, rb, re_, device=0)
            else:
                revNN = H.rev(H.ordered_nn_kdtree(locs, m, rb, re_)).astype(np.int32)
            revCond = np.zeros(revNN.shape, dtype=np.int32)
            revCond[revNN == 0] = np.iinfo(np.int32).min
            revCond[:, -1] = 1
            nuggets = H.make_nuggets(n_total, stream=2)
            rng_ = H.default_range(n_total, d)
            covparms = np.array([SIG2, rng_, wl["nu"]]) if wl["covType"] == "matern" else np.array([1.0, rng_, 0.5, rng_])
        else:
            pb = make_inputs(wl, n_total, 1, 0, 0, use_gpu_nn=use_gpu_nn)
            full = (pb["revNN"] != 0).sum(axis=1) >= 2
            locs, revNN, revCond, nuggets, covparms = pb["locs"], pb["revNN"][full], pb["revCond"][full], pb["nug_all"], pb["covparms"]
            n_s = min(revNN.shape[0], 400_000)
            revNN, revCond = revNN[-n_s:], revCond[-n_s:]
            rb, re_ = pb["N"] - n_s, pb["N"]
        threads = host_threads()
        rate0, _, _ = cpu_reference_rate(locs, revNN, revCond, rb, nuggets, covparms, target_s=2.0, covType=wl["covType"])
        rows_step = int(min(n_s, max(10000, rate0 * 3.0)))      # ~3 s of CPU work per step
        pr = O.RowsProblem(locs, revNN[-rows_step:], revCond[-rows_step:], re_ - rows_step, nuggets, wl["covType"], covparms)
        for _ in range(args.warmup):
            pr.run(threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pr.run(threads)
        dt = time.perf_counter() - t0
        value = rows_step * args.steps / dt
        sample = (f"{rows_step} full rows per step (rows [{re_ - rows_step},{re_}) of the n={n_total} workload), "
                  f"{threads} OpenMP threads, LAPACK={'openblas' if O.has_lapack() else 'textbook'}")
        print(json.dumps({
            "impl": "reference", "metric": "conditioning sets/sec (createU)", "value": value, "unit": "sets/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "sets/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ---- own arm --------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    import gpvecchia_b200 as G
    from gpvecchia_b200 import shard

    if G.lib.gpv_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the gpvecchia_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        config["host_binding"] = bind_to_gpu_cpus(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t_gen = time.perf_counter()
    pb = make_inputs(wl, n_total, world, rank, local_rank, use_gpu_nn=not args.host_nn)
    t_gen = time.perf_counter() - t_gen
    locs, revNN, revCond, covparms, z = pb["locs"], pb["revNN"], pb["revCond"], pb["covparms"], pb["z"]
    nuggets, nug_obs = pb["nug_all"], pb["nug_obs"]
    rb, re_ = pb["rb"], pb["re"]
    nrows = re_ - rb
    n_sets = pb["nfull_total"]            # conditioning sets with n0 >= 2 over all ranks: the unit of `value`
    # each rank hands over only its own rows of revNNarray / revCond (gpv_create_shard)
    h = G.UHandle(locs, revNN, revCond, obs=pb["obs"], row_begin=rb, row_end=re_, device=local_rank)
    covType = wl["covType"]

    d_nug = torch.from_numpy(nuggets).to(dev)
    d_out = torch.empty(nrows * p, dtype=torch.float64, device=dev)
    d_z = torch.from_numpy(z).to(dev)
    d_ll = torch.zeros(8, dtype=torch.float64, device=dev)
    # a dedicated (non-default) torch stream: its handle is what the C ABI launches on, and the
    # torch events below are recorded on the same stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_dev():
        h.u_dev(covType, covparms, d_nug.data_ptr(), d_out.data_ptr(), packed=False, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        h.kernel_time_stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = G.lib.gpv_launch_count()
    ms_total = timed(step_dev, args.steps, args.warmup)
    launches = int(G.lib.gpv_launch_count() - launches0) - 2 * args.warmup
    clocks = sampler.stop()
    value = n_sets * args.steps / (ms_total * 1e-3)
    # duration of the dominant kernel over the SAME timed region: one CUDA-event pair per launch,
    # recorded by the library on the launching stream
    k_count, k_total = h.kernel_time_stats(reset=True)
    k_ms = k_total / max(k_count, 1)
    kname = h.last_kernel_name()

    # ---- e2e: the reference-facing host-buffer call createU() makes --------------------------------
    total_packed = h.packed_len
    n_obs = n_total
    host_out = torch.empty(total_packed + 2 * n_obs, dtype=torch.float64).pin_memory()
    host_nug = torch.from_numpy(nuggets).pin_memory()
    host_tau = torch.from_numpy(nug_obs).pin_memory()
    out_np, nug_np, tau_np = host_out.numpy(), host_nug.numpy(), host_tau.numpy()

    def step_e2e():
        h.values_packed(covType, covparms, nug_np, tau_np, zentries_tail=True, out=out_np)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_sets * e2e_steps / float(t_e2e.item())
    h2d = 8 * nuggets.size + 8 * n_obs
    d2h = 8 * (total_packed + 2 * n_obs)

    # ---- extras: loglik evals/sec (fused numerator, scalars out), other covariances ----------------
    extras = {}
    if not args.no_extras:
        # the same end-to-end call delivering dgCMatrix@x (compressed-column order)
        try:
            ncols, nnz_csc, _ = h.csc_dims()
            host_csc = torch.empty(nnz_csc, dtype=torch.float64).pin_memory().numpy()
            for _ in range(2):
                h.values_csc(covType, covparms, nug_np, tau_np, out=host_csc)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                h.values_csc(covType, covparms, nug_np, tau_np, out=host_csc)
            barrier()
            t_csc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_csc, op=dist.ReduceOp.MAX)
            extras["e2e_csc_sets_per_s"] = n_sets * e2e_steps / float(t_csc.item())
            del host_csc
        except G.GpvError as e:                      # duplicate U rows in a set: triplet route only
            extras["e2e_csc_sets_per_s"] = None
        ll_steps = max(5, args.steps // 2)

        def step_ll():
            h.u_dev(covType, covparms, d_nug.data_ptr(), None, d_zord=d_z.data_ptr(), skip_rows=pb["skip_rows"],
                    d_loglik=d_ll.data_ptr(), stream=stream)
            if world > 1:
                shard.allreduce_loglik(d_ll)          # 3 doubles over NCCL: the path's only collective
        ms_ll = timed(step_ll, ll_steps, 3)
        # pure `z` layout: the same launch also accumulates the denominator terms, so this is the
        # whole vecchia_likelihood (R/vecchia_likelihood.R:14-27) with the data resident in HBM
        extras["loglik_evals_per_s"] = ll_steps / (ms_ll * 1e-3)
        extras["loglik_sets_per_s"] = n_sets * ll_steps / (ms_ll * 1e-3)
        parts = [float(v) for v in d_ll.cpu().tolist()]
        tau_terms = float(np.sum(z * z / nug_obs)), float(np.sum(np.log(nug_obs)))
        qn, ldn, qd, ldd = parts[0] + tau_terms[0], parts[1] + tau_terms[1], parts[3], parts[4]
        extras["loglik_parts"] = dict(quadform_num=qn, logdet_num=ldn, quadform_denom=qd, logdet_denom=ldd, nfail=parts[2])
        if wl["layout"] == "z":
            extras["loglik_value"] = -0.5 * (ldn - ldd + qn - qd + n_total * float(np.log(2 * np.pi)))
        if wl["layout"] == "z":
            # estimation loop: data and nuggets resident on the handle, only covparms go up (gpv_loglik_z
            # with NULL vectors), 6 doubles come back
            try:
                h.loglik_z(covType, covparms, nug_np, tau_np, z)
                barrier()
                t0 = time.perf_counter()
                for _ in range(ll_steps):
                    r_res = h.loglik_z(covType, covparms, None, None, None)
                barrier()
                t_res = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
                extras["loglik_e2e_resident_evals_per_s"] = ll_steps / float(t_res.item())
            except G.GpvError:
                pass
        if wl["layout"] == "z":
            # end-to-end likelihood call with host buffers (nuggets, tau, z up; 6 doubles back)
            host_z = torch.from_numpy(z).pin_memory().numpy()
            for _ in range(2):
                r_ll = h.loglik_z(covType, covparms, nug_np, tau_np, host_z)
            barrier()
            t0 = time.perf_counter()
            for _ in range(ll_steps):
                r_ll = h.loglik_z(covType, covparms, nug_np, tau_np, host_z)
            barrier()
            extras["loglik_e2e_evals_per_s"] = ll_steps / (time.perf_counter() - t0)
            extras["loglik_e2e_value_rank0_shard"] = r_ll["loglik"] if world == 1 else None
        if args.workload == "cfg2":
            per = {}
            rng_ = float(covparms[1])
            for tag, ct, cp in (("nu0.5", "matern", [SIG2, rng_, 0.5]), ("nu2.5", "matern", [SIG2, rng_, 2.5]),
                                ("general_nu0.8", "matern", [SIG2, rng_, 0.8]), ("general_nu1.3", "matern", [SIG2, rng_, 1.3]),
                                ("esqe", "esqe", [1.0, rng_, 0.5, rng_])):
                cpa = np.array(cp)

                def fn(ct=ct, cpa=cpa):
                    h.u_dev(ct, cpa, d_nug.data_ptr(), d_out.data_ptr(), packed=False, stream=stream)
                ms = timed(fn, ll_steps, 3)
                per[tag] = n_sets * ll_steps / (ms * 1e-3)
            extras["sets_per_s_other_covariances"] = per

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    F = flops_per_set(p, d, wl["tag"])
    B = bytes_per_set(p, d)
    peak_tf = C.c_double(0)
    G._lib.check(G.lib.gpv_measure_fp64_peak(local_rank, C.byref(peak_tf)))
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    nfull_rank = pb["nfull_rank"]
    achieved_tf = F * nfull_rank / (k_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "kernel": kname, "achieved": achieved_tf, "peak": peak_tf.value, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf.value,
        "peak_source": "fp64 FMA peak from a DFMA micro-kernel in this run (gpv_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has HBM and bf16 only",
        "flops_per_set": F, "sets_per_launch": nfull_rank, "kernel_ms": k_ms, "kernel_launches_timed": k_count,
        "kernel_share_of_step": k_ms / (ms_total / args.steps),
        "hbm": {"achieved": B * nfull_rank / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": B * nfull_rank / (k_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_set": B, "peak_source": hbm_src},
        "traffic": None,
    }
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof) and args.workload == "cfg2":
        try:
            pj = json.load(open(prof))
            roofline["traffic"] = pj.get("dram_bytes_per_launch")
            roofline["fp64_pipe_active_pct_ncu"] = pj.get("fp64_pipe_pct_of_peak_active")
            if pj.get("note"):
                roofline["ncu_note"] = pj["note"]
        except Exception:
            pass

    cpu = None
    if getattr(bind_to_gpu_cpus, "unbound", None):
        os.sched_setaffinity(0, bind_to_gpu_cpus.unbound)
    if rank == 0 and not args.no_cpu_baseline:
        full = (revNN != 0).sum(axis=1) >= 2          # time the full sets only (the unit of `value`)
        rate, cores, sample = cpu_reference_rate(locs, revNN[full], revCond[full], rb, nuggets, covparms, covType=covType)
        cpu = {"value": rate, "unit": "sets/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "conditioning sets/sec (createU)", "value": value, "unit": "sets/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "sets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "gpv_u_values_packed (createU's U_NZentries + packing, pinned host buffers)"},
            "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
            "input_generation_s": t_gen,
        }
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
