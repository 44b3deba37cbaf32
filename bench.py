#!/usr/bin/env python
"""bench.py -- conditioning sets/sec of createU (U_NZentries) and loglik evals/sec on B200, per the round contract.

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N ...            CPU arm: the reference's own sources compiled here
                                                           (oracle/_ref; the restatement if that is missing)

Primary workload (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] -- n = 1e6 uniform 2-D
locations per GPU (weak scaling: n = N * 1e6, rows sharded by contiguous equal ranges, every rank holds all
locations), m = 30, Matern nu = 1.5 closed form, standard Vecchia ('z') conditioning, per-location nuggets.  One
step = one U_NZentries pass over the rank's rows.  `value` times the device-resident call (gpv_u_dev); `e2e` times
the reference-facing call with host buffers (gpv_u_values_packed: nuggets H2D, packed U values D2H) -- the call
createU() makes.

North-star workload (`roofline.cfg3`, `e2e.cfg3`): BASELINE.json configs[2] -- n = 1e7 total (STRONG scaling),
m = 30, general-nu Matern (nu = 0.8, the Bessel branch), sharded over the N ranks; createU sets/s, the fused
likelihood (evals/s, one all-reduce of the partial sums), roofline fraction and, inside the run, parity: every rank
compares sampled rows of its shard with the oracle, and the all-reduced log-likelihood is compared with the value
a single GPU gives (tests/golden/bench_loglik.json).
"""
import argparse
import ctypes as C
import importlib.util
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
GOLD_LL = os.path.join(ROOT, "tests", "golden", "bench_loglik.json")


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


def load_synth():
    """The pure numpy / scipy input generators, loaded BY PATH: importing the package would map the product's
    .so, which the reference arm must not do."""
    spec = importlib.util.spec_from_file_location("gpv_synth", os.path.join(ROOT, "gpvecchia_b200", "_synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_text(name, wl, n_total):
    cov_desc = f"Matern nu={wl['nu']}" if wl["covType"] == "matern" else "esqe"
    return (f"{name}: createU/U_NZentries, n={n_total} uniform {wl['d']}-D locs"
            f"{' (' + str(wl['n']) + '/GPU)' if wl['scaling'] == 'weak' else ''}"
            f"{' + ' + str(wl['n_pred']) + ' prediction locs' if wl['n_pred'] else ''}, m={wl['m']}, {cov_desc}, "
            f"'{wl['layout']}' conditioning")


def build_config(name, wl, n_total, world, north):
    """`config` of the JSON line: the same for both arms (the driver compares them)."""
    p = wl["m"] + 1
    per_row_mb = (p * 4 + p * 8) / 1e6
    cfg = dict(workload=workload_text(name, wl, n_total), n=n_total, m=wl["m"], d=wl["d"], covmodel=wl["covType"],
               nu=wl["nu"], cond_yz=wl["layout"],
               sharding=(f"rows by contiguous equal ranges over {world} rank(s) (sum n0^3 balanced for layouts with "
                         "trivial rows); locs and nuggets replicated, ids / conditioning masks / outputs sliced"),
               l2=(f"touched per step: {per_row_mb * n_total / world:.0f} MB of ids + U values per rank "
                   + ("(larger than the 126 MB L2)" if per_row_mb * n_total / world > 126 else
                      "(fits the 126 MB L2: this workload is launch-latency bound, not a bandwidth case)")))
    if north is not None:
        cfg["north_star_workload"] = workload_text("cfg3", north, north["n"]) + f", strong scaling over {world} rank(s)"
    return cfg


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
        # 'z' conditioning: neighbours on the response, self on the latent (vecchia_specify.R:189-190).  Plain equal
        # row ranges: with the library's locality layer a late row costs what an early one does.
        cuts = shard.uniform_cuts(n_total, world)
        rb, re_ = int(cuts[rank]), int(cuts[rank + 1])
        if use_gpu_nn:
            revNN = H.ordered_nn_gpu(locs_obs, m, rb, re_, device=device)
        else:
            revNN = H.rev(H.ordered_nn_kdtree(locs_obs, m, rb, re_)).astype(np.int32)
        revCond = np.zeros(revNN.shape, dtype=np.int32)
        revCond[revNN == 0] = np.iinfo(np.int32).min
        revCond[:, -1] = 1
        return dict(locs=locs_obs, revNN=revNN, revCond=revCond, obs=np.ones(n_total, dtype=np.int32),
                    nug_all=tau, nug_obs=tau, z=z, covparms=covparms, rb=rb, re=re_, N=n_total,
                    skip_rows=0, nfull_total=n_total - 1, nfull_rank=int(((revNN != 0).sum(axis=1) >= 2).sum()),
                    obs_lo=rb, obs_hi=re_)
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
    ocuts = shard.uniform_cuts(n_total, world)    # the observations' Z entries: any disjoint split will do
    return dict(locs=locs2, revNN=revNN, revCond=revCond, obs=obs.astype(np.int32), nug_all=nug_all,
                nug_obs=tau, z=z, covparms=covparms, rb=rb, re=re_, N=N, skip_rows=n_total,
                nfull_total=int((n0 >= 2).sum()), nfull_rank=int((n0[rb:re_] >= 2).sum()),
                obs_lo=int(ocuts[rank]), obs_hi=int(ocuts[rank + 1]))


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so
    omp_get_max_threads() would understate the box; the CPU arms take the team size as an argument
    (num_threads(Ncores), like src/U_NZentries.cpp:37)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_cpus(cuda_index):
    """N > 1 only: run this rank on the CPUs NVML reports as local to its GPU before any pinned host buffer is
    allocated, so that the ranks' device-to-host copies land in the memory of the socket their GPU hangs on.
    GPV_BENCH_NUMA=0 turns it off.  Returns a description."""
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


# ---------------------------------------------------------------------------------------------------------------
# CPU arms.  The compiled reference (oracle/_ref: the reference's own U_NZentries.cpp / Matern.cpp / Esqe.cpp /
# dist.cpp, unmodified) processes every row of the problem it is given, so its sample is the FIRST k rows of the
# workload -- a self-contained problem, because ordered-NN rows only name earlier rows -- with the workload's
# covparms.  The restatement (oracle/) has a row-range entry point and is timed on the LAST rows as well.
# ---------------------------------------------------------------------------------------------------------------
def reference_prefix_problem(S, wl, n_total, k):
    """First k rows (full conditioning sets from row m on) of the n_total-location workload, marshalled for the
    compiled reference."""
    from oracle import ref_native as RN
    d, m = wl["d"], wl["m"]
    rng_ = S.default_range(n_total, d)
    covparms = np.array([SIG2, rng_, wl["nu"]]) if wl["covType"] == "matern" else np.array([1.0, rng_, 0.5, rng_])
    locs = S.make_locs(n_total, d, stream=2)[:k]
    tau = S.make_nuggets(n_total, stream=2)[:k]
    NN = S.ordered_nn_kdtree(locs, m)
    if wl["layout"] == "z":
        revNN = S.rev(NN).astype(np.int32)
        revCond = np.zeros(revNN.shape, dtype=np.int32)
        revCond[revNN == 0] = np.iinfo(np.int32).min
        revCond[:, -1] = 1
        return RN.Problem(k, locs, revNN, revCond, tau, tau, wl["covType"], covparms), k - 1
    locs2, NN2, Cond, obs = S.layout_zy(locs, m, k)          # zy: k dummy rows + k full rows
    revCond = S.rev(Cond).astype(np.int32)
    revCond[revCond < 0] = np.iinfo(np.int32).min
    nug_all = np.concatenate([tau, np.zeros(k)])
    return RN.Problem(k, locs2, S.rev(NN2).astype(np.int32), revCond, nug_all, tau, wl["covType"], covparms), k


def time_reference(S, wl, n_total, threads, target_s, backends=("openblas", "textbook")):
    """(sets/s, backend, k, passes, seconds) of the compiled reference on about target_s seconds of CPU work: a pilot
    picks the faster of LAPACK (OpenBLAS from scipy) and the published unblocked chol / back substitution."""
    from oracle import ref_native as RN
    k0 = min(4000, n_total)
    pr, nfull = reference_prefix_problem(S, wl, n_total, k0)
    rates = {}
    for b in backends:
        RN.force_textbook(b == "textbook")
        pr.run(threads)
        t0 = time.perf_counter()
        pr.run(threads)
        rates[b] = nfull / (time.perf_counter() - t0)
    best = max(rates, key=rates.get)
    RN.force_textbook(best == "textbook")
    k = int(min(n_total, 200_000, max(k0, rates[best] * min(target_s, 4.0))))
    pr, nfull = reference_prefix_problem(S, wl, n_total, k)
    pr.run(threads)                                      # page-fault / thread-pool warm-up
    passes = max(1, int(round(target_s / max(nfull / rates[best], 1e-3))))
    per = []
    for _ in range(passes):
        t0 = time.perf_counter()
        pr.run(threads)
        per.append(time.perf_counter() - t0)
    RN.force_textbook(False)
    # the median pass: a pass that shared the cores with something else does not set the number
    return nfull / float(np.median(per)), best, k, passes, float(np.sum(per)), rates


def port_rate(locs, revNN_rows, revCond_rows, row_begin, nuggets, covparms, covType, threads, target_s, mode=0):
    """The restatement (oracle/, OpenMP schedule(static) + LAPACK, no per-row heap traffic) on the LAST rows of the
    workload; (sets/s, description)."""
    import oracle as O
    nr = revNN_rows.shape[0]

    def timed(nrows_s):
        pr = O.RowsProblem(locs, revNN_rows[nr - nrows_s:], revCond_rows[nr - nrows_s:], row_begin + nr - nrows_s,
                           nuggets, covType, covparms)
        t0 = time.perf_counter()
        pr.run(threads, mode=mode)
        return time.perf_counter() - t0
    pilot = min(20000, nr)
    timed(min(2000, nr))
    t_p = timed(pilot)
    nrows_s = int(min(nr, max(pilot, pilot / t_p * target_s)))
    t_s = timed(nrows_s)
    return nrows_s / t_s, f"last {nrows_s} rows of the n={locs.shape[0]} workload, {t_s:.1f} s"


def cpu_baseline_block(wl, n_total, pb, full_mask):
    """`cpu_baseline` of the own arm (rank 0, N = 1): the compiled reference on all host cores (the headline CPU
    number), on one core, a core-scaling table, and the restatement beside it."""
    S = load_synth()
    threads = host_threads()
    out = {}
    try:
        from oracle import ref_native as RN
        have_ref = RN.available()
    except Exception:
        have_ref = False
    if have_ref:
        _, backend, _, _, _, _ = time_reference(S, wl, n_total, threads, target_s=1.0)
        by = {}
        for c in sorted({1, 8, 16, 32, threads}):
            if c <= threads:
                by[str(c)] = time_reference(S, wl, n_total, c, target_s=2.0, backends=(backend,))[0]
        rate, backend, k, passes, dt, rates = time_reference(S, wl, n_total, threads, target_s=10.0)
        out.update(value=rate, unit="sets/s", cores=threads, kind="reference",
                   sample=(f"the reference's own U_NZentries.cpp (oracle/_ref) on the first {k} rows of the n={n_total} "
                           f"workload (self-contained: ordered-NN rows only name earlier rows), same covparms, median of {passes} "
                           f"passes ({dt:.1f} s in all); chol/solve backend {backend} (pilot: "
                           + ", ".join(f"{b} {v:.3g}" for b, v in rates.items()) + " sets/s)"))
        out["by_cores"] = by
        out["one_core"] = by.get("1")
    locs, revNN, revCond = pb["locs"], pb["revNN"][full_mask], pb["revCond"][full_mask]
    pr, ps = port_rate(locs, revNN, revCond, pb["rb"], pb["nug_all"], pb["covparms"], wl["covType"], threads, 4.0)
    p1, _ = port_rate(locs, revNN, revCond, pb["rb"], pb["nug_all"], pb["covparms"], wl["covType"], 1, 2.0)
    pt, _ = port_rate(locs, revNN, revCond, pb["rb"], pb["nug_all"], pb["covparms"], wl["covType"], threads, 3.0, mode=1)
    out["port"] = {"value": max(pr, pt), "lapack": pr, "textbook": pt, "one_core": p1, "cores": threads,
                   "sample": ps, "what": "oracle/ restatement (same arithmetic as the reference, bit for bit; no "
                                         "Armadillo-style temporaries), the faster of LAPACK and textbook chol"}
    if not have_ref:
        out.update(value=out["port"]["value"], unit="sets/s", cores=threads, kind="port", sample=ps)
    return out


def reference_arm(args, name, wl, n_total, world, north):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    S = load_synth()
    threads = host_threads()
    config = build_config(name, wl, n_total, world, north)
    try:
        from oracle import ref_native as RN
        have_ref = RN.available()
    except Exception:
        have_ref = False
    if have_ref:
        from oracle import ref_native as RN
        rate0, backend, _, _, _, rates = time_reference(S, wl, n_total, threads, target_s=1.5)
        RN.force_textbook(backend == "textbook")
        k = int(min(n_total, 200_000, max(4000, rate0 * 3.0)))        # ~3 s of CPU work per step
        pr, nfull = reference_prefix_problem(S, wl, n_total, k)
        run = lambda: pr.run(threads)                        # noqa: E731
        kind = "reference"
        sample = (f"{nfull} full sets per step: the reference's own U_NZentries.cpp (oracle/_ref) on the first {k} rows of "
                  f"the n={n_total} workload (self-contained: ordered-NN rows only name earlier rows), same covparms, "
                  f"{threads} OpenMP threads, chol/solve backend {backend}")
    else:
        import oracle as O
        k = min(60_000, n_total)
        locs = S.make_locs(n_total, wl["d"], stream=2)[:k]
        tau = S.make_nuggets(n_total, stream=2)[:k]
        revNN = S.rev(S.ordered_nn_kdtree(locs, wl["m"])).astype(np.int32)
        revCond = np.zeros(revNN.shape, dtype=np.int32)
        revCond[revNN == 0] = np.iinfo(np.int32).min
        revCond[:, -1] = 1
        rng_ = S.default_range(n_total, wl["d"])
        cp = np.array([SIG2, rng_, wl["nu"]]) if wl["covType"] == "matern" else np.array([1.0, rng_, 0.5, rng_])
        po = O.RowsProblem(locs, revNN, revCond, 0, tau, wl["covType"], cp)
        run = lambda: po.run(threads)                        # noqa: E731
        nfull, kind = k - 1, "port"
        sample = f"{nfull} full sets per step: oracle/ restatement on the first {k} rows, {threads} OpenMP threads"
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = nfull * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "conditioning sets/sec (createU)", "value": value, "unit": "sets/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "sets/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank's GPU: device-resident steps, end-to-end steps, likelihood, parity sample."""

    def __init__(self, name, wl, n_total, world, rank, local_rank, host_nn=False):
        import torch
        import gpvecchia_b200 as G
        self.torch, self.G = torch, G
        self.name, self.wl, self.n_total, self.world, self.rank, self.local_rank = name, wl, n_total, world, rank, local_rank
        self.dev = torch.device("cuda", local_rank)
        t0 = time.perf_counter()
        self.pb = pb = make_inputs(wl, n_total, world, rank, local_rank, use_gpu_nn=not host_nn)
        self.t_gen = time.perf_counter() - t0
        self.p = wl["m"] + 1
        self.nrows = pb["re"] - pb["rb"]
        t0 = time.perf_counter()
        self.h = G.UHandle(pb["locs"], pb["revNN"], pb["revCond"], obs=pb["obs"], row_begin=pb["rb"], row_end=pb["re"],
                           device=local_rank)
        self.t_create = time.perf_counter() - t0
        self.d_nug = torch.from_numpy(pb["nug_all"]).to(self.dev)
        self.d_out = torch.empty(self.nrows * self.p, dtype=torch.float64, device=self.dev)
        self.d_z = torch.from_numpy(pb["z"]).to(self.dev)
        self.d_ll = torch.zeros(8, dtype=torch.float64, device=self.dev)
        self.tstream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """W untimed steps, then exactly `steps` between two CUDA events on the launching stream, a barrier +
        synchronize on both sides, MAX over ranks (ms)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        self.h.kernel_time_stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.allmax(e0.elapsed_time(e1))

    def step_dev(self, covType=None, covparms=None):
        pb = self.pb
        self.h.u_dev(covType or self.wl["covType"], pb["covparms"] if covparms is None else covparms,
                     self.d_nug.data_ptr(), self.d_out.data_ptr(), packed=False, stream=self.stream)

    def step_ll(self):
        from gpvecchia_b200 import shard
        pb = self.pb
        self.h.u_dev(self.wl["covType"], pb["covparms"], self.d_nug.data_ptr(), None, d_zord=self.d_z.data_ptr(),
                     skip_rows=pb["skip_rows"], d_loglik=self.d_ll.data_ptr(), stream=self.stream)
        if self.world > 1:
            shard.allreduce_loglik(self.d_ll)          # a handful of doubles over NCCL: the path's only collective

    def loglik_value(self):
        """Whole log-likelihood of a pure-z layout from the all-reduced partial sums of the last step_ll()."""
        pb = self.pb
        parts = [float(v) for v in self.d_ll.cpu().tolist()]
        z, tau = pb["z"], pb["nug_obs"]
        qn = parts[0] + float(np.sum(z * z / tau))
        ldn = parts[1] + float(np.sum(np.log(tau)))
        out = dict(quadform_num=qn, logdet_num=ldn, quadform_denom=parts[3], logdet_denom=parts[4], nfail=parts[2])
        if self.wl["layout"] == "z":
            out["loglik"] = -0.5 * (ldn - parts[4] + qn - parts[3] + self.n_total * float(np.log(2 * np.pi)))
        return out

    def parity_sample(self, nsample=2000):
        """Max row-scaled error of sampled rows of THIS rank's shard (device-resident output of the last step_dev)
        against the oracle restatement on the same rows; MAX over ranks.  North_star bar: 1e-10."""
        import oracle as O
        pb = self.pb
        full = np.nonzero((pb["revNN"] != 0).sum(axis=1) >= 2)[0]
        rs = np.random.default_rng(1234 + self.rank)
        rows = np.sort(rs.choice(full, size=min(nsample, full.size), replace=False))
        self.torch.cuda.synchronize()
        idx = self.torch.from_numpy(rows).to(self.dev)
        got = self.d_out.view(self.nrows, self.p).index_select(0, idx).cpu().numpy()
        pr = O.RowsProblem(pb["locs"], pb["revNN"][rows], pb["revCond"][rows], 0, pb["nug_all"], self.wl["covType"],
                           pb["covparms"])
        pr.run(host_threads())
        ref = pr.Lentries().copy()
        err = float((np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max())
        # structural pattern: n0 values, then exact zeros (U_NZentries.cpp:33,63).  (Not `got == 0` against
        # `ref == 0`: in the first rows of a large problem neighbours are hundreds of ranges apart and the
        # covariances underflow -- exp gives 0 on the host and 1e-304 on the device, both "zero" at scale 1.)
        n0 = (pb["revNN"][rows] != 0).sum(axis=1)
        cols = np.arange(self.p)[None, :]
        same_pattern = bool(np.all(got[cols >= n0[:, None]] == 0) and np.all(ref[cols >= n0[:, None]] == 0)
                            and np.all(got[np.arange(rows.size), n0 - 1] > 0))
        self.parity_extra = None
        if self.wl["layout"] == "zy":
            # latent neighbours without nugget plus a duplicated location: cond ~ 1e5-1e6, two correct fp64 results
            # differ by cond * eps.  The arbiter is the __float128 mode of the oracle on the same fp64 inputs.
            pr.run(host_threads(), mode=2)
            quad = pr.Lentries().copy()
            sc = np.abs(quad).max(axis=1, keepdims=True)
            self.parity_extra = {"gpu_vs_float128": self.allmax(float((np.abs(got - quad) / sc).max())),
                                 "fp64_oracle_vs_float128": self.allmax(float((np.abs(ref - quad) / sc).max()))}
        return self.allmax(err if same_pattern else 1.0), rows.size

    def e2e(self, steps):
        """The reference-facing call with HOST buffers: gpv_u_values_packed (createU's U_NZentries + packing).  Every
        rank uploads the per-location nuggets and the nuggets of ITS slice of the observations, and downloads its
        rows' packed U values and its slice of Zentries; wall clock around `steps` calls, barrier on both sides,
        MAX over ranks.  Also times plain pinned D2H copies of the same size: the box's ceiling for this call."""
        torch, pb, h = self.torch, self.pb, self.h
        n_obs_slice = pb["obs_hi"] - pb["obs_lo"]
        total = h.packed_len + 2 * n_obs_slice
        host_out = torch.empty(total, dtype=torch.float64).pin_memory()
        host_nug = torch.from_numpy(pb["nug_all"]).pin_memory()
        host_tau = torch.from_numpy(np.ascontiguousarray(pb["nug_obs"][pb["obs_lo"]:pb["obs_hi"]])).pin_memory()
        out_np, nug_np, tau_np = host_out.numpy(), host_nug.numpy(), host_tau.numpy()
        self.e2e_bufs = (out_np, nug_np, tau_np)

        def step():
            h.values_packed(self.wl["covType"], pb["covparms"], nug_np, tau_np, zentries_tail=True, out=out_np)
        for _ in range(2):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.barrier()
        dt = self.allmax(time.perf_counter() - t0)
        h2d = 8 * (h.nuggets_read + n_obs_slice)     # what the call copies: the nuggets its rows name + its tau slice
        d2h = 8 * total
        # ceiling: the same number of bytes as plain pinned copies, all ranks at once
        d_src = torch.empty(total, dtype=torch.float64, device=self.dev)
        for _ in range(2):
            host_out.copy_(d_src, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            host_out.copy_(d_src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self.barrier()
        dt_copy = self.allmax(time.perf_counter() - t0)
        del d_src
        return dict(seconds_per_step=dt / steps, h2d=h2d, d2h=d2h, copy_seconds_per_step=dt_copy / steps)

    def close(self):
        self.h.close()
        self.torch.cuda.set_stream(self.torch.cuda.default_stream(self.dev))


def one_process_section(wl, n_total, ndev, steps):
    """The form an R session uses on a multi-GPU box: ONE process, gpv_multi_* (a worker thread per device inside the
    library), host buffers in and out.  Runs on rank 0 while the other ranks wait; returns end-to-end sets/s of
    gpv_multi_u_values_packed, the whole-likelihood rate of gpv_multi_loglik_z and a parity sample vs the oracle."""
    import torch
    import gpvecchia_b200 as G
    import oracle as O
    pb = make_inputs(wl, n_total, 1, 0, 0)
    p = wl["m"] + 1
    with G.MultiHandle(pb["locs"], pb["revNN"], pb["revCond"], obs=pb["obs"], devices=list(range(ndev))) as mh:
        n_obs = pb["nug_obs"].size
        host_out = torch.empty(mh.packed_len + 2 * n_obs, dtype=torch.float64).pin_memory().numpy()
        nug = torch.from_numpy(pb["nug_all"]).pin_memory().numpy()
        tau = torch.from_numpy(pb["nug_obs"]).pin_memory().numpy()
        z = torch.from_numpy(pb["z"]).pin_memory().numpy()
        for _ in range(2):
            mh.values_packed(wl["covType"], pb["covparms"], nug, tau, zentries_tail=True, out=host_out)
        t0 = time.perf_counter()
        for _ in range(steps):
            mh.values_packed(wl["covType"], pb["covparms"], nug, tau, zentries_tail=True, out=host_out)
        dt = (time.perf_counter() - t0) / steps
        res = {"value": pb["nfull_total"] / dt, "unit": "sets/s", "devices": ndev, "seconds_per_step": dt,
               "d2h_bytes_per_step": 8 * host_out.size, "achieved_gbs": 8 * host_out.size / dt / 1e9,
               "call": "gpv_multi_u_values_packed: one process, one worker thread per device, pinned host buffers"}
        # parity: sampled rows of the packed vector against the oracle
        n0 = (pb["revNN"] != 0).sum(axis=1)
        off = np.concatenate([[0], np.cumsum(n0)])
        full = np.nonzero(n0 == p)[0]
        rows = np.sort(np.random.default_rng(99).choice(full, size=min(2000, full.size), replace=False))
        got = np.stack([host_out[off[r]:off[r] + p] for r in rows])
        pr = O.RowsProblem(pb["locs"], pb["revNN"][rows], pb["revCond"][rows], 0, pb["nug_all"], wl["covType"], pb["covparms"])
        pr.run(host_threads())
        ref = pr.Lentries()
        res["parity_max_err"] = float((np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max())
        if wl["layout"] == "z":
            for _ in range(2):
                r = mh.loglik_z(wl["covType"], pb["covparms"], nug, tau, z)
            t0 = time.perf_counter()
            for _ in range(steps):
                r = mh.loglik_z(wl["covType"], pb["covparms"], nug, tau, z)
            res["loglik_evals_per_s"] = steps / (time.perf_counter() - t0)
            res["loglik"] = r["loglik"]
    return res


def sum_over_ranks(v, runner):
    t = runner.torch.tensor([float(v)], dtype=runner.torch.float64, device=runner.dev)
    if runner.world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def roofline_block(runner, k_ms, k_count, ms_step, peak_tf):
    wl, pb = runner.wl, runner.pb
    F = flops_per_set(runner.p, wl["d"], wl["tag"])
    B = bytes_per_set(runner.p, wl["d"])
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    nfull_rank = pb["nfull_rank"]
    achieved_tf = F * nfull_rank / (k_ms * 1e-3) / 1e12
    kname = runner.h.last_kernel_name()
    rl = {
        "bound": "fp64", "kernel": kname, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf,
        "peak_source": "fp64 FMA peak from a DFMA micro-kernel in this run (gpv_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has HBM and bf16 only",
        "flops_per_set": F, "sets_per_launch": nfull_rank, "kernel_ms": k_ms, "kernel_launches_timed": k_count,
        "kernel_share_of_step": k_ms / ms_step,
        "hbm": {"achieved": B * nfull_rank / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": B * nfull_rank / (k_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_set": B, "peak_source": hbm_src},
        "traffic": None,
    }
    # evidence from the committed ncu capture of THIS kernel (profiles/roofline_traffic.json, keyed by kernel name):
    # DRAM bytes per launch, executed fp64 warp instructions per set
    try:
        pj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("kernels", {}).get(kname)
    except Exception:
        pj = None
    if pj:
        if pj.get("sets_per_launch") == nfull_rank or abs(pj.get("sets_per_launch", 0) - nfull_rank) <= 1:
            rl["traffic"] = pj.get("dram_bytes_per_launch")
        ipset = pj.get("fp64_warp_inst_per_set")
        if ipset:
            # what the fp64 pipe physically did: executed fp64 warp instructions x 32 lanes x 2 flop / time / peak
            rl["fp64_inst_frac"] = ipset * nfull_rank * 64.0 / (k_ms * 1e-3) / 1e12 / peak_tf
            rl["max_algorithmic_frac_at_full_pipe"] = F / (ipset * 64.0)
            rl["fp64_warp_inst_per_set"] = ipset
        rl["ncu_source"] = pj.get("source")
    return rl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the workload's n (per GPU if weak, total if strong)")
    ap.add_argument("--north-n", type=int, default=0, help="override n of the north-star (cfg3) section")
    ap.add_argument("--host-nn", action="store_true", help="build neighbour arrays with cKDTree on the host")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other covariances and the dgCMatrix@x call")
    ap.add_argument("--no-north-star", action="store_true", help="skip the cfg3 (n = 1e7, general nu) section")
    ap.add_argument("--no-one-process", action="store_true", help="N > 1: skip the one-process gpv_multi_* section")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    wl = dict(WORKLOADS[name])
    if args.n:
        wl["n"] = args.n
    n_total = wl["n"] * world if wl["scaling"] == "weak" else wl["n"]
    north = None
    if name == "cfg2" and not args.no_north_star:
        north = dict(WORKLOADS["cfg3"])
        if args.north_n:
            north["n"] = args.north_n

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, name, wl, n_total, world, north)
        return

    import torch
    import torch.distributed as dist
    import gpvecchia_b200 as G

    if G.lib.gpv_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the gpvecchia_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    host_binding = None
    if world > 1:
        host_binding = bind_to_gpu_cpus(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    config = build_config(name, wl, n_total, world, north)

    peak_tf = C.c_double(0)
    G._lib.check(G.lib.gpv_measure_fp64_peak(local_rank, C.byref(peak_tf)))
    peak_tf = peak_tf.value

    # ---- primary workload ---------------------------------------------------------------------------------------
    R = Runner(name, wl, n_total, world, rank, local_rank, host_nn=args.host_nn)
    n_sets = R.pb["nfull_total"]            # conditioning sets with n0 >= 2 over all ranks: the unit of `value`
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = G.lib.gpv_launch_count()
    ms_total = R.timed(R.step_dev, args.steps, args.warmup)
    launches = int(G.lib.gpv_launch_count() - launches0)
    launches = launches * args.steps // (args.steps + args.warmup)      # timed region only
    clocks = sampler.stop()
    value = n_sets * args.steps / (ms_total * 1e-3)
    k_count, k_total = R.h.kernel_time_stats(reset=True)
    k_ms = k_total / max(k_count, 1)
    roofline = roofline_block(R, k_ms, k_count, ms_total / args.steps, peak_tf)
    perr, nsamp = R.parity_sample()
    roofline["parity_max_err"] = perr
    roofline["parity"] = f"max row-scaled |U - oracle| over {nsamp} sampled rows per rank, MAX over ranks (bar 1e-10)"
    if R.parity_extra:
        roofline["parity_ill_conditioned_layout"] = R.parity_extra

    e2e_steps = max(3, min(args.steps, 10))
    em = R.e2e(e2e_steps)
    e2e_value = n_sets / em["seconds_per_step"]
    d2h_all = sum_over_ranks(em["d2h"], R)
    e2e = {"value": e2e_value, "unit": "sets/s", "h2d_bytes_per_step": em["h2d"], "d2h_bytes_per_step": em["d2h"],
           "call": "gpv_u_values_packed (createU's U_NZentries + packing, pinned host buffers; per rank: the per-location "
                   "nuggets its rows name up (a prefix for a row shard; staged with the output chunks), its rows' U values "
                   "and its slice of Zentries down)",
           "bytes_per_step_all_ranks": {"h2d": sum_over_ranks(em["h2d"], R), "d2h": d2h_all,
                                        "what": "h2d/d2h_bytes_per_step above are rank 0's; a later rank of an ordered "
                                                "layout uploads a longer prefix of the nuggets"},
           "achieved_gbs": d2h_all / em["seconds_per_step"] / 1e9,
           "ceiling_gbs": d2h_all / em["copy_seconds_per_step"] / 1e9,
           "frac": em["copy_seconds_per_step"] / em["seconds_per_step"],
           "ceiling": "the same bytes as plain pinned device-to-host copies, all ranks at once, measured in this run"}
    if host_binding:
        e2e["host_binding"] = host_binding

    # likelihood: fused numerator (+ denominator terms for the pure-z layout), scalars out, one all-reduce
    ll_steps = max(5, args.steps // 2)
    ms_ll = R.timed(R.step_ll, ll_steps, 3)
    llv = R.loglik_value()
    e2e["loglik_evals_per_s"] = ll_steps / (ms_ll * 1e-3)
    e2e["loglik"] = llv.get("loglik")
    e2e["loglik_what"] = ("whole vecchia_likelihood (numerator + per-row denominator terms, pure-z layout), data resident "
                          "in HBM, partial sums all-reduced; device-timed like `value`")
    if wl["layout"] == "z":
        out_np, nug_np, tau_np = R.e2e_bufs
        host_z = torch.from_numpy(R.pb["z"]).pin_memory().numpy()
        full_tau = torch.from_numpy(R.pb["nug_obs"]).pin_memory().numpy()
        for _ in range(2):
            R.h.loglik_z(wl["covType"], R.pb["covparms"], nug_np, full_tau, host_z)
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(ll_steps):
            R.h.loglik_z(wl["covType"], R.pb["covparms"], nug_np, full_tau, host_z)
        R.barrier()
        e2e["loglik_e2e_evals_per_s"] = ll_steps / R.allmax(time.perf_counter() - t0)
        R.h.loglik_z(wl["covType"], R.pb["covparms"], nug_np, full_tau, host_z)
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(ll_steps):
            R.h.loglik_z(wl["covType"], R.pb["covparms"], None, None, None)
        R.barrier()
        e2e["loglik_e2e_resident_evals_per_s"] = ll_steps / R.allmax(time.perf_counter() - t0)
        e2e["loglik_e2e_what"] = ("gpv_loglik_z per rank with host buffers (ALL nuggets, tau, z up on every rank; 6 doubles "
                                  "back), and its estimation-loop form (data resident on the handle, only covparms go up)")
        if world > 1:
            # the multi-process form of the same call: every rank uploads only ITS slice, the slices are exchanged over
            # NVLink (grouped NCCL broadcasts inside the library), the partial sums are all-reduced (gpv_loglik_z_dist)
            try:
                from gpvecchia_b200 import shard
                uid = torch.from_numpy(G.UHandle.dist_unique_id().copy()).to(R.dev) if rank == 0 else \
                    torch.zeros(128, dtype=torch.uint8, device=R.dev)
                dist.broadcast(uid, src=0)
                R.h.dist_init(uid.cpu().numpy(), rank, world)
                cuts = shard.uniform_cuts(n_total, world)
                a, b = int(cuts[rank]), int(cuts[rank + 1])
                nug_s = torch.from_numpy(np.ascontiguousarray(R.pb["nug_all"][a:b])).pin_memory().numpy()
                tau_s = torch.from_numpy(np.ascontiguousarray(R.pb["nug_obs"][a:b])).pin_memory().numpy()
                z_s = torch.from_numpy(np.ascontiguousarray(R.pb["z"][a:b])).pin_memory().numpy()
                for _ in range(2):
                    rd = R.h.loglik_z_dist(wl["covType"], R.pb["covparms"], nug_s, tau_s, z_s, cuts, cuts)
                R.barrier()
                t0 = time.perf_counter()
                for _ in range(ll_steps):
                    rd = R.h.loglik_z_dist(wl["covType"], R.pb["covparms"], nug_s, tau_s, z_s, cuts, cuts)
                R.barrier()
                e2e["loglik_e2e_nvlink_evals_per_s"] = ll_steps / R.allmax(time.perf_counter() - t0)
                e2e["loglik_e2e_nvlink_rel_err"] = abs(rd["loglik"] - llv["loglik"]) / abs(llv["loglik"])
                e2e["loglik_e2e_nvlink_what"] = ("gpv_loglik_z_dist: each rank uploads 1/N of nuggets, tau and z from pinned "
                                                 "host memory, NCCL broadcasts over NVLink fill the rest, one ncclAllReduce "
                                                 "of the partial sums; every rank returns the whole log-likelihood")
            except G.GpvError as ex:
                e2e["loglik_e2e_nvlink_evals_per_s"] = None
                e2e["loglik_e2e_nvlink_what"] = f"unavailable: {ex}" 

    extras = {"input_generation_s": R.t_gen, "handle_creation_s": R.t_create}
    if not args.no_extras:
        try:
            ncols, nnz_csc, _ = R.h.csc_dims()
            host_csc = torch.empty(nnz_csc, dtype=torch.float64).pin_memory().numpy()
            nug_np = R.e2e_bufs[1]
            full_tau = R.pb["nug_obs"]
            for _ in range(2):
                R.h.values_csc(wl["covType"], R.pb["covparms"], nug_np, full_tau, out=host_csc)
            R.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                R.h.values_csc(wl["covType"], R.pb["covparms"], nug_np, full_tau, out=host_csc)
            R.barrier()
            e2e["csc_sets_per_s"] = n_sets * e2e_steps / R.allmax(time.perf_counter() - t0)
            del host_csc
        except G.GpvError:                      # duplicate U rows in a set: triplet route only
            e2e["csc_sets_per_s"] = None
        if world == 1:
            # what an R session gets: every R vector is PAGEABLE and freshly allocated (Rf_allocVector: untouched pages)
            pg = {}
            tau_pg = R.pb["nug_obs"]
            tot = R.h.packed_len + 2 * tau_pg.size
            for tag, mk in (("fresh", lambda: np.empty(tot)), ("reused", lambda b=np.empty(tot): b)):
                for _ in range(2):
                    R.h.values_packed(wl["covType"], R.pb["covparms"], R.pb["nug_all"], tau_pg, out=mk())
                t0 = time.perf_counter()
                for _ in range(5):
                    R.h.values_packed(wl["covType"], R.pb["covparms"], R.pb["nug_all"], tau_pg, out=mk())
                pg[tag + "_sets_per_s"] = n_sets * 5 / (time.perf_counter() - t0)
            pg["what"] = ("gpv_u_values_packed into pageable numpy arrays, allocated per call (`fresh`: first-touch page faults "
                          "included, like a new R vector) or reused; worker threads of the library stage the pieces through "
                          "page-locked slots (a plain cudaMemcpy into fresh pageable memory ran at 1.7e7 sets/s)")
            e2e["pageable"] = pg
        if name == "cfg2":
            per = {}
            rng_ = float(R.pb["covparms"][1])
            for tag, ct, cp in (("nu0.5", "matern", [SIG2, rng_, 0.5]), ("nu2.5", "matern", [SIG2, rng_, 2.5]),
                                ("general_nu0.8", "matern", [SIG2, rng_, 0.8]), ("general_nu1.3", "matern", [SIG2, rng_, 1.3]),
                                ("esqe", "esqe", [1.0, rng_, 0.5, rng_])):
                cpa = np.array(cp)
                ms = R.timed(lambda ct=ct, cpa=cpa: R.step_dev(ct, cpa), ll_steps, 3)
                per[tag] = n_sets * ll_steps / (ms * 1e-3)
            roofline["sets_per_s_other_covariances"] = per

    cpu = None
    if getattr(bind_to_gpu_cpus, "unbound", None):
        os.sched_setaffinity(0, bind_to_gpu_cpus.unbound)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        full = (R.pb["revNN"] != 0).sum(axis=1) >= 2
        cpu = cpu_baseline_block(wl, n_total, R.pb, full)
    R.close()
    del R
    torch.cuda.empty_cache()

    # ---- north star: cfg3, n = 1e7, general nu, strong scaling -------------------------------------------------
    if north is not None:
        if getattr(bind_to_gpu_cpus, "unbound", None) and world > 1:
            bind_to_gpu_cpus(local_rank)
        N3 = Runner("cfg3", north, north["n"], world, rank, local_rank, host_nn=args.host_nn)
        sets3 = N3.pb["nfull_total"]
        steps3 = max(5, min(args.steps, 10))
        ms3 = N3.timed(N3.step_dev, steps3, 3)
        kc3, kt3 = N3.h.kernel_time_stats(reset=True)
        r3 = roofline_block(N3, kt3 / max(kc3, 1), kc3, ms3 / steps3, peak_tf)
        r3.update(workload=config["north_star_workload"], n=north["n"], scaling="strong", value=sets3 * steps3 / (ms3 * 1e-3),
                  value_unit="sets/s", ms_per_step=ms3 / steps3, steps=steps3)
        perr3, nsamp3 = N3.parity_sample()
        r3["parity_max_err"] = perr3
        r3["parity"] = f"max row-scaled |U - oracle| over {nsamp3} sampled rows per rank, MAX over ranks (bar 1e-10)"
        ms3l = N3.timed(N3.step_ll, steps3, 3)
        ll3 = N3.loglik_value()
        r3["loglik_evals_per_s"] = steps3 / (ms3l * 1e-3)
        r3["loglik"] = ll3.get("loglik")
        r3["loglik_nfail"] = ll3.get("nfail")
        try:
            gold = json.load(open(GOLD_LL)).get(f"cfg3_n{north['n']}_nu{north['nu']}")
        except Exception:
            gold = None
        if gold is not None and ll3.get("loglik") is not None:
            r3["loglik_rel_err_vs_n1"] = abs(ll3["loglik"] - gold) / abs(gold)     # bar 1e-8
        em3 = N3.e2e(3)
        d2h3 = sum_over_ranks(em3["d2h"], N3)
        e2e["cfg3"] = {"value": sets3 / em3["seconds_per_step"], "unit": "sets/s", "h2d_bytes_per_step": em3["h2d"],
                       "d2h_bytes_per_step": em3["d2h"], "achieved_gbs": d2h3 / em3["seconds_per_step"] / 1e9,
                       "ceiling_gbs": d2h3 / em3["copy_seconds_per_step"] / 1e9,
                       "frac": em3["copy_seconds_per_step"] / em3["seconds_per_step"],
                       "loglik_evals_per_s": r3["loglik_evals_per_s"]}
        roofline["cfg3"] = r3
        extras["cfg3_input_generation_s"] = N3.t_gen
        extras["cfg3_handle_creation_s"] = N3.t_create
        N3.close()

    # ---- one process, all GPUs (the form an R session uses): rank 0 drives every device, the others wait ----------
    if world > 1 and not args.no_one_process:
        torch.cuda.empty_cache()
        dist.barrier()
        if rank == 0:
            if getattr(bind_to_gpu_cpus, "unbound", None):
                os.sched_setaffinity(0, bind_to_gpu_cpus.unbound)
            try:
                e2e["one_process"] = one_process_section(wl, n_total, world, max(3, min(args.steps, 5)))
            except Exception as ex:                                   # reported, not fatal: the torchrun numbers stand
                e2e["one_process"] = {"error": f"{type(ex).__name__}: {ex}"}
        dist.barrier()

    if rank == 0:
        line = {
            "metric": "conditioning sets/sec (createU)", "value": value, "unit": "sets/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "gpu_launches": launches,
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
