#!/usr/bin/env python
"""bench.py -- the bootstrap particle filter hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|c2|c1|c3|c4|c5] [--impl reference]

Workloads:
  target  Poisson + seasonal(24,3) + OU, 2^24 particles x 1000 observations, systematic   (default)
  c2      the same model, 2^20 particles                                  (BASELINE.json configs[1])
  c1      Poisson + OU, 1000 particles x 500 observations                 (configs[0], the CPU-sized case)
  c3      LGCP + Brownian motion, 2^22 particles x 200 events, stratified, 10^-3 sub-steps (configs[2])
  c4      PMMH, negative binomial + linear trend, 2^16 particles x 500 observations per likelihood
          evaluation, one chain per GPU; metric = PMMH iterations/sec     (configs[3])
  c5      ONE filter of 2^27 particles, Normal + seasonal + OU, 100 observations, sharded over the
          GPUs of the job (in-kernel NVLink exchange); strong scaling      (configs[4])

One "step" = one llFilter (model/ParticleFilter.scala:137-140) over the T observations -- for c4 one
PMMH iteration (model/PMMH.scala:68-81).  For target/c1/c2/c3/c4 every rank runs its own independent
filter / chain (the reference's only parallelism: Streaming.pilotRun, PMMH chains), so scaling is
weak and there is no data-path collective; c5 shards one cloud and the ranks exchange inside the
kernels.  Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the
reference (oracle/, all host cores) on a bounded sample of the same workload.
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

WORKLOADS = {
    # name: (description, model builder name, particles, observations, resampler)
    "target": ("composed Poisson + seasonal(24,3) + OU, 2^24 particles x 1000 observations, systematic, fp32 "
               "(BASELINE.json target)", "c2", 1 << 24, 1000, "systematic"),
    "c2": ("composed Poisson + seasonal(24,3) + OU, 2^20 particles x 1000 observations, systematic, fp32 "
           "(BASELINE.json configs[1])", "c2", 1 << 20, 1000, "systematic"),
    "c1": ("Poisson + OU, 1000 particles x 500 observations, systematic (BASELINE.json configs[0])", "c1", 1000, 500,
           "systematic"),
    "c4": ("PMMH, negative binomial + linear trend, 2^16 particles x 500 observations per likelihood evaluation, "
           "one chain per GPU (BASELINE.json configs[3])", "c4", 1 << 16, 500, "systematic"),
    "c5": ("ONE filter, Normal + seasonal(24,3) + OU, 2^27 particles x 100 observations, systematic, sharded over the "
           "GPUs of the job (BASELINE.json configs[4])", "c5", 1 << 27, 100, "systematic"),
    "c3": ("LGCP + Brownian motion, 2^22 particles x 200 events, stratified, sub-step 10^-3 (BASELINE.json configs[2])",
           "c3", 1 << 22, 200, "stratified"),
}


def build_unparam(name):
    """(unparameterised model, parameters) with the reference's example values."""
    from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters
    ou1 = SdeParameter.ouParameter([1.0], [0.5], [0.2], [1.5], [0.05])   # examples/Simulation.scala:16
    ou6 = SdeParameter.ouParameter([0.1], [1.0], [0.4], [0.1], [0.5])    # examples/Simulation.scala:64-67
    if name == "c1":
        return Model.poisson(Sde.ouProcess(1)), Parameters(None, ou1)
    if name == "c2":
        return (Model.poisson(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6)),
                Parameters(None, ou1) | Parameters(None, ou6))
    if name == "c4":
        return (Model.negativeBinomial(Sde.brownianMotion(1)) | Model.linear(Sde.genBrownianMotion(1)),
                Parameters(2.0, SdeParameter.brownianParameter([0.0], [1.0], [0.01])) |
                Parameters(None, SdeParameter.genBrownianParameter([0.0], [1.0], [0.01], [0.01])))
    if name == "c5":
        return (Model.linear(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6)),
                Parameters(0.0, ou1) | Parameters(None, ou6))
    if name == "c3":
        return Model.lgcp(Sde.brownianMotion(1)), Parameters(None, SdeParameter.brownianParameter([0.0], [1.0], [0.01]))
    raise SystemExit(f"unknown model {name}")


def build_model(name):
    import composablestatespacemodels_b200 as cs
    um, p = build_unparam(name)
    m = um(p)
    if name == "c3":
        return cs.model.Model(m.leaves, m.step_mode, 3)  # FilterLgcp precision 3: sub-step 10^-3
    return m


def synth_series(mod, wl_model, T):
    from composablestatespacemodels_b200 import simulate
    if wl_model == "c3":
        return simulate.simLgcpEvents(T, 0.1, seed=1)
    t, y, _ = simulate.simRegular(mod, 0.1, T, seed=1)
    return t, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(wl_model, resampler, T_full, budget_s, threads, pmmh=False):
    """The CPU restatement of the reference (oracle/, kind "port") on a bounded sample.
    Particle-steps/s of llFilter; for PMMH the iteration rate that follows from it (an iteration IS
    one llFilter of N x T particle-steps plus O(#parameters) host work, model/PMMH.scala:68-81)."""
    import oracle
    from composablestatespacemodels_b200.resampling import Resampling
    mod = build_model(wl_model)
    orc = oracle.Oracle(mod)
    kind = Resampling.kind_of(resampler)
    T = min(T_full, 100 if wl_model != "c3" else 10)
    t, y = synth_series(mod, wl_model, T)
    # calibrate on a small cloud, then size the sample for ~budget_s seconds of CPU work
    n0 = 2048 if wl_model != "c3" else 256
    t0 = time.perf_counter()
    orc.filter_ll(n0, kind, t, y, seed=1, variant=1)
    rate0 = n0 * T / (time.perf_counter() - t0)
    n = int(min(1 << 18, max(1024, rate0 * budget_s / T)))
    t0 = time.perf_counter()
    orc.filter_ll_many(n, kind, t, y, seed=2, variant=1, R=threads, threads=threads)
    el = time.perf_counter() - t0
    flat = threads * n * T / el
    # the reference-faithful cost model (per-particle objects, TreeMap ECDF) on a smaller sample
    nf = max(512, n // 8)
    t0 = time.perf_counter()
    orc.filter_ll_many(nf, kind, t, y, seed=3, variant=0, R=threads, threads=threads)
    faithful = threads * nf * T / (time.perf_counter() - t0)
    out = {"value": flat, "unit": "particle-steps/s", "cores": threads, "kind": "port",
           "sample": f"{threads} independent filter(s), {n} particles x {T} observations each, flat-array C++ restatement "
                     f"(oracle/, fp64); reference-faithful variant (per-particle objects + std::map ECDF, {nf} particles): "
                     f"{faithful:.3e} particle-steps/s",
           "faithful_value": faithful}
    if pmmh:
        N, Tf = WORKLOADS["c4"][2], WORKLOADS["c4"][3]
        out["particle_steps_per_s"] = flat
        out["value"] = flat / (N * Tf)
        out["unit"] = "pmmh-iterations/s"
        out["sample"] += f"; iterations/s = particle-steps/s / ({N} x {Tf}), {threads} chain(s), one per core"
    return out


def run_reference(args, wl, wl_name):
    """--impl reference: the reference's CPU algorithm (oracle port; the Scala original cannot run:
    no JVM in this image) on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, wl_model, N, T, resampler = wl
    cores = os.cpu_count() or 1
    pmmh = wl_name == "c4"
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        budget = float(os.environ.get("CSSM_BENCH_BUDGET_S", max(2.0, 20.0 / max(1, args.steps))))  # tests shrink it
        last = cpu_baseline(wl_model, resampler, T, budget_s=budget, threads=cores, pmmh=pmmh)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    metric, unit = ("pmmh-iterations/sec", "pmmh-iterations/s") if pmmh else ("particle-steps/sec", "particle-steps/s")
    out = {"metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
           "scaling": "strong" if wl_name == "c5" else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": desc, "note": "CPU restatement of the reference (not the JVM), bounded sample"},
           "cpu_baseline": last,
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default 5; 100 PMMH iterations for c4)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles (per GPU; for c5 the global count)")
    ap.add_argument("--obs", type=int, default=0, help="override number of observations")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tie-first", action="store_true",
                    help="informational: CSSM_TIE_FIRST (textbook inverse CDF) instead of the reference's TreeMap rule")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.particles:
        wl[2] = args.particles
    if args.obs:
        wl[3] = args.obs
    if not args.steps:
        args.steps = 100 if args.workload == "c4" else 5
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)

    import torch
    import torch.distributed as dist
    import composablestatespacemodels_b200 as cs
    from composablestatespacemodels_b200 import _abi, sharding
    from composablestatespacemodels_b200.resampling import Resampling

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    desc, wl_model, N, T, resampler = wl
    sharded = args.workload == "c5"
    pmmh = args.workload == "c4"
    mod = build_model(wl_model)
    t, y = synth_series(mod, wl_model, T)
    dtype = _abi.F32 if args.dtype == "f32" else _abi.F64
    b = 4 if dtype == _abi.F32 else 8
    d = mod.dimension
    kind = Resampling.kind_of(resampler)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            tv = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            return float(tv.item())
        return v

    if sharded:
        # ONE cloud of N particles over `world` GPUs; rank r owns the slots [r*N/world, (r+1)*N/world)
        if world > 1:
            h = sharding.create_sharded(mod, kind, N, dtype=dtype, device=local, seed=2)
        else:
            h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, device=local, seed=2)
        n_local, n_total = N // world, N
    else:
        h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, device=local, seed=2, stream_id=rank)
        n_local, n_total = N, N * world
    h.set_stream(stream.cuda_stream)
    if args.tie_first:
        h.set_tie_rule(_abi.TIE_FIRST)
    h.load_series(t, y)  # inputs resident in HBM before the timed region

    sampler = ClockSampler(local)
    launches = 0
    lls = []
    extra = {}
    if pmmh:
        # ---- PMMH: K iterations of the chain; every iteration = new parameters + one llFilter ----------
        from composablestatespacemodels_b200 import MetropolisHastings, GpuBootstrapFilter, Data, perturb
        h.close()
        um, p0 = build_unparam(wl_model)
        data = [Data(tt, yy) for tt, yy in zip(t, y)]
        rng = np.random.default_rng(100 + rank)
        pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, N, dtype=dtype, device=local, seed=5,
                                stream_id=rank)
        pf.handle.set_stream(stream.cuda_stream)
        # examples/DetermineParameters.scala:59,73: Parameters.perturb(0.05), flat prior, symmetric proposal
        mh = MetropolisHastings(p0, perturb(0.05, rng), lambda a, c: 0.0, lambda p: 0.0, pf, rng)
        it = mh.iters()
        for _ in range(args.warmup):
            next(it)
        pf.handle.profile(1)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        acc0 = None
        for _ in range(args.steps):
            s = next(it)
            lls.append(s.ll)
            launches += pf.handle.last_launches() + 1
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        barrier()
        clocks = sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
        wall = max_over_ranks(wall)
        prof = pf.handle.profile_read()
        value = world * args.steps / (ms * 1e-3)
        e2e_v = world * args.steps / wall
        extra = {"accepted": int(s.accepted), "particle_steps_per_s": value * N * T,
                 "device_ms_per_likelihood": pf.handle.last_elapsed_ms()}
        e2e = {"value": e2e_v, "unit": "pmmh-iterations/s", "h2d_bytes_per_step": int(T * 17), "d2h_bytes_per_step": 8 + 8 * d,
               "note": "MetropolisHastings(...).iters() with a GpuBootstrapFilter: per iteration the proposed parameters go "
                       "in (per-observation constants rebuilt on the host), one log-likelihood and one sampled state come "
                       "out; wall clock"}
        pf.close()
    else:
        for _ in range(args.warmup):
            h.ll_resident()
        # ---- timed region: K steps, device events on the launching stream, max over ranks ----------
        h.profile(10)  # per-kernel CUDA events on every 10th observation (roofline of the dominant kernel)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            lls.append(h.ll_resident())
            launches += h.last_launches()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
        prof = h.profile_read()
        h.profile(0)
        value = n_total * T * args.steps / (ms * 1e-3)
        if not sharded and wl_model != "c3":
            # the step after the filter in the streaming examples (examples/Filtering.scala:29): getIntervals of the
            # final cloud on the device (mean + 2(d+1) order statistics by radix select), not part of the timed region
            h.intervals(float(t[-1]))
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for _ in range(3):
                h.intervals(float(t[-1]))
            extra["get_intervals_ms"] = (time.perf_counter() - w0) / 3 * 1e3

        # ---- end to end through the public API: host observations in, log-likelihood out ----------
        from composablestatespacemodels_b200 import Filter, FilterLgcp, Data
        data = [Data(tt, yy) for tt, yy in zip(t, y)]
        k_e2e = max(1, min(args.steps, 3))
        if sharded:
            # the sharded handle IS the public object (there is no reference class for it): host series in, ll out
            h.ll_arrays(t, y)
            barrier()
            w0 = time.perf_counter()
            for _ in range(k_e2e):
                h.ll_arrays(t, y)
            torch.cuda.synchronize()
            el = max_over_ranks(time.perf_counter() - w0)
            h.close()
        else:
            h.close()
            rs = Resampling.systematicResampling if resampler == "systematic" else Resampling.stratifiedResampling
            if wl_model == "c3":
                flt = FilterLgcp(mod, rs, 3, dtype=dtype, device=local, seed=3, stream_id=rank)
            else:
                flt = Filter(mod, rs, dtype=dtype, device=local, seed=3, stream_id=rank)
            flt.llFilter(data[: max(2, T // 50)], N)  # allocate the cloud once (not timed), as a long-lived filter would
            barrier()
            w0 = time.perf_counter()
            for _ in range(k_e2e):
                flt.llFilter(data, N)
            torch.cuda.synchronize()
            el = max_over_ranks(time.perf_counter() - w0)
            flt.close()
        e2e_v = n_total * T * k_e2e / el
        e2e = {"value": e2e_v, "unit": "particle-steps/s", "h2d_bytes_per_step": int(T * 17), "d2h_bytes_per_step": 8,
               "note": "Filter.llFilter(data, n): host observations (t, y, has_obs) in, per-observation constants built on the "
                       "host and passed as kernel arguments, one fp64 log-likelihood out; wall clock"}

    if rank == 0:
        peak, peak_src = measured_peak()
        k1_ms, k1_n = prof["propagate_weight"]
        k1_bpp = 2 * d * b + b + 4  # anc + gathered state in, state + log-weight out
        ws_bytes_small = k1_bpp * n_local
        roof = None
        if k1_n and wl_model != "c3":
            ach = k1_bpp * n_local / (k1_ms / k1_n * 1e-3) / 1e9
            per_launch = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items()}
            tot = sum(per_launch.values())
            # what one observed step moves with the gather fused into the next propagate and no CDF written:
            # K1 2db+b+4, K2 b, K3 b+4
            real_bpp = k1_bpp + b + b + 4
            traffic = None
            try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture (same shape only)
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
                if tr and tr["particles"] == n_local and args.dtype == "f32":
                    traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": "k_propagate_weight (gather + propagate + weight, fused)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": k1_bpp * n_local,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": k1_bpp,
                    "avg_launch_ms": k1_ms / k1_n, "sampled_launches": k1_n,
                    "kernel_ms_per_launch": per_launch,
                    # scan / search / gather throughput of the resampling stage (north_star): particles per second through
                    # K2 (exp + exact sums) and K3 (CDF scan + ancestor search); the gather is fused into K1's load
                    "resampling": {"weight_sums_particles_per_s": n_local / (per_launch["weight_sums"] * 1e-3) if per_launch["weight_sums"] else None,
                                   "scan_search_particles_per_s": n_local / (per_launch["scan_search"] * 1e-3) if per_launch["scan_search"] else None,
                                   "bytes_per_particle": {"weight_sums": b, "scan_search": b + 4}},
                    "whole_step": {"bytes_per_particle_fused": real_bpp,
                                   "achieved_gbs_fused": real_bpp * n_local / (tot * 1e-3) / 1e9 if tot else None,
                                   "frac_fused": real_bpp * n_local / (tot * 1e-3) / 1e9 / peak if tot else None,
                                   "bytes_per_particle_survey_8d": 4 * d * b + 5 * b + 8,
                                   "frac_survey_8d": (4 * d * b + 5 * b + 8) * n_local / (tot * 1e-3) / 1e9 / peak if tot else None}}
        elif prof.get("series", (0, 0))[1]:
            # small cloud: the whole llFilter is ONE cooperative launch (cssm_series.cuh); per launch it moves
            # T x N x (K1 2db+b+4, K2 b, K3 b+4) bytes, all of it L2 resident: latency bound, not HBM bound
            s_ms, s_n = prof["series"]
            real_bpp = k1_bpp + b + b + 4
            ach = real_bpp * n_local * T / (s_ms / s_n * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_series_small (whole llFilter in one cooperative launch: propagate + weight, "
                                              "exact sums, scan + search per observation, grid barriers in between)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": real_bpp,
                    "avg_launch_ms": s_ms / s_n, "sampled_launches": s_n,
                    "us_per_observation": s_ms / s_n * 1e3 / T,
                    "note": "latency bound: the 2^16-particle cloud (%.1f MB) lives in L2; three grid barriers per observation"
                            % (ws_bytes_small / 1e6)}
        elif k1_n:
            # LGCP: the state lives in registers for ~100 sub-steps per event: SFU/ALU bound, not HBM bound
            nsub = float(np.sum(np.ceil(np.diff(np.concatenate([[t[0]], t])) / 1e-3)))
            roof = {"bound": "hbm", "kernel": "k_lgcp_weight (sub-stepped propagate + hazard, fused)",
                    "achieved": k1_bpp * n_local / (k1_ms / k1_n * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": k1_bpp * n_local / (k1_ms / k1_n * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "note": "instruction bound: Philox + Box-Muller + exp per sub-step, state in registers",
                    "particle_substeps_per_s": n_total * nsub * args.steps / (ms * 1e-3),
                    "avg_launch_ms": k1_ms / k1_n, "sampled_launches": k1_n}
        cpu = None
        if not args.no_cpu:
            cpu = cpu_baseline(wl_model, resampler, T, budget_s=12.0, threads=1, pmmh=pmmh)
        metric, unit = ("pmmh-iterations/sec", "pmmh-iterations/s") if pmmh else ("particle-steps/sec", "particle-steps/s")
        if sharded:
            par = (f"one filter of {n_total} particles sharded over {world} GPU(s), {n_local} per GPU; per step three in-kernel "
                   "exchanges (max, sums, done) + ancestor scatter / parent gather over NVLink peer pointers" if world > 1
                   else "one filter on one GPU (the unsharded case of the strong-scaling series)")
        else:
            par = f"{world} independent {'chain' if pmmh else 'filter'}(s), one per GPU, no collective"
        ws_bytes = k1_bpp * n_local
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if sharded else "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": desc, "particles_per_gpu": n_local, "particles_total": n_total, "observations": T,
                          "latent_dim": d, "resampler": resampler,
                          "tie_rule": "first index (NOT the reference's rule)" if args.tie_first else "reference (TreeMap: last particle of a repeated key)",
                          "l2": "working set %.0f MB per GPU %s the 126 MB L2, no flush" % (
                              ws_bytes / 1e6, "exceeds" if ws_bytes > 126e6 else "is below"),
                          "parallelism": par},
               "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "log_likelihood_mean": float(np.mean(lls))}
        out.update(extra)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
