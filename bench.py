#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the bootstrap particle filter hot path on B200.

One "step" = one llFilter (model/ParticleFilter.scala:137-140) over T observations with N
particles on each GPU: init + T x (propagate, weight, log-sum-exp, resample).  Metric =
particle-steps/sec = N * T / time, summed over the GPUs of the job (each rank filters its own
independent cloud: the reference's only parallelism is across independent filters / PMMH chains,
so scaling is weak and there is no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|c2|c1|c4|c5|c3] [--impl reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, all host cores, independent filters in parallel) on a bounded sample of the same workload.
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
    # name: (description, model builder name, particles, observations, resampler, dt)
    "target": ("composed Poisson + seasonal(24,3) + OU, 2^24 particles x 1000 observations, systematic, fp32 "
               "(BASELINE.json target)", "c2", 1 << 24, 1000, "systematic"),
    "c2": ("composed Poisson + seasonal(24,3) + OU, 2^20 particles x 1000 observations, systematic, fp32 "
           "(BASELINE.json configs[1])", "c2", 1 << 20, 1000, "systematic"),
    "c1": ("Poisson + OU, 1000 particles x 500 observations, systematic (BASELINE.json configs[0])", "c1", 1000, 500,
           "systematic"),
    "c4": ("negative binomial + linear trend, 2^16 particles x 500 observations, systematic (configs[3] likelihood)", "c4",
           1 << 16, 500, "systematic"),
    "c5": ("Normal + seasonal(24,3) + OU, 2^24 particles per GPU x 100 observations, systematic (configs[4] per-rank shard)",
           "c5", 1 << 24, 100, "systematic"),
    "c3": ("LGCP + Brownian motion, 2^22 particles x 200 events, stratified, precision 3 (configs[2])", "c3", 1 << 22, 200,
           "stratified"),
}


def build_model(name):
    from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters
    import composablestatespacemodels_b200 as cs
    ou1 = SdeParameter.ouParameter([1.0], [0.5], [0.2], [1.5], [0.05])   # examples/Simulation.scala:16
    ou6 = SdeParameter.ouParameter([0.1], [1.0], [0.4], [0.1], [0.5])    # examples/Simulation.scala:64-67
    if name == "c1":
        return Model.poisson(Sde.ouProcess(1))(Parameters(None, ou1))
    if name == "c2":
        return (Model.poisson(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6)))(
            Parameters(None, ou1) | Parameters(None, ou6))
    if name == "c4":
        return (Model.negativeBinomial(Sde.brownianMotion(1)) | Model.linear(Sde.genBrownianMotion(1)))(
            Parameters(2.0, SdeParameter.brownianParameter([0.0], [1.0], [0.01])) |
            Parameters(None, SdeParameter.genBrownianParameter([0.0], [1.0], [0.01], [0.01])))
    if name == "c5":
        return (Model.linear(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6)))(
            Parameters(0.0, ou1) | Parameters(None, ou6))
    if name == "c3":
        m = Model.lgcp(Sde.brownianMotion(1))(Parameters(None, SdeParameter.brownianParameter([0.0], [1.0], [0.01])))
        return cs.model.Model(m.leaves, m.step_mode, 3)
    raise SystemExit(f"unknown model {name}")


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


def cpu_baseline(wl_model, resampler, T_full, budget_s, threads):
    """The CPU restatement of the reference (oracle/, kind "port") on a bounded sample."""
    import oracle
    from composablestatespacemodels_b200.resampling import Resampling
    mod = build_model(wl_model)
    orc = oracle.Oracle(mod)
    kind = Resampling.kind_of(resampler)
    T = min(T_full, 100 if wl_model != "c3" else 10)
    t, y = synth_series(mod, wl_model, T)
    # calibrate on a small cloud, then size the sample for ~budget_s seconds of CPU work
    n0 = 2048
    t0 = time.perf_counter()
    orc.filter_ll(n0, kind, t, y, seed=1, variant=1)
    rate0 = n0 * T / (time.perf_counter() - t0)
    n = int(min(1 << 18, max(4096, rate0 * budget_s / T)))
    t0 = time.perf_counter()
    orc.filter_ll_many(n, kind, t, y, seed=2, variant=1, R=threads, threads=threads)
    el = time.perf_counter() - t0
    flat = threads * n * T / el
    # the reference-faithful cost model (per-particle objects, TreeMap ECDF) on a smaller sample
    nf = max(1024, n // 8)
    t0 = time.perf_counter()
    orc.filter_ll_many(nf, kind, t, y, seed=3, variant=0, R=threads, threads=threads)
    faithful = threads * nf * T / (time.perf_counter() - t0)
    return {"value": flat, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"{threads} independent filter(s), {n} particles x {T} observations each, flat-array C++ restatement "
                      f"(oracle/, fp64); reference-faithful variant (per-particle objects + std::map ECDF, {nf} particles): "
                      f"{faithful:.3e} particle-steps/s",
            "faithful_value": faithful}


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm (oracle port; the Scala original cannot run:
    no JVM in this image) on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, wl_model, N, T, resampler = wl
    cores = os.cpu_count() or 1
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(wl_model, resampler, T, budget_s=max(2.0, 20.0 / max(1, args.steps)), threads=cores)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    out = {"metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": desc, "note": "CPU restatement of the reference (not the JVM), bounded sample"},
           "cpu_baseline": last,
           "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--obs", type=int, default=0, help="override number of observations")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.particles:
        wl[2] = args.particles
    if args.obs:
        wl[3] = args.obs
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    import composablestatespacemodels_b200 as cs
    from composablestatespacemodels_b200 import _abi
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
    mod = build_model(wl_model)
    t, y = synth_series(mod, wl_model, T)
    dtype = _abi.F32 if args.dtype == "f32" else _abi.F64
    b = 4 if dtype == _abi.F32 else 8
    d = mod.dimension
    h = cs.GpuFilterHandle(mod, Resampling.kind_of(resampler), N, dtype=dtype, device=local, seed=2, stream_id=rank)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    h.load_series(t, y)  # inputs resident in HBM before the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        h.ll_resident()
    # ---- timed region: K steps, device events on the launching stream, max over ranks ----------
    sampler = ClockSampler(local)
    h.profile(10)  # per-kernel CUDA events on every 10th observation (roofline of the dominant kernel)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    lls = []
    for _ in range(args.steps):
        lls.append(h.ll_resident())
        launches += h.last_launches()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    prof = h.profile_read()
    h.profile(0)
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * N * T * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: host observations in, log-likelihood out ----------
    from composablestatespacemodels_b200 import Filter, Data
    data = [Data(tt, yy) for tt, yy in zip(t, y)]
    flt = Filter(mod, Resampling.systematicResampling if resampler == "systematic" else Resampling.stratifiedResampling,
                 dtype=dtype, device=local, seed=3, stream_id=rank) if wl_model != "c3" else None
    h.close()
    if flt is not None:
        flt.llFilter(data[: max(2, T // 50)], N)  # allocate the cloud once (not timed), as a long-lived filter would
        barrier()
        w0 = time.perf_counter()
        k_e2e = max(1, min(args.steps, 3))
        for _ in range(k_e2e):
            flt.llFilter(data, N)
        torch.cuda.synchronize()
        el = time.perf_counter() - w0
        if world > 1:
            tel = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(tel, op=dist.ReduceOp.MAX)
            el = float(tel.item())
        e2e_v = world * N * T * k_e2e / el
        flt.close()
    else:
        e2e_v = None
    e2e = {"value": e2e_v, "unit": "particle-steps/s", "h2d_bytes_per_step": int(T * 17), "d2h_bytes_per_step": 8,
           "note": "Filter.llFilter(data, n): host observations (t, y, has_obs) in, per-observation constants built on the "
                   "host and passed as kernel arguments, one fp64 log-likelihood out; wall clock"}

    if rank == 0:
        peak, peak_src = measured_peak()
        k1_ms, k1_n = prof["propagate_weight"]
        k1_bytes = (2 * d * b + b + 4) * N  # anc + gathered state in, state + log-weight out
        roof = None
        if k1_n:
            ach = k1_bytes / (k1_ms / k1_n * 1e-3) / 1e9
            step_bytes = (4 * d * b + 5 * b + 8) * N * T * args.steps  # SURVEY.md 8(d) B_step accounting
            roof = {"bound": "hbm", "kernel": "k_propagate_weight (gather + propagate + weight, fused)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": 2 * d * b + b + 4,
                    "avg_launch_ms": k1_ms / k1_n, "sampled_launches": k1_n,
                    "step_frac_survey_accounting": step_bytes / (ms * 1e-3) / 1e9 / peak,
                    "kernel_share_of_step": {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items()}}
        cpu = None
        if not args.no_cpu:
            cpu = cpu_baseline(wl_model, resampler, T, budget_s=12.0, threads=1)
        out = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": desc, "particles_per_gpu": N, "observations": T, "latent_dim": d, "resampler": resampler,
                          "l2": "working set %.0f MB per GPU %s the 126 MB L2, no flush" % (
                              (2 * d * b + b + 4) * N / 1e6, "exceeds" if (2 * d * b + b + 4) * N > 126e6 else "is below"),
                          "parallelism": f"{world} independent filter(s), one per GPU, no collective"},
               "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "log_likelihood_mean": float(np.mean(lls))}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
