#!/usr/bin/env python
"""bench.py -- the bootstrap particle filter hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|c2|c1|c3|c4|c5|resample] [--impl reference]

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

The default line (workload `target`) also carries BASELINE.json's other two numbers as sub-records, each with its
own clocks: `pmmh` (c4: PMMH iterations/s, one chain per GPU, and the rate with several chains sharing a GPU) and,
with --gpus N > 1, `sharded` (c5: ONE 2^27-particle filter over the N GPUs, strong scaling against its own 1-GPU
run, per-kernel times, exposed exchange time per observation).  --no-extra skips them.
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
    "resample": ("isolated resampling through the Resample[A] seam (cssm_resample, host vectors in and out): the sizes and unit "
                 "weights of the reference's ResamplingBenchmark (src/bench/scala/Resampling.scala:11-17) and 2^20 / 2^24 "
                 "weights of SURVEY 8(d): unit, exp(N(0, 2^2)), degenerate", "resample", 1 << 24, 1, "systematic"),
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
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML sampled in-process every 20 ms
    (starts at once, so even a 100 ms region of an 8-GPU job gets samples); `nvidia-smi -lms` when NVML is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.th, self.stop_flag = gpu_index, [], None, None, False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _nvml_loop(self):
        nv = self.nv
        R = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            smax = None
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((sm, smax, [n for n, b in R.items() if mask & b]))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        self.rows, self.stop_flag = [], False
        if self.nv is not None:
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]), [n for n, v in zip(names, r[4:8]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                pass

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            if self.th:
                self.th.join(timeout=1)
            src = "nvml"
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            src = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"], "samples": 0}
        sm = [r[0] for r in self.rows]
        smax = next((r[1] for r in self.rows if r[1]), None)
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": reasons, "samples": len(sm),
                "source": src}


def kernels_blob():
    """git blob hash of the two files the step kernels live in: what a committed ncu traffic figure is valid for."""
    import hashlib
    h = hashlib.sha1()
    for f in ("cssm_kernels.cuh", "cssm_common.cuh"):
        data = open(os.path.join(ROOT, "composablestatespacemodels_b200", "csrc", f), "rb").read()
        h.update(hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest().encode())
    return h.hexdigest()[:16]


def run_resample(args, wl):
    """The reference's ResamplingBenchmark (src/bench/scala/Resampling.scala: sizes 100 .. 6400, unit weights, systematic and
    multinomial) and the large-cloud cases of SURVEY 8(d), through the stand-alone seam S2: `Resample[A]` = host weights in,
    host ancestors out (cssm_resample: allocation, upload, max + exact sums + scan + search, download -- everything is
    inside the timed call).  The CPU restatement (oracle.resample, one core) is timed on the same inputs.  `--impl
    reference` times the CPU side only.  One JSON line; `value` = resampled particles/s, systematic, exp(N(0,4)) weights,
    largest size."""
    import oracle
    from composablestatespacemodels_b200 import _abi
    gpu = args.impl != "reference"
    if gpu:
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        from composablestatespacemodels_b200.resampling import ancestors
    rng = np.random.default_rng(3)
    kinds = (("systematic", _abi.RESAMPLE_SYSTEMATIC), ("stratified", _abi.RESAMPLE_STRATIFIED), ("multinomial", _abi.RESAMPLE_MULTINOMIAL))
    big = wl[2]
    sizes = [100, 200, 400, 800, 1600, 3200, 6400, 1 << 20] + ([big] if big > (1 << 20) else [])

    def weights(which, n):
        if which == "unit":
            return np.ones(n)
        if which == "lognormal":
            return np.exp(rng.normal(0.0, 2.0, n))
        w = np.full(n, 1e-30)
        w[n // 3] = 1.0
        return w

    cases, head = [], None
    for n in sizes:
        for which in (("unit",) if n <= 6400 else ("unit", "lognormal", "degenerate")):
            w = weights(which, n)
            for kname, kind in kinds:
                if kname == "multinomial" and n > (1 << 20):
                    continue  # the reference's multinomial is O(n^2); the device's is a binary search per draw, the oracle's too slow here
                u = rng.random(1 if kname == "systematic" else n)
                rec = {"n": n, "weights": which, "kind": kname}
                if gpu:
                    for _ in range(max(1, args.warmup if n <= (1 << 20) else 1)):
                        a = ancestors(kind, w, u)
                    reps = 20 if n <= 6400 else (5 if n <= (1 << 20) else max(2, args.steps // 2))
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        a = ancestors(kind, w, u)
                    dt = (time.perf_counter() - t0) / reps
                    rec.update({"gpu_particles_per_s": n / dt, "gpu_ms": dt * 1e3})
                # (the CPU restatement walks runs of repeated keys output by output: quadratic on the degenerate case)
                if not args.no_cpu and n <= (1 << 20) and (which != "degenerate" or n <= 6400):
                    t0 = time.perf_counter()
                    ac = oracle.resample(kind, w, u)
                    dtc = time.perf_counter() - t0
                    rec.update({"cpu_particles_per_s": n / dtc, "cpu_ms": dtc * 1e3})
                    if gpu:
                        rec["same_ancestors"] = bool(np.array_equal(a, ac))
                cases.append(rec)
                if kname == "systematic" and which == "lognormal":
                    head = rec
    key = "gpu_particles_per_s" if gpu else "cpu_particles_per_s"
    if head is None or key not in head:
        head = next(c for c in reversed(cases) if key in c)
    cpu_head = next((c for c in reversed(cases) if "cpu_particles_per_s" in c and c["kind"] == "systematic" and c["weights"] == "lognormal"), None)
    n = head["n"]
    peak, peak_src = measured_peak()
    line = {"metric": "resampled-particles/sec", "value": head[key], "unit": "particles/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head.get("gpu_ms" if gpu else "cpu_ms"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl[0], "headline_case": {k: head[k] for k in ("n", "weights", "kind")}},
            "roofline": {"bound": "hbm", "achieved": (8 + 8 + 4) * n / (head["gpu_ms"] * 1e-3) / 1e9 if gpu else None, "peak": peak, "unit": "GB/s",
                         "frac": ((8 + 8 + 4) * n / (head["gpu_ms"] * 1e-3) / 1e9 / peak) if gpu else None, "traffic": None, "peak_source": peak_src,
                         "note": "end to end through host vectors (PCIe copies, allocation): not a kernel figure; the kernels' own rates "
                                 "are roofline.resampling of the default line (K2 and K3 with weights resident in HBM)"},
            "cpu_baseline": ({"value": cpu_head["cpu_particles_per_s"], "unit": "particles/s", "cores": 1, "kind": "port",
                              "sample": f"oracle.resample, systematic, exp(N(0,4)) weights, n = {cpu_head['n']}"} if cpu_head else None),
            "e2e": {"value": head[key], "unit": "particles/s", "h2d_bytes_per_step": 8 * n + 8, "d2h_bytes_per_step": 4 * n} if gpu else
                   {"value": head[key], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": (4 if gpu else 0), "cases": cases}
    if not gpu:
        line["impl"] = "reference"
    print(json.dumps(line))


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
    d = build_model(wl_model).dimension
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sharded = wl_name == "c5"
    # the same config keys as the b200 arm (same workload, metric and unit); what was actually timed on the CPU is the
    # bounded sample described in config.sample / cpu_baseline.sample -- a RATE, cache-resident clouds favour the CPU
    out = {"metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
           "scaling": "strong" if sharded else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": desc, "particles_per_gpu": N // world if sharded else N,
                      "particles_total": N if sharded else N * world, "observations": T, "latent_dim": d,
                      "resampler": resampler, "tie_rule": "reference (TreeMap: last particle of a repeated key)",
                      "sample": {"what": last["sample"], "cores": cores, "rate_not_full_run": True},
                      "note": "CPU restatement of the reference (not the JVM), bounded sample of this workload on all host cores"},
           "cpu_baseline": last,
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def pmmh_leg(args, ctx, iters, warmup, chains_concurrent=(2, 4), with_cpu=True):
    """BASELINE.json configs[3]: PMMH on the negative-binomial + linear-trend model, 2^16 particles x 500 observations per
    likelihood evaluation, one chain per GPU (examples/DetermineParameters.scala:55-85; model/PMMH.scala:68-81).  Returns the
    record (rank 0) -- iterations/s with ONE chain per GPU from device events, the same through the public API by wall
    clock, and the rate with k chains sharing each GPU (the reference's mapAsync over chains; a 2^16-particle filter is
    latency bound and leaves most SMs idle, so chains overlap)."""
    import torch
    from composablestatespacemodels_b200 import MetropolisHastings, GpuBootstrapFilter, Data, perturb, runChains
    from composablestatespacemodels_b200.resampling import Resampling
    _, wl_model, N, T, resampler = WORKLOADS["c4"]
    if args.workload == "c4":
        N, T = ctx["N"], ctx["T"]
    rank, world, local, stream, dtype = ctx["rank"], ctx["world"], ctx["local"], ctx["stream"], ctx["dtype"]
    um, p0 = build_unparam(wl_model)
    mod = um(p0)
    d = mod.dimension
    t, y = synth_series(mod, wl_model, T)
    data = [Data(tt, yy) for tt, yy in zip(t, y)]

    def chain(c, own_stream):
        rng = np.random.default_rng(100 + rank * 64 + c)
        pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, N, dtype=dtype, device=local, seed=5,
                                stream_id=rank * 64 + c)
        if not own_stream:
            pf.handle.set_stream(stream.cuda_stream)
        # examples/DetermineParameters.scala:59,73: Parameters.perturb(0.05), flat prior, symmetric proposal
        mh = MetropolisHastings(p0, perturb(0.05, rng), lambda a, c_: 0.0, lambda p: 0.0, pf, rng)
        return pf, mh.iters()

    # ---- one chain per GPU: device events on the launching stream ------------------------------------------
    pf, it = chain(0, own_stream=False)
    for _ in range(warmup):
        next(it)
    pf.handle.profile(1)
    sampler = ClockSampler(local)
    ctx["barrier"]()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, lls = 0, []
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(iters):
        st = next(it)
        lls.append(st.ll)
        launches += pf.handle.last_launches() + 1
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ctx["barrier"]()
    clocks = sampler.stop()
    ms = ctx["max_over_ranks"](e0.elapsed_time(e1))
    wall = ctx["max_over_ranks"](wall)
    prof = pf.handle.profile_read()
    dev_ms = pf.handle.last_elapsed_ms()
    accepted = int(st.accepted)
    pf.close()
    rec = {"metric": "pmmh-iterations/sec", "unit": "pmmh-iterations/s", "workload": WORKLOADS["c4"][0],
           "chains_per_gpu": 1, "n_gpus": world, "iterations": iters, "warmup": warmup,
           "value": world * iters / (ms * 1e-3), "ms_per_iteration": ms / iters,
           "e2e": {"value": world * iters / wall, "unit": "pmmh-iterations/s", "h2d_bytes_per_step": int(T * (4 * d + 8) * 4),
                   "d2h_bytes_per_step": 8 + 8 * d,
                   "note": "MetropolisHastings(...).iters() over a GpuBootstrapFilter: per iteration the proposed parameters go in "
                           "(per-observation constants rebuilt on the host, one pinned asynchronous upload), one log-likelihood "
                           "and one sampled state come out; wall clock"},
           "device_ms_per_likelihood": dev_ms, "us_per_observation": dev_ms * 1e3 / T, "accepted": accepted,
           "particle_steps_per_s": world * iters / (ms * 1e-3) * N * T, "gpu_launches": int(launches), "clocks": clocks,
           "log_likelihood_mean": float(np.mean(lls))}
    s_ms, s_n = prof.get("series", (0, 0))
    if s_n:
        rec["series_kernel_ms_per_launch"] = s_ms / s_n
    # ---- k chains sharing each GPU (threads; every handle has its own stream) --------------------------------
    conc = []
    for k in chains_concurrent:
        pairs = [chain(1 + c, own_stream=True) for c in range(k)]
        runChains([p[1] for p in pairs], max(2, warmup), parallelism=k)
        sampler = ClockSampler(local)
        ctx["barrier"]()
        sampler.start()
        w0 = time.perf_counter()
        runChains([p[1] for p in pairs], iters, parallelism=k)
        torch.cuda.synchronize()
        wall_k = ctx["max_over_ranks"](time.perf_counter() - w0)
        ctx["barrier"]()
        ck = sampler.stop()
        for p_ in pairs:
            p_[0].close()
        conc.append({"chains_per_gpu": k, "value": world * k * iters / wall_k, "unit": "pmmh-iterations/s",
                     "per_chain": iters / wall_k, "timing": "wall clock around runChains (host threads, one stream per chain)",
                     "clocks": ck})
    rec["concurrent_chains"] = conc
    if rank == 0 and with_cpu:
        cpu = cpu_baseline(wl_model, resampler, T, budget_s=6.0, threads=1, pmmh=True)
        rec["cpu_baseline"] = cpu
        per_gpu = rec["value"] / world
        best = max([per_gpu] + [c["value"] / world for c in conc])
        rec["vs_cpu"] = {"one_chain_per_gpu_vs_one_core": per_gpu / cpu["value"],
                         "best_per_gpu_vs_one_core": best / cpu["value"],
                         "target": "BASELINE.md: >= 1000 x the CPU port at one core per chain",
                         "met_like_for_like": bool(per_gpu / cpu["value"] >= 1000.0),
                         "note": "like for like = one chain per GPU against one chain per core of the flat C++ port; "
                                 "the reference-shaped variant (per-particle objects, std::map ECDF) is %.1fx slower still"
                                 % (cpu["value"] / (cpu["faithful_value"] / (N * T)) if cpu.get("faithful_value") else float("nan"))}
    return rec


def sharded_leg(args, ctx, steps, warmup):
    """BASELINE.json configs[4]: ONE filter of 2^27 particles (Normal + seasonal + OU), 100 observations, sharded over the
    GPUs of the job -- in-kernel NVLink exchange of the per-rank maxima / exact sums, ancestors scattered to the rank that
    owns the slot, parents gathered from the rank that owns them.  Strong scaling against the same filter on ONE GPU
    (timed on rank 0 in this run) and the exposed exchange time per observation against an unsharded filter of the
    per-rank size.  Only called with world > 1."""
    import torch
    import composablestatespacemodels_b200 as cs
    from composablestatespacemodels_b200 import sharding
    from composablestatespacemodels_b200.resampling import Resampling
    desc, wl_model, N, T, resampler = WORKLOADS["c5"]
    if args.workload == "c5":
        N, T = ctx["N"], ctx["T"]
    rank, world, local, stream, dtype = ctx["rank"], ctx["world"], ctx["local"], ctx["stream"], ctx["dtype"]
    mod = build_model(wl_model)
    d = mod.dimension
    b = 4 if dtype == 0 else 8
    t, y = synth_series(mod, wl_model, T)
    kind = Resampling.kind_of(resampler)
    n_local = N // world

    def timed(h, k, w, everyone):
        for _ in range(w):
            h.ll_resident()
        h.profile(10)
        sm = ClockSampler(local)
        if everyone:
            ctx["barrier"]()
        else:
            torch.cuda.synchronize()
        sm.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        lls, launches = [], 0
        for _ in range(k):
            lls.append(h.ll_resident())
            launches += h.last_launches()
        e1.record(stream)
        if everyone:
            ctx["barrier"]()
        else:
            torch.cuda.synchronize()
        ck = sm.stop()
        ms = e0.elapsed_time(e1)
        if everyone:
            ms = ctx["max_over_ranks"](ms)
        prof = h.profile_read()
        h.profile(0)
        per = {k_: (v[0] / v[1] if v[1] else 0.0) for k_, v in prof.items() if v[1]}
        return ms / k, per, ck, lls, launches

    # (1) the sharded filter, all ranks
    h = sharding.create_sharded(mod, kind, N, dtype=dtype, device=local, seed=2)
    h.set_stream(stream.cuda_stream)
    h.load_series(t, y)
    ms_sh, per_sh, clocks_sh, lls_sh, launches = timed(h, steps, warmup, True)
    h.reseed(2, 0)
    ll_sh = h.ll_resident()  # a freshly seeded run: the unsharded filter below must return the same bits
    h.close()
    ctx["barrier"]()
    # (2) rank 0 alone: the same filter unsharded (strong-scaling base) and an unsharded filter of the per-rank size
    #     (the local work of one rank without any exchange); the other ranks wait at the barrier
    ms_1 = ms_loc = None
    per_1 = per_loc = {}
    ll_1 = None
    clocks_1 = None
    if rank == 0:
        h1 = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, device=local, seed=2)
        h1.set_stream(stream.cuda_stream)
        h1.load_series(t, y)
        ms_1, per_1, clocks_1, lls_1, _ = timed(h1, max(2, steps // 2), 1, False)
        h1.reseed(2, 0)
        ll_1 = h1.ll_resident()
        h1.close()
        hl = cs.GpuFilterHandle(mod, kind, n_local, dtype=dtype, device=local, seed=2)
        hl.set_stream(stream.cuda_stream)
        hl.load_series(t, y)
        ms_loc, per_loc, _, _, _ = timed(hl, max(2, steps // 2), 1, False)
        hl.close()
    ctx["barrier"]()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    k1_bpp = 2 * d * b + b + 4
    k1 = per_sh.get("propagate_weight", 0.0)
    rec = {"metric": "particle-steps/sec", "unit": "particle-steps/s", "workload": desc, "scaling": "strong",
           "n_gpus": world, "particles_total": N, "particles_per_gpu": n_local, "observations": T, "steps": steps, "warmup": warmup,
           "value": N * T / (ms_sh * 1e-3), "ms_per_step": ms_sh, "us_per_observation": ms_sh * 1e3 / T,
           "kernel_ms_per_launch": per_sh,
           "one_gpu": {"value": N * T / (ms_1 * 1e-3), "ms_per_step": ms_1, "kernel_ms_per_launch": per_1, "clocks": clocks_1},
           "speedup_vs_one_gpu": ms_1 / ms_sh, "strong_scaling_efficiency": ms_1 / ms_sh / world,
           "local_unsharded": {"particles": n_local, "us_per_observation": ms_loc * 1e3 / T, "kernel_ms_per_launch": per_loc,
                               "what": "an unsharded filter of the per-rank size on one GPU: the local work of a rank with no exchange"},
           "exposed_exchange_us_per_observation": (ms_sh - ms_loc) * 1e3 / T,
           "same_bits_as_one_gpu": bool(ll_sh == ll_1),
           "roofline": {"bound": "hbm", "kernel": "k_propagate_weight on each rank (parents gathered over NVLink where they live on a peer)",
                        "achieved": k1_bpp * n_local / (k1 * 1e-3) / 1e9 if k1 else None, "peak": peak, "unit": "GB/s",
                        "frac": k1_bpp * n_local / (k1 * 1e-3) / 1e9 / peak if k1 else None, "traffic": None, "peak_source": peak_src},
           "exchange": "per observation: max log-weight, exact (sum w, sum w^2), resampling done -- 8..32-byte stores into every "
                       "peer's memory + release flag from the last block of the producing kernel, polled by the first warp of the "
                       "consuming kernel; no collective launch",
           "gpu_launches": int(launches), "clocks": clocks_sh, "log_likelihood": float(ll_sh)}
    return rec


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="default workload only: skip the pmmh / sharded sub-records")
    ap.add_argument("--pmmh-iters", type=int, default=50, help="timed PMMH iterations of the pmmh sub-record")
    ap.add_argument("--chains", default="2,4", help="chains sharing a GPU in the concurrent PMMH measurements")
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
    if args.workload == "resample":
        return run_resample(args, wl)
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
    dtype = _abi.F32 if args.dtype == "f32" else _abi.F64
    b = 4 if dtype == _abi.F32 else 8
    stream = torch.cuda.current_stream()
    chains = tuple(int(c) for c in args.chains.split(",") if c.strip())

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

    ctx = {"rank": rank, "world": world, "local": local, "stream": stream, "dtype": dtype, "barrier": barrier,
           "max_over_ranks": max_over_ranks, "N": N, "T": T}

    if pmmh:
        # ---- --workload c4: the PMMH record IS the line ------------------------------------------------------
        rec = pmmh_leg(args, ctx, args.steps, args.warmup, chains_concurrent=chains, with_cpu=not args.no_cpu)
        if rank == 0:
            peak, peak_src = measured_peak()
            d = build_model(wl_model).dimension
            real_bpp = (2 * d * b + b + 4) + b + b + 4
            s_ms = rec.get("series_kernel_ms_per_launch")
            roof = None
            if s_ms:
                ach = real_bpp * N * T / (s_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "kernel": "k_series (whole llFilter in one cooperative launch)", "achieved": ach, "peak": peak,
                        "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_particle": real_bpp, "avg_launch_ms": s_ms, "us_per_observation": s_ms * 1e3 / T,
                        "note": "latency bound: the 2^16-particle cloud lives in L2; grid-wide exchanges per observation"}
            out = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": world, "steps": args.steps,
                   "warmup": args.warmup, "ms_per_step": rec["ms_per_iteration"], "higher_is_better": True, "scaling": "weak",
                   "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                   "config": {"workload": desc, "particles_per_gpu": N, "particles_total": N * world, "observations": T, "latent_dim": d,
                              "resampler": resampler, "tie_rule": "reference (TreeMap: last particle of a repeated key)",
                              "l2": "working set %.1f MB per GPU is below the 126 MB L2, no flush" % ((2 * d * b + b + 4) * N / 1e6),
                              "parallelism": f"{world} independent chain(s), one per GPU, no collective"},
                   "roofline": roof, "cpu_baseline": rec.get("cpu_baseline"), "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"],
                   "clocks": rec["clocks"], "log_likelihood_mean": rec["log_likelihood_mean"]}
            for k in ("accepted", "particle_steps_per_s", "device_ms_per_likelihood", "concurrent_chains", "vs_cpu"):
                if k in rec:
                    out[k] = rec[k]
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    mod = build_model(wl_model)
    t, y = synth_series(mod, wl_model, T)
    d = mod.dimension
    kind = Resampling.kind_of(resampler)
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
    try:
        extra["scan_tiles_last_run"] = dict(zip(("certified_fp64", "exact_fallback"), h.scan_stats()))
    except Exception:
        pass
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

    # ---- BASELINE.json's other two numbers as sub-records of the default line --------------------------------
    pmmh_rec = shard_rec = None
    if args.workload == "target" and not args.no_extra and not args.particles and not args.obs:
        pmmh_rec = pmmh_leg(args, ctx, args.pmmh_iters, 3, chains_concurrent=chains, with_cpu=not args.no_cpu)
        if world > 1:
            shard_rec = sharded_leg(args, ctx, 4, 2)

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
            traffic, traffic_note = None, None
            try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture: same shape AND same kernel source
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
                if tr and tr["particles"] == n_local and args.dtype == "f32":
                    if tr.get("kernels_blob") == kernels_blob():
                        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                        traffic_note = "ncu capture " + tr.get("source", "")
                    else:
                        traffic_note = ("null: profiles/ncu_traffic.json was captured on kernel sources %s, this tree is %s"
                                        % (tr.get("kernels_blob"), kernels_blob()))
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": "k_propagate_weight (gather + propagate + weight, fused)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_note,
                    "algorithmic_bytes_per_launch": k1_bpp * n_local,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": k1_bpp,
                    "avg_launch_ms": k1_ms / k1_n, "sampled_launches": k1_n,
                    "kernel_ms_per_launch": per_launch,
                    # scan / search / gather throughput of the resampling stage (north_star): particles per second through
                    # K2 (exp + exact sums) and K3 (CDF scan + ancestor search); the gather is fused into K1's load
                    "resampling": {"weight_sums_particles_per_s": n_local / (per_launch["weight_sums"] * 1e-3) if per_launch["weight_sums"] else None,
                                   "scan_search_particles_per_s": n_local / (per_launch["scan_search"] * 1e-3) if per_launch["scan_search"] else None,
                                   "bytes_per_particle": {"weight_sums": b, "scan_search": b + 4},
                                   "frac_of_hbm_peak": {"weight_sums": b * n_local / (per_launch["weight_sums"] * 1e-3) / 1e9 / peak if per_launch["weight_sums"] else None,
                                                        "scan_search": (b + 4) * n_local / (per_launch["scan_search"] * 1e-3) / 1e9 / peak if per_launch["scan_search"] else None}},
                    "whole_step": {"bytes_per_particle_fused": real_bpp,
                                   "achieved_gbs_fused": real_bpp * n_local / (tot * 1e-3) / 1e9 if tot else None,
                                   "frac_fused": real_bpp * n_local / (tot * 1e-3) / 1e9 / peak if tot else None,
                                   "bytes_per_particle_survey_8d": 4 * d * b + 5 * b + 8,
                                   "frac_survey_8d": (4 * d * b + 5 * b + 8) * n_local / (tot * 1e-3) / 1e9 / peak if tot else None,
                                   "note": "frac_fused counts the bytes the three kernels move (the gather never materialises a cloud, no "
                                           "CDF is written); frac_survey_8d rates the same step time with SURVEY 8(d)'s B_step, the bytes of "
                                           "an unfused step -- it can exceed 1 and is a speed-up over that step at the roof, not a bandwidth"}}
        elif prof.get("series", (0, 0))[1]:
            # small cloud: the whole llFilter is ONE cooperative launch (cssm_series.cuh); per launch it moves
            # T x N x (K1 2db+b+4, K2 b, K3 b+4) bytes, all of it L2 resident: latency bound, not HBM bound
            s_ms, s_n = prof["series"]
            real_bpp = k1_bpp + b + b + 4
            ach = real_bpp * n_local * T / (s_ms / s_n * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_series (whole llFilter in one cooperative launch: propagate + weight, "
                                              "exact sums, scan + search per observation, grid-wide exchanges in between)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": real_bpp,
                    "avg_launch_ms": s_ms / s_n, "sampled_launches": s_n,
                    "us_per_observation": s_ms / s_n * 1e3 / T,
                    "note": "latency bound: the cloud (%.1f MB) lives in L2; grid-wide exchanges per observation"
                            % (ws_bytes_small / 1e6)}
        elif k1_n:
            # LGCP: the state lives in registers for ~100 sub-steps per event: bound by instruction issue, not by HBM.
            # Instruction roofline: SASS instructions per particle-sub-step (counted in the disassembly of the unrolled
            # d = 1 loop: 36) x sub-steps/s against the issue rate of the SMs (4 schedulers x 32 lanes per SM and clock).
            nsub = float(np.sum(np.ceil(np.diff(np.concatenate([[t[0]], t])) / 1e-3)))
            sub_rate_kernel = n_local * (nsub / T) / (k1_ms / k1_n * 1e-3)  # inside the kernel, per GPU
            sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
            issue_peak = 148 * 4 * 32 * sm_clock  # thread-instructions/s
            INSTR_PER_SUBSTEP = 36
            roof = {"bound": "issue", "kernel": "k_lgcp_weight (sub-stepped propagate + hazard, fused)",
                    "achieved": sub_rate_kernel * INSTR_PER_SUBSTEP / 1e12, "peak": issue_peak / 1e12, "unit": "T thread-instr/s",
                    "frac": sub_rate_kernel * INSTR_PER_SUBSTEP / issue_peak, "traffic": None,
                    "peak_source": "148 SMs x 4 schedulers x 32 lanes x the SM clock sampled during the run",
                    "instructions_per_substep": INSTR_PER_SUBSTEP,
                    "hbm": {"achieved": k1_bpp * n_local / (k1_ms / k1_n * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": k1_bpp * n_local / (k1_ms / k1_n * 1e-3) / 1e9 / peak, "peak_source": peak_src},
                    "note": "instruction bound: Philox + Box-Muller + exp per sub-step, state in registers; the HBM fraction is secondary",
                    "particle_substeps_per_s": n_total * nsub * args.steps / (ms * 1e-3),
                    "avg_launch_ms": k1_ms / k1_n, "sampled_launches": k1_n}
        cpu = None
        if not args.no_cpu:
            cpu = cpu_baseline(wl_model, resampler, T, budget_s=12.0, threads=1, pmmh=False)
        if sharded:
            par = (f"one filter of {n_total} particles sharded over {world} GPU(s), {n_local} per GPU; per step three in-kernel "
                   "exchanges (max, sums, done) + ancestor scatter / parent gather over NVLink peer pointers" if world > 1
                   else "one filter on one GPU (the unsharded case of the strong-scaling series)")
        else:
            par = f"{world} independent filter(s), one per GPU, no collective"
        ws_bytes = k1_bpp * n_local
        out = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if sharded else "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": desc, "particles_per_gpu": n_local, "particles_total": n_total, "observations": T,
                          "latent_dim": d, "resampler": resampler,
                          "tie_rule": "first index (NOT the reference's rule)" if args.tie_first else "reference (TreeMap: last particle of a repeated key)",
                          "l2": "working set %.0f MB per GPU %s the 126 MB L2, no flush" % (
                              ws_bytes / 1e6, "exceeds" if ws_bytes > 126e6 else "is below"),
                          "parallelism": par,
                          "timing": "value: CUDA events around the K steps, with per-kernel events on every 10th observation inside "
                                    "(those observations run without programmatic dependent launch: at 2^20 particles that costs "
                                    "several per cent, at 2^24 nothing measurable); e2e: wall clock, no per-kernel events"},
               "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "log_likelihood_mean": float(np.mean(lls))}
        out.update(extra)
        if pmmh_rec is not None:
            out["pmmh"] = pmmh_rec
        if shard_rec is not None:
            out["sharded"] = shard_rec
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
