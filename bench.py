#!/usr/bin/env python
"""bench.py — individual·locus·generations/s of the GeneEvolve reproduction hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is ONE generation (mating -> crossover sampling -> bit-packed haplotype propagation -> mutation ->
causal-variant genetic values -> phenotypes -> mating/selection values) of the named workload, default the
north-star target `config3_100k_x_1M` (100 000 individuals x 1 000 000 loci on 22 autosomes, assortative
mating rho = 0.4, logit selection, h2 = 0.5, Philox draws on the device).  One JSON line is printed by rank 0.

`value`  : whole-job throughput with the generation state resident in HBM (the library keeps it there).
`e2e`    : the same through the reference-facing C-ABI with HOST buffers — per step the generation
           parameters go host->device and the `.info` columns of every individual (what the reference
           writes each generation, src/Population.cpp:510-568) come back into pinned host memory.
`roofline`: propagate_bits_kernel, algorithmic bytes (0.5 B per individual-locus) / CUDA-event
           time of that kernel measured on the library's stream inside the timed region, vs MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own binary (oracle/_ref/GeneEvolve_ref, built from its unmodified sources) timed
           on this box on a bounded sample of the same workload, one core (the reference is single-threaded).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "individual_locus_generations_per_s"
UNIT = "individual*locus*generations/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "GeneEvolve_ref")


def rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc, self.samples = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)  # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            self.samples = [[x.strip() for x in line.split(",")] for line in out.strip().splitlines() if line.count(",") >= 5]

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference's own CPU implementation, timed generation by generation from its stdout
# ------------------------------------------------------------------------------------------------
def run_reference_sample(cfg_name, n_sample, n_gen, n_proc, seed0=12345):
    """Runs n_proc independent replicate processes of the unmodified reference on a bounded sample of the
    workload (n_sample individuals and founders, same genetic map, CV set and mating scheme) for n_gen
    generations; returns per-process lists of per-generation wall seconds (high-resolution timestamps of its
    own "Start generation" lines) — the reference is single-threaded, so n_proc = cores used."""
    from geneevolve_b200 import workloads
    cfg = workloads.make_workload(cfg_name, n_override=n_sample)
    cfg["founders"] = n_sample
    rng = np.random.default_rng(5)
    for cv in cfg["cvs"]:  # founders changed -> fresh CV panel of the right width
        k = len(cv["bp"])
        cv["val"] = (rng.random((2 * n_sample, k)) < 0.5).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        args = workloads.write_reference_inputs(cfg, d, n_gen)
        procs = []
        for r in range(n_proc):
            cmd = [REF_BIN] + args + ["--seed", str(seed0 + r), "--prefix", os.path.join(d, f"o{r}")]
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1))
        stamps = [[] for _ in procs]

        def reader(k):
            for line in procs[k].stdout:
                if "Start generation" in line or "Time taken for simulation" in line:
                    stamps[k].append(time.perf_counter())
        th = [threading.Thread(target=reader, args=(k,)) for k in range(n_proc)]
        [t.start() for t in th]
        [t.join() for t in th]
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("reference binary failed")
    return [np.diff(s) for s in stamps], cfg


def reference_arm(args):
    rank, world, _ = rank_world()
    if rank != 0:
        return
    from geneevolve_b200 import workloads
    cfgM = sum(workloads.make_workload(args.workload, n_override=64)["n_loci"])
    if not os.path.exists(REF_BIN):
        # the reference did not compile here: time the oracle port of the HBM-bound kernel instead
        from oracle import oracle
        t, _ = oracle.bench_propagate_bits(2000, 2000, cfgM, 36, 1)
        v = 2000 * cfgM / t
        print(json.dumps({"impl": "reference", "metric": METRIC, "unit": UNIT, "value": v, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                          "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
                          "data": "synthetic", "config": {"workload": args.workload},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": "2000 offspring x %d loci bit-packed propagation" % cfgM},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    cores = os.cpu_count() or 1
    n_sample = args.ref_sample
    per_gen, _ = run_reference_sample(args.workload, n_sample, args.warmup + args.steps, cores)
    timed = np.array([g[args.warmup:args.warmup + args.steps] for g in per_gen])  # [proc][step] seconds
    ms = float(timed.mean(axis=0).mean() * 1e3)
    # every process advances n_sample individuals per step; all cores run concurrently
    value = float((n_sample * cfgM / timed).sum(axis=0).mean())
    sample = "%d replicate processes x %d individuals x 22 chr, %d timed generations each (same map, CVs, mating scheme); M = %d nominal loci " \
             "(the reference stores founder segments and never touches non-causal loci)" % (cores, n_sample, args.steps, cfgM)
    print(json.dumps({"impl": "reference", "metric": METRIC, "unit": UNIT, "value": value, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "u64 segments + f64", "data": "synthetic",
                      "config": {"workload": args.workload, "individuals_per_process": n_sample, "processes": cores, "loci_nominal": cfgM,
                                 "individual_generations_per_s": value / cfgM,
                                 "note": "the reference stores founder segments and never touches non-causal loci: its cost does not depend on M, "
                                         "so individual*locus*generations/s is individual*generations/s times the nominal M"},
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_single(args, M):
    """One core, bounded sample (~10-30 s): the figure quoted next to the GPU number."""
    try:
        if os.path.exists(REF_BIN):
            n_sample, n_gen = args.ref_sample, 3
            per_gen, _ = run_reference_sample(args.workload, n_sample, n_gen, 1)
            t = float(per_gen[0][1:].mean())  # generations 2..3 (the first one mates the founders)
            return {"value": n_sample * M / t, "unit": UNIT, "cores": 1, "kind": "reference",
                    "sample": "%d individuals x 22 chr x %d generations of the same map/CV/mating scheme, 1 process; M = %d nominal loci; "
                              "%.0f individual*generations/s" % (n_sample, n_gen - 1, M, n_sample / t)}
        from oracle import oracle
        t, _ = oracle.bench_propagate_bits(2000, 2000, M, 36, 1)
        return {"value": 2000 * M / t, "unit": UNIT, "cores": 1, "kind": "port", "sample": "2000 offspring x %d loci, bit-packed propagation only" % M}
    except Exception as e:  # the baseline must never take the GPU number down with it
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}


# ------------------------------------------------------------------------------------------------
def state_hash(eng, n_pop):
    """Integer fingerprint of the simulated state: pedigree, sex and couples of every population (crc32).  The fp64 columns are
    left out on purpose — sharded runs add the chromosomes' partial genetic values in another order (1e-16) — but couples, family
    sizes and pedigree depend on every phenotype through selection and the mating-value sorts, so equal hashes across --gpus N mean
    the sharded runs simulated the same populations."""
    import zlib
    h = 0
    for p in range(n_pop):
        ind = eng.individuals(p)
        c = eng.get_couples(p)
        for a in (ind["ids"], ind["sex"], c["pos_male"], c["pos_female"], c["num_offspring"]):
            h = zlib.crc32(np.ascontiguousarray(a).tobytes(), h)
    return h


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


SETTLE = 6   # untimed generations before any timed region, at least: the library records its two generation graphs (with the NCCL
             # collective inside on a sharded run: tens of milliseconds of capture and instantiation) once its buffers stop moving


def untimed_generations(warmup):
    """The W warm-up steps, or SETTLE of them when W is smaller (reported as config.untimed_generations_before_timing)."""
    return max(int(warmup), SETTLE)


def measure_workload(name, local, steps, warmup, n=None, loci=None, e2e=True, phases_steps=5, flags=0):
    """One single-GPU workload through the C-ABI: device-resident arm, end-to-end arm, dominant-kernel roofline (CUDA events on the
    kernel's own stream inside the timed region) and, in a short extra pass, the phases of the control chain."""
    import torch
    from geneevolve_b200 import capi, workloads
    cfg = workloads.make_workload(name, n_override=n, loci_override=loci)
    M, N = sum(cfg["n_loci"]), cfg["n"]
    cap = int(max(N, cfg["founders"]) * 1.03) + 1024
    segs = bool(cfg.get("segments"))
    warmup = untimed_generations(warmup)
    total_steps = warmup + steps * (2 if e2e else 1) + phases_steps
    seg_cap = 0
    if segs:  # parts per haplotype-genome after g generations ~ n_chr + g * (map length in Morgans)
        morgans = sum(float(p.sum()) for _, _, p in cfg["maps"])
        seg_cap = int(2 * cap * (len(cfg["chrs"]) + (total_steps + 1) * morgans) * 1.05)
    eng = capi.Engine(seg_capacity=seg_cap, n_pop=1, n_chr=len(cfg["chrs"]), n_phen=1, device=local, representation=capi.GE_REP_SEGMENTS if segs else capi.GE_REP_BITS,
                      rng_mode=capi.GE_RNG_PHILOX, seed=12345, capacity=cap, flags=flags)
    kid = capi.GE_KERNEL_RECOMBINE_SEGMENTS if segs else capi.GE_KERNEL_PROPAGATE_BITS
    kname = "seg_plan_kernel + seg_gather_kernel" if segs else "propagate_bits_kernel"
    workloads.configure_engine(eng, cfg)
    eng.init_generation0()
    gp = [capi.gen_params(N, cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    gen = 0
    for _ in range(warmup):
        gen += 1
        eng.step_generation(gen, gp)
    # ---- device-resident arm
    eng.set_profiling(1)
    eng.reset_kernel_times()
    eng.synchronize()
    work = 0
    with ClockSampler(local) as clocks:
        eng.timer_start()
        for _ in range(steps):
            gen += 1
            eng.step_generation(gen, gp)
            work += eng.population_size(0) * M
        ms_dev = eng.timer_stop()
    launches = eng.launch_count()
    k_ms, k_n, k_bytes = eng.kernel_time(kid)
    eng.set_profiling(0)
    res = {"cfg": cfg, "M": M, "N": N, "cap": cap, "segs": segs, "kname": kname, "ms_dev": ms_dev, "work": work, "launches": launches,
           "k_ms": k_ms, "k_n": k_n, "k_bytes": k_bytes, "clocks": clocks.summary(), "first_timed_gen": warmup + 1}
    # ---- end-to-end arm: host parameters in, `.info` columns out to pinned host memory, every step
    if e2e:
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
        out = {"ids": pin((cap, 7), torch.int64).view(np.uint64), "sex": pin((cap,), torch.uint8)}
        for k in "ADGCEFP":
            out[k] = pin((1, cap), torch.float64)
        for k in ("mv", "sv", "svf"):
            out[k] = pin((cap,), torch.float64)
        eng.synchronize()
        work2, d2h = 0, 0
        eng.timer_start()
        for _ in range(steps):
            gen += 1
            eng.step_generation(gen, gp)
            nn = eng.population_size(0)
            eng.individuals(0, out=out)
            work2 += nn * M
            d2h += capi.Engine.individual_bytes(nn, 1)
        res.update(ms_e2e=eng.timer_stop(), work2=work2, d2h=d2h // steps, checksum=float(out["P"].reshape(-1)[:16].sum()), state_hash=state_hash(eng, 1))
    # ---- the control chain's phases (CUDA events between its kernels: queued kernel by kernel, hence outside the timed regions)
    eng.set_profiling(2)
    eng.reset_kernel_times()
    for _ in range(phases_steps):
        gen += 1
        eng.step_generation(gen, gp)
    res["phases"] = {nm: eng.kernel_time(pid)[0] / phases_steps for nm, pid in capi.GE_PHASES.items()}
    eng.set_profiling(0)
    res["device_memory_gb"] = eng.device_memory_bytes() / 1e9
    eng.close()
    return res


def roofline_of(r, traffic=None, traffic_source=None):
    peak, peak_src = hbm_peak()
    achieved = r["k_bytes"] / (r["k_ms"] * 1e-3) / 1e9 if r["k_ms"] > 0 else None
    return {"bound": "hbm", "kernel": r["kname"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
            "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src, "kernel_ms_per_launch": r["k_ms"] / max(r["k_n"], 1),
            "kernel_share_of_step": r["k_ms"] / r["ms_dev"], "algorithmic_bytes_per_launch": r["k_bytes"] // max(r["k_n"], 1)}


def other_workloads(local, budget_s=60.0):
    """The other BASELINE configurations a single GPU can hold, each as a short record next to the headline (config 2 whole; config 5
    as a founder-segment sample: 250 000 individuals, generations 41-45)."""
    out = []
    t0 = time.perf_counter()
    for name, kw in (("config2_chr22_10k", dict(steps=40, warmup=5)),
                     ("config1_bundled_1chr", dict(steps=100, warmup=5)),
                     ("config5_1M_x_10M_segments", dict(steps=5, warmup=40, n=250000))):
        if time.perf_counter() - t0 > budget_s:
            out.append({"workload": name, "skipped": "time budget"})
            continue
        try:
            r = measure_workload(name, local, e2e=False, phases_steps=3, **kw)
            rec = {"workload": name, "individuals": r["N"], "loci": r["M"], "steps": kw["steps"], "warmup": kw["warmup"], "ms_per_step": r["ms_dev"] / kw["steps"],
                   "value": r["work"] / (r["ms_dev"] * 1e-3), "unit": UNIT, "gpu_launches_per_step": r["launches"] / kw["steps"],
                   "roofline": roofline_of(r), "control_chain_ms_per_step": r["phases"], "device_memory_gb": r["device_memory_gb"]}
            if r["segs"]:
                rec["note"] = "founder segments: loci nominal, cost grows with the generation (timed steps are generations %d..%d); algorithmic bytes = 8 B per " \
                              "part read by the plan, read by the gather and written" % (r["first_timed_gen"], r["first_timed_gen"] + kw["steps"] - 1)
            out.append(rec)
        except Exception as e:  # never take the headline down
            out.append({"workload": name, "failed": repr(e)})
    return out


def ours(args):
    import torch
    from geneevolve_b200 import capi, workloads
    rank, world, local = rank_world()
    if world > 1:
        from geneevolve_b200 import dist as gdist
        return gdist.bench_sharded(args, METRIC, UNIT)
    torch.cuda.set_device(local)
    if "pops" in workloads.CONFIGS[args.workload]:
        raise SystemExit("multi-population workloads run sharded: launch with torchrun (config 4 needs >= 4 GPUs at full size)")
    from geneevolve_b200 import capi as _capi
    flags = _capi.GE_FLAG_SERIAL if args.serial else 0
    r = measure_workload(args.workload, local, args.steps, args.warmup, n=args.n, loci=args.loci, flags=flags)
    cfg, M, N, segs = r["cfg"], r["M"], r["N"], r["segs"]
    traffic = traffic_source = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not segs and args.workload == "config3_100k_x_1M" and not args.n and not args.loci:
        traffic = json.load(open(tpath)).get("propagate_bits_dram_bytes_per_launch")
        traffic_source = "profiles/traffic.json (static: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload, not measured in this run)"
    value = r["work"] / (r["ms_dev"] * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_dev"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-packed + f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "individuals": N, "loci": M, "chromosomes": len(cfg["chrs"]), "founders": cfg["founders"],
                   "causal_variants": sum(len(c["bp"]) for c in cfg["cvs"]), "mating": "random" if cfg["rm"] else "assortative rho=%.1f" % cfg["mat_cor"],
                   "selection": "logit(0,1)", "h2": 0.5, "rng": "philox4x32-10 on device",
                   "l2": ("inputs larger than L2 (%.1f GB of parental rows per step vs 126 MB)" % (N * M / 4 / 1e9)) if not segs else
                         "inputs larger than L2 (founder-segment lists, %.1f GB written per step)" % (r["k_bytes"] / max(r["k_n"], 1) / 2e9),
                   "representation": "founder segments (loci nominal; cost grows with the generation: steps are generations %d..%d)" % (untimed_generations(args.warmup) + 1, untimed_generations(args.warmup) + args.steps)
                   if segs else "bit-packed haplotypes",
                   "untimed_generations_before_timing": untimed_generations(args.warmup),
                   "device_memory_gb": r["device_memory_gb"], "individual_generations_per_s": value / M,
                   "scaling_note": "the workload is fixed; --gpus N splits its loci over N ranks"},
        "e2e": {"value": r["work2"] / (r["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 40, "d2h_bytes_per_step": r["d2h"],
                "ms_per_step": r["ms_e2e"] / args.steps, "checksum": r["checksum"], "state_hash": r["state_hash"],
                "generations_simulated": untimed_generations(args.warmup) + 2 * args.steps,
                "note": "generation state stays in HBM between steps by design (as it stays in process memory in the reference); per step the host sends the "
                        "generation-table row and receives every individual's .info columns in pinned memory; state_hash = crc32 of pedigree, sex and couples "
                        "after generations_simulated generations (equal across --gpus N when the sharded runs simulated the same populations)"},
        "gpu_launches": r["launches"],
        "roofline": roofline_of(r, traffic, traffic_source),
        "control_chain_ms_per_step": r["phases"],
        "clocks": r["clocks"],
    }
    line["cpu_baseline"] = cpu_baseline_single(args, M) if not args.no_cpu_baseline else None
    if not args.no_other_workloads and args.workload == "config3_100k_x_1M" and not args.n and not args.loci:
        line["other_workloads"] = other_workloads(local)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3_100k_x_1M")
    ap.add_argument("--n", "--individuals", dest="n", type=int, default=None,
                    help="override individuals per generation (debug; torch.distributed.run reads a bare --n as one of its own options: use --individuals there)")
    ap.add_argument("--loci", type=int, default=None, help="override loci (debug)")
    ap.add_argument("--ref-sample", type=int, default=1500, help="individuals in the bounded reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--serial", action="store_true", help="measurement aid: no overlap of the bulk copy with the next generation's control chain")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the short records of the other BASELINE configurations")
    ap.add_argument("--collective", default="native", choices=["native", "hook"],
                    help="measurement aid (N > 1): 'hook' routes the all-reduce through a Python callback into torch.distributed instead of the library's own ncclAllReduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
