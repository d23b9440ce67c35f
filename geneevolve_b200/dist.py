"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the collective.

Sharding axis: every rank holds ALL individuals but only ITS chromosomes.  A gamete's
chromosomes are independent given the parents, so the HBM-bound propagation, the crossover/mutation sampling,
the causal-variant planes and the allele counts are rank-local with no parent exchange at all; the one real
exchange step is the sum over ranks of the per-individual partial genetic values (NCCL all-reduce of
3 * n_phen * N doubles per population per generation), after which phenotypes, selection and mating are
computed redundantly and identically on every rank (Philox draws are keyed by global indices).
Sharding individuals instead would move 7/8 of every parental row over NVLink each generation (SURVEY.md §8e:
~2.7 GB inbound per GPU per generation at 8 GPUs, ~3.5 ms at 770 GB/s, against ~1 ms of local HBM time).
"""
import ctypes
import json
import os

import numpy as np


def assign_chromosomes(weights, world_size):
    """Longest-processing-time assignment of chromosomes (weights = loci per chromosome) to ranks."""
    order = sorted(range(len(weights)), key=lambda c: -weights[c])
    load = [0.0] * world_size
    mine = [[] for _ in range(world_size)]
    for c in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        mine[r].append(c)
        load[r] += weights[c]
    return [sorted(m) for m in mine]


def assign_locus_ranges(n_loci, world_size, align=128):
    """Balanced split of the bit-packed rows: the genome's 16-byte chunks (128 loci) in chromosome order are cut into world_size
    contiguous runs of equal length (to within one chunk), so a chromosome may span ranks — every rank that holds a slice of it
    draws that chromosome's crossovers itself (Philox counters are keyed by the global chromosome id, so the lists agree) and no
    parental row ever crosses a link.  Returns, per rank, a list of (chromosome index, first locus, end locus).  Unlike whole
    chromosomes (22 of very different lengths) this balances exactly, works with one chromosome and has no rank limit."""
    chunks = [(int(n) + align - 1) // align for n in n_loci]
    total = sum(chunks)
    cuts = [total * r // world_size for r in range(world_size + 1)]
    out = [[] for _ in range(world_size)]
    base = 0
    for c, (nc, nl) in enumerate(zip(chunks, n_loci)):
        for r in range(world_size):
            lo, hi = max(cuts[r], base), min(cuts[r + 1], base + nc)
            if lo < hi:
                out[r].append((c, (lo - base) * align, min((hi - base) * align, int(nl))))
        base += nc
    return out


class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def cuda_allreduce_hook(device, group=None):
    """Sum-allreduce of a raw device buffer on the library's stream through torch.distributed (NCCL)."""
    import torch
    import torch.distributed as dist

    cache = {}

    def hook(ptr, count, stream):
        key = (ptr, count, stream)
        if key not in cache:  # the library reuses one scratch buffer and one stream: wrap them once (a regrown buffer replaces the entry)
            cache.clear()
            cache[key] = (torch.as_tensor(_DevPtr(ptr, count), device=f"cuda:{device}"), torch.cuda.ExternalStream(stream, device=device))
        t, ext = cache[key]
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


def host_allreduce_hook(group=None):
    """Same for a host buffer (gloo) — used by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist

    def hook(ptr, count, stream):
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(count,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


def bench_sharded(args, METRIC, UNIT):
    """bench.py for WORLD_SIZE > 1: the same fixed workload, its loci spread over the ranks (strong scaling).  Bit-packed rows are
    split by locus range (assign_locus_ranges: exact balance, a chromosome may span ranks), founder segments by whole chromosomes."""
    import torch
    import torch.distributed as dist
    from . import capi, workloads
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = workloads.make_workload(args.workload, n_override=args.n, loci_override=args.loci)
    segs = bool(cfg.get("segments"))
    if segs:
        if len(cfg["chrs"]) < world:
            raise SystemExit("segment workloads are sharded by chromosome: fewer chromosomes than ranks")
        chrs = assign_chromosomes(cfg["n_loci"], world)[rank]
        pieces = [(c, 0, cfg["n_loci"][c]) for c in chrs]
    else:
        pieces = assign_locus_ranges(cfg["n_loci"], world)[rank]
    M, N = sum(cfg["n_loci"]), cfg["n"]
    pops = cfg.get("pops", [N])
    n_phen = cfg.get("n_phen", 1)
    cap = int(max(max(pops), cfg["founders"]) * (1.03 if len(pops) == 1 else 1.10)) + 1024   # migration moves ~2 % either way
    seg_cap = 0
    if segs:  # parts per haplotype after g generations ~ chromosomes + g * Morgans of this rank's chromosomes (both arms run back to back)
        morgans = sum(float(cfg["maps"][c][2].sum()) for c, _, _ in pieces)
        seg_cap = int(2 * cap * (len(pieces) + (args.warmup + 2 * args.steps + 1) * morgans) * 1.05)
    kid = capi.GE_KERNEL_RECOMBINE_SEGMENTS if segs else capi.GE_KERNEL_PROPAGATE_BITS
    eng = capi.Engine(n_pop=len(pops), n_chr=len(pieces), n_phen=n_phen, device=local, representation=capi.GE_REP_SEGMENTS if segs else capi.GE_REP_BITS,
                      rng_mode=capi.GE_RNG_PHILOX, seed=12345, capacity=cap, seg_capacity=seg_cap, rank=rank, world_size=world)
    if len(pops) > 1:
        workloads.configure_engine_multipop(eng, cfg, pieces=pieces)
    else:
        workloads.configure_engine(eng, cfg, pieces=pieces)
    eng.set_allreduce(cuda_allreduce_hook(local))
    eng.init_generation0()
    gp = [capi.gen_params(n, cfg["mat_cor"], "p", "logit", 0.0, 1.0) for n in pops]
    mig = cfg.get("migration")
    gen = 0
    for _ in range(args.warmup):
        gen += 1
        eng.step_generation(gen, gp, mig)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
    out = {"ids": pin((cap, 7), torch.int64).view(np.uint64), "sex": pin((cap,), torch.uint8)}
    for k in "ADGCEFP":
        out[k] = pin((n_phen, cap), torch.float64)
    for k in ("mv", "sv", "svf"):
        out[k] = pin((cap,), torch.float64)

    def timed(e2e):
        nonlocal gen
        eng.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        work = 0
        eng.timer_start()
        for _ in range(args.steps):
            gen += 1
            eng.step_generation(gen, gp, mig)
            for q in range(len(pops)):
                work += eng.population_size(q) * M
                if e2e and rank == 0:
                    eng.individuals(q, out=out)  # every rank holds identical columns; rank 0 feeds the host writers
        ms = eng.timer_stop()
        eng.synchronize()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return work, float(t.item())

    eng.set_profiling(1)
    eng.reset_kernel_times()
    with bench.ClockSampler(local) as clocks:
        work, ms_dev = timed(False)
    launches = eng.launch_count()
    k_ms, k_n, k_bytes = eng.kernel_time(kid)
    eng.set_profiling(0)
    work2, ms_e2e = timed(True)
    checksum = float(out["P"].reshape(-1)[:16].sum()) if rank == 0 else 0.0
    # every rank simulated the same populations: the integer fingerprint (pedigree, sex, couples) must agree on all of them,
    # and with the unsharded run of the same number of generations (bench.py --gpus 1 prints the same field)
    my_hash = bench.state_hash(eng, len(pops))
    hs = torch.tensor([my_hash], dtype=torch.int64, device="cuda")
    all_h = [torch.zeros_like(hs) for _ in range(world)]
    dist.all_gather(all_h, hs)
    hashes = [int(h.item()) for h in all_h]
    kb = torch.tensor([k_bytes / max(k_ms, 1e-9) / 1e6, k_ms / max(k_n, 1)], dtype=torch.float64, device="cuda")  # GB/s and ms per launch of this rank's kernel
    kmax = kb.clone()
    dist.all_reduce(kb, op=dist.ReduceOp.SUM)
    dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_base = bench.cpu_baseline_single(args, M)
    if rank == 0:
        peak, peak_src = bench.hbm_peak()
        achieved = float(kb[0].item()) / world  # mean per-GPU achieved GB/s
        print(json.dumps({
            "metric": METRIC, "value": work / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-packed + f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "individuals": sum(pops), "populations": pops, "phenotypes": n_phen, "loci": M, "chromosomes": len(cfg["chrs"]),
                       "parallelism": ("chromosome-sharded x%d (founder segments)" if segs else "locus-range sharded x%d (every rank: all individuals, 1/N of the 16-byte chunks of the rows)") % world,
                       "device_memory_gb_rank0": eng.device_memory_bytes() / 1e9,
                       "representation": "founder segments (loci nominal; steps are generations %d..%d)" % (args.warmup + 1, args.warmup + args.steps) if segs else "bit-packed haplotypes",
                       "pieces_rank0": pieces, "collective": "all-reduce of 3 * n_phen * capacity doubles per population and generation (NCCL)",
                       "l2": ("inputs larger than L2 (%.1f GB of parental rows per step per GPU)" % (sum(pops) * M / 4 / 1e9 / world)) if not segs else
                             "inputs larger than L2 (founder-segment lists, %.1f GB moved per step per GPU)" % (k_bytes / max(k_n, 1) * 2 / 1e9)},
            "e2e": {"value": work2 / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 40 * world,
                    "d2h_bytes_per_step": sum(capi.Engine.individual_bytes(n, n_phen) for n in pops), "ms_per_step": ms_e2e / args.steps,
                    "checksum": checksum, "state_hash": hashes[0], "state_hash_equal_on_all_ranks": len(set(hashes)) == 1,
                    "generations_simulated": args.warmup + 2 * args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "seg_plan_kernel + seg_gather_kernel" if segs else "propagate_bits_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "traffic_source": "not captured for sharded runs (ncu is single-GPU; profiles/traffic.json holds the 1-GPU capture)", "peak_source": peak_src,
                         "kernel_ms_per_launch": k_ms / max(k_n, 1), "kernel_ms_per_launch_slowest_rank": float(kmax[1].item()),
                         "kernel_share_of_step": k_ms / ms_dev, "note": "per-GPU mean"},
            "clocks": clocks.summary(), "cpu_baseline": cpu_base}))
    dist.barrier()
    dist.destroy_process_group()
