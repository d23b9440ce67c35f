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


def piece_cost_chunks(chunks_per_rank):
    """What one more chromosome piece costs a rank's copy kernel, in 16-byte chunks of row (measured, see assign_locus_ranges): it falls
    with the warps per offspring CTA (ge_api.cu build_genome: one warp per 16 KB of the offspring's two rows, 1..8) because the warps of a
    CTA work on different pieces side by side — 38 chunks with 2 warps (8 ranks of config 3), 16 with 8 warps (2 ranks)."""
    warps = min(8, max(1, (32 * int(chunks_per_rank) + 8192) // 16384))
    return int(round(8.7 + 58.7 / warps))


def assign_locus_ranges(n_loci, world_size, align=128, piece_cost=None):
    """Balanced split of the bit-packed rows: the genome's 16-byte chunks (128 loci) in chromosome order are cut into world_size
    contiguous runs, so a chromosome may span ranks — every rank that holds a slice of it draws that chromosome's crossovers itself
    (Philox counters are keyed by the global chromosome id, so the lists agree) and no parental row ever crosses a link.  Returns, per
    rank, a list of (chromosome index, first locus, end locus).  Unlike whole chromosomes (22 of very different lengths) this balances,
    works with one chromosome and has no rank limit.

    What is balanced is the measured cost of a rank, chunks + piece_cost * pieces: every chromosome piece a rank holds costs its copy
    kernel a fixed amount per offspring (crossover lists staged, short runs, a ragged last tile).  One rank's share of config 3 on a B200
    (scripts/emulate_rank.py, 8 ranks of 125k loci each): 2 pieces 1.066 ms per generation, 3 pieces 1.09, 4 pieces 1.137, 5 pieces
    1.170, 6 pieces 1.219 — 0.038 ms per piece against 0.99 ms for the 977 chunks, i.e. 38 chunks per piece; with equal chunk counts the
    rank holding chromosomes 17-22 set the pace of all eight (8 GPUs: 1.232 ms per generation with equal chunks, 1.170 with equal cost).
    With 2 ranks the same measurement gives 16 chunks per piece (piece_cost_chunks).  The cuts minimise the largest cost
    (piece_cost=0: equal chunk counts)."""
    chunks = [(int(n) + align - 1) // align for n in n_loci]
    total = sum(chunks)
    if piece_cost is None:
        piece_cost = piece_cost_chunks(total / max(world_size, 1))
    min_piece = min(8, piece_cost)   # never open a piece for fewer chunks than this unless the chromosome is that short

    def sweep(limit):
        out = [[] for _ in range(world_size)]
        r, used = 0, 0
        for c, nc in enumerate(chunks):
            pos = 0
            while pos < nc:
                if r == world_size - 1:
                    take = nc - pos
                else:
                    room = limit - used - piece_cost
                    if room < max(1, min(min_piece, nc - pos)):
                        if used == 0:
                            return None, None           # the limit does not even hold one piece
                        r, used = r + 1, 0
                        continue
                    take = min(room, nc - pos)
                out[r].append((c, pos * align, min((pos + take) * align, int(n_loci[c]))))
                used += piece_cost + take
                pos += take
                if r < world_size - 1 and used >= limit:
                    r, used = r + 1, 0
        return out, (used if r == world_size - 1 else 0)

    lo, hi = max(1, total // world_size), total + piece_cost * (len(chunks) + world_size)
    while lo < hi:   # smallest limit under which the last rank is not the most expensive
        mid = (lo + hi) // 2
        out, last = sweep(mid)
        if out is not None and last <= mid:
            hi = mid
        else:
            lo = mid + 1
    out, _ = sweep(lo)
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


class NcclComm:
    """This rank's own NCCL communicator, made with the NCCL library torch already loaded (ncclGetUniqueId on rank 0, the id
    broadcast through the torch.distributed group, ncclCommInitRank on every rank).  Engine.set_allreduce_nccl(comm.handle,
    comm.all_reduce_address) hands it to the CUDA library, which then calls ncclAllReduce itself: no Python, no host code at all in
    the generation loop, and the sharded control chain is captured into a CUDA graph with the collective inside."""

    class _UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_ubyte * 128)]

    def __init__(self, rank, world, device, group=None):
        import torch
        import torch.distributed as dist
        self.lib = ctypes.CDLL("libnccl.so.2")   # the soname torch's CUDA library is linked against: already in the process
        self.lib.ncclGetErrorString.restype = ctypes.c_char_p
        uid = self._UniqueId()
        if rank == 0:
            self._check(self.lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        box = [ctypes.string_at(ctypes.byref(uid), 128) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        torch.cuda.set_device(device)
        comm = ctypes.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, self._UniqueId, ctypes.c_int]
        self._check(self.lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank), "ncclCommInitRank")
        self.handle = comm.value
        self.all_reduce_address = ctypes.cast(self.lib.ncclAllReduce, ctypes.c_void_p).value

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.ncclGetErrorString(rc).decode()}")

    def close(self):
        if self.handle:
            self.lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self.lib.ncclCommDestroy(ctypes.c_void_p(self.handle))
            self.handle = None


def host_allreduce_hook(group=None):
    """Same for a host buffer (gloo) — used by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist

    def hook(ptr, count, stream):
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(count,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


def measure_sharded(name, args, steps, warmup, rank, world, local, e2e=True, n=None, loci=None):
    """One workload over all ranks of the process group; every rank returns the same dict (timings are max over ranks)."""
    import torch
    import torch.distributed as dist
    from . import capi, workloads
    import bench
    cfg = workloads.make_workload(name, n_override=n, loci_override=loci)
    warmup = bench.untimed_generations(warmup)
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
        seg_cap = int(2 * cap * (len(pieces) + (warmup + 2 * steps + 1) * morgans) * 1.05)
    kid = capi.GE_KERNEL_RECOMBINE_SEGMENTS if segs else capi.GE_KERNEL_PROPAGATE_BITS
    eng = capi.Engine(n_pop=len(pops), n_chr=len(pieces), n_phen=n_phen, device=local, representation=capi.GE_REP_SEGMENTS if segs else capi.GE_REP_BITS,
                      rng_mode=capi.GE_RNG_PHILOX, seed=12345, capacity=cap, seg_capacity=seg_cap, rank=rank, world_size=world)
    if len(pops) > 1:
        workloads.configure_engine_multipop(eng, cfg, pieces=pieces)
    else:
        workloads.configure_engine(eng, cfg, pieces=pieces)
    comm, collective = None, "hook"
    if getattr(args, "collective", "native") != "hook":
        # every rank must take the same path: a rank that cannot build its communicator (libnccl.so.2 not loadable by soname) says so first
        try:
            lib_ok = 1 if ctypes.CDLL("libnccl.so.2") else 0
        except OSError:
            lib_ok = 0
        flag = torch.tensor([lib_ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            comm = NcclComm(rank, world, local)
            eng.set_allreduce_nccl(comm.handle, comm.all_reduce_address)
            collective = "native"
    if comm is None:   # the Python hook into torch.distributed (also NCCL): measurement aid, or the NCCL library is not reachable from ctypes
        eng.set_allreduce(cuda_allreduce_hook(local))
    eng.init_generation0()
    gp = [capi.gen_params(q, cfg["mat_cor"], "p", "logit", 0.0, 1.0) for q in pops]
    mig = cfg.get("migration")
    gen = 0
    for _ in range(warmup):
        gen += 1
        eng.step_generation(gen, gp, mig)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
    out = {"ids": pin((cap, 7), torch.int64).view(np.uint64), "sex": pin((cap,), torch.uint8)}
    for k in "ADGCEFP":
        out[k] = pin((n_phen, cap), torch.float64)
    for k in ("mv", "sv", "svf"):
        out[k] = pin((cap,), torch.float64)

    def timed(with_download):
        nonlocal gen
        eng.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        work = 0
        eng.timer_start()
        for _ in range(steps):
            gen += 1
            eng.step_generation(gen, gp, mig)
            for q in range(len(pops)):
                work += eng.population_size(q) * M
                if with_download and rank == 0:
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
    r = dict(cfg=cfg, pieces=pieces, M=M, pops=pops, n_phen=n_phen, segs=segs, work=work, ms_dev=ms_dev, launches=launches, k_ms=k_ms, k_n=k_n, k_bytes=k_bytes,
             clocks=clocks.summary(), steps=steps, warmup=warmup)
    if e2e:
        r["work2"], r["ms_e2e"] = timed(True)
        r["checksum"] = float(out["P"].reshape(-1)[:16].sum()) if rank == 0 else 0.0
        # every rank simulated the same populations: the integer fingerprint (pedigree, sex, couples) must agree on all of them,
        # and with the unsharded run of the same number of generations (bench.py --gpus 1 prints the same field)
        hs = torch.tensor([bench.state_hash(eng, len(pops))], dtype=torch.int64, device="cuda")
        all_h = [torch.zeros_like(hs) for _ in range(world)]
        dist.all_gather(all_h, hs)
        r["hashes"] = [int(h.item()) for h in all_h]
    kb = torch.tensor([k_bytes / max(k_ms, 1e-9) / 1e6, k_ms / max(k_n, 1)], dtype=torch.float64, device="cuda")  # GB/s and ms per launch of this rank's kernel
    kmax = kb.clone()
    dist.all_reduce(kb, op=dist.ReduceOp.SUM)
    dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    r["achieved"] = float(kb[0].item()) / world  # mean per-GPU achieved GB/s
    r["k_ms_slowest"] = float(kmax[1].item())
    r["device_memory_gb"] = eng.device_memory_bytes() / 1e9
    r["graph_replays"] = eng.graph_replays()
    r["collective"] = collective
    eng.close()
    if comm:
        comm.close()
    return r


def sharded_roofline(r):
    import bench
    peak, peak_src = bench.hbm_peak()
    return {"bound": "hbm", "kernel": "seg_plan_kernel + seg_gather_kernel" if r["segs"] else "propagate_bits_kernel", "achieved": r["achieved"], "peak": peak, "unit": "GB/s",
            "frac": r["achieved"] / peak, "traffic": None, "traffic_source": "not captured for sharded runs (ncu is single-GPU; profiles/traffic.json holds the 1-GPU capture)",
            "peak_source": peak_src, "kernel_ms_per_launch": r["k_ms"] / max(r["k_n"], 1), "kernel_ms_per_launch_slowest_rank": r["k_ms_slowest"],
            "kernel_share_of_step": r["k_ms"] / r["ms_dev"], "algorithmic_bytes_per_launch_rank0": r["k_bytes"] // max(r["k_n"], 1), "note": "per-GPU mean"}


def bench_sharded(args, METRIC, UNIT):
    """bench.py for WORLD_SIZE > 1: the same fixed workload, its loci spread over the ranks (strong scaling).  Bit-packed rows are
    split by locus range (assign_locus_ranges: exact balance, a chromosome may span ranks), founder segments by whole chromosomes.
    From 4 ranks on, BASELINE config 4 (three populations, 150 GB of rows per generation: it needs them) is appended as a short record."""
    import torch
    import torch.distributed as dist
    from . import capi
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = measure_sharded(args.workload, args, args.steps, args.warmup, rank, world, local, n=args.n, loci=args.loci)
    cfg, M, pops, n_phen, segs = r["cfg"], r["M"], r["pops"], r["n_phen"], r["segs"]
    others = []
    if world >= 4 and args.workload == "config3_100k_x_1M" and not args.n and not args.loci and not args.no_other_workloads:
        try:
            o = measure_sharded("config4_3pop_300k_x_2M", args, 5, 3, rank, world, local, e2e=False)
            others.append({"workload": "config4_3pop_300k_x_2M", "individuals": sum(o["pops"]), "populations": o["pops"], "phenotypes": o["n_phen"], "loci": o["M"],
                           "steps": 5, "warmup": 3, "ms_per_step": o["ms_dev"] / 5, "value": o["work"] / (o["ms_dev"] * 1e-3), "unit": UNIT,
                           "gpu_launches_per_step": o["launches"] / 5, "roofline": sharded_roofline(o), "device_memory_gb_rank0": o["device_memory_gb"],
                           "note": "three populations 150k / 100k / 50k, ring migration 2 %, two phenotypes, 2M loci: 150 GB of bit-packed rows per generation over the ranks"})
        except Exception as e:  # never take the headline down
            others.append({"workload": "config4_3pop_300k_x_2M", "failed": repr(e)})
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_base = bench.cpu_baseline_single(args, M)
    if rank == 0:
        line = {
            "metric": METRIC, "value": r["work"] / (r["ms_dev"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_dev"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-packed + f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "individuals": sum(pops), "populations": pops, "phenotypes": n_phen, "loci": M, "chromosomes": len(cfg["chrs"]),
                       "parallelism": ("chromosome-sharded x%d (founder segments)" if segs else "locus-range sharded x%d (every rank: all individuals, 1/N of the 16-byte chunks of the rows)") % world,
                       "device_memory_gb_rank0": r["device_memory_gb"],
                       "representation": "founder segments (loci nominal; steps are generations %d..%d)" % (bench.untimed_generations(args.warmup) + 1, bench.untimed_generations(args.warmup) + args.steps) if segs else "bit-packed haplotypes",
                       "untimed_generations_before_timing": bench.untimed_generations(args.warmup),
                       "pieces_rank0": r["pieces"], "collective": "all-reduce of 3 * n_phen * capacity doubles per population and generation (ncclAllReduce issued by the CUDA library on its control stream, inside the captured generation graph when the chain is graphable)",
                       "graph_replays_rank0": r["graph_replays"], "collective_path": r["collective"],
                       "l2": ("inputs larger than L2 (%.1f GB of parental rows per step per GPU)" % (sum(pops) * M / 4 / 1e9 / world)) if not segs else
                             "inputs larger than L2 (founder-segment lists, %.1f GB moved per step per GPU)" % (r["k_bytes"] / max(r["k_n"], 1) * 2 / 1e9)},
            "e2e": {"value": r["work2"] / (r["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 40 * world,
                    "d2h_bytes_per_step": sum(capi.Engine.individual_bytes(q, n_phen) for q in pops), "ms_per_step": r["ms_e2e"] / args.steps,
                    "checksum": r["checksum"], "state_hash": r["hashes"][0], "state_hash_equal_on_all_ranks": len(set(r["hashes"])) == 1,
                    "generations_simulated": bench.untimed_generations(args.warmup) + 2 * args.steps},
            "gpu_launches": r["launches"],
            "roofline": sharded_roofline(r),
            "clocks": r["clocks"], "cpu_baseline": cpu_base}
        if others:
            line["other_workloads"] = others
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
