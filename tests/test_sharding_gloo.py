"""Multi-rank logic on CPU (world_size 2, gloo): chromosome-sharded engines + the sum-allreduce hook must
reproduce the single-rank run — identical couples, draws and haplotypes (Philox counters carry global
chromosome ids), fp64 columns to 1e-10 (the all-reduce changes the summation order of the chromosomes).
The engines are CPU oracles standing in for the CUDA contexts: the host-side sharding logic (assignment,
global ids, hook wiring, replicated mating) is the same code."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from geneevolve_b200 import capi, dist as gdist
from golden_util import Golden
from oracle.oracle import OracleEngine


def configure_subset(G, eng, chrs):
    z = G.z
    eng.set_chromosome_ids(chrs)
    for k, c in enumerate(chrs):
        eng.set_loci(k, z[f"in.p0.c{c}.panel_pos"])
    for p in range(G.n_pop):
        pre = f"in.p{p}."
        eng.set_population(p, bool(z[pre + "avoid_inbreeding"]), bool(z[pre + "RM"]), float(z[pre + "MM_percent"]))
        for k, c in enumerate(chrs):
            cp = pre + f"c{c}."
            eng.set_genetic_map(p, k, z[cp + "rmap_bp"], z[cp + "recom_prob"], int(z[cp + "bp_dist"]))
            if int(z[pre + "has_mutation_map"]):
                eng.set_mutation_map(p, k, z[cp + "mut_bp"], z[cp + "mut_rate"])
            eng.set_founder_panel(p, k, z[cp + "panel"])
            for f in range(G.n_phen):
                fp = cp + f"f{f}."
                eng.set_cv(p, f, k, z[fp + "cv_bp"], z[fp + "cv_a"], z[fp + "cv_d"], z[fp + "cv_val"])
        for f in range(G.n_phen):
            s = z[pre + "scheme"][f]
            eng.set_pheno_scheme(p, f, va=s[0], vd=s[1], ve=s[2], vc=s[3], vf=s[4], omega=s[5], beta=s[6], lam=s[7])
    if len(z["in.gamma"]):
        eng.set_gamma(z["in.gamma"])


def configure_pieces(G, eng, pieces):
    """Locus-range sharding (dist.assign_locus_ranges): piece (c, s0, s1) holds loci [s0, s1) of chromosome c with the whole genetic
    (and mutation) map of c; a causal variant goes to the piece whose locus range covers its position."""
    z = G.z
    eng.set_chromosome_ids([c for c, _, _ in pieces])
    bounds = []
    for k, (c, s0, s1) in enumerate(pieces):
        pos = z[f"in.p0.c{c}.panel_pos"]
        eng.set_loci(k, pos[s0:s1])
        bounds.append((-1 if s0 == 0 else int(pos[s0]), 2 ** 62 if s1 >= len(pos) else int(pos[s1])))
    for p in range(G.n_pop):
        pre = f"in.p{p}."
        eng.set_population(p, bool(z[pre + "avoid_inbreeding"]), bool(z[pre + "RM"]), float(z[pre + "MM_percent"]))
        for k, (c, s0, s1) in enumerate(pieces):
            cp = pre + f"c{c}."
            eng.set_genetic_map(p, k, z[cp + "rmap_bp"], z[cp + "recom_prob"], int(z[cp + "bp_dist"]))
            if int(z[pre + "has_mutation_map"]):
                eng.set_mutation_map(p, k, z[cp + "mut_bp"], z[cp + "mut_rate"])
            eng.set_founder_panel(p, k, np.ascontiguousarray(z[cp + "panel"][:, s0:s1]))
            for f in range(G.n_phen):
                fp = cp + f"f{f}."
                keep = (z[fp + "cv_bp"].astype(np.int64) >= bounds[k][0]) & (z[fp + "cv_bp"].astype(np.int64) < bounds[k][1])
                eng.set_cv(p, f, k, z[fp + "cv_bp"][keep], z[fp + "cv_a"][keep], z[fp + "cv_d"][keep], np.ascontiguousarray(z[fp + "cv_val"][:, keep]))
        for f in range(G.n_phen):
            s = z[pre + "scheme"][f]
            eng.set_pheno_scheme(p, f, va=s[0], vd=s[1], ve=s[2], vc=s[3], vf=s[4], omega=s[5], beta=s[6], lam=s[7])
    if len(z["in.gamma"]):
        eng.set_gamma(z["in.gamma"])


def shard_of(G, split, world):
    """Per rank: the pieces (chromosome, first locus, end locus) under either split."""
    n_loci = [len(G[f"in.p0.c{c}.panel_pos"]) for c in range(G.n_chr)]
    if split == "locus-tiles":
        return gdist.assign_locus_ranges(n_loci, world, align=32)
    return [[(c, 0, n_loci[c]) for c in part] for part in gdist.assign_chromosomes(n_loci, world)]


def run_generations(G, eng, n_gen, segments=False):
    """State of the LAST population after n_gen generations (with migration every population feeds into it)."""
    eng.init_generation0()
    for gen in range(1, n_gen + 1):
        eng.step_generation(gen, G.all_params(gen), G.migration_row(gen))
    p = G.n_pop - 1
    out = {"couples": eng.get_couples(p), "ind": eng.individuals(p)}
    out["hap"] = [eng.haplotypes(p, k) for k in range(eng.n_chr)]
    if segments:
        out["seg"] = [eng.segments(p, k) for k in range(eng.n_chr)]
    return out


def worker(rank, world, port, name, n_gen, q, split):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = Golden(name)
    mine = shard_of(G, split, world)[rank]
    kw = G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX)
    kw.update(n_chr=len(mine), rank=rank, world_size=world)
    eng = OracleEngine(**kw)
    configure_pieces(G, eng, mine)
    eng.set_allreduce(gdist.host_allreduce_hook())
    out = run_generations(G, eng, n_gen)
    q.put((rank, mine, out))
    dist.barrier()
    dist.destroy_process_group()


def test_assign_chromosomes_balances():
    w = [249, 243, 198, 191, 181, 171, 159, 146, 141, 135, 135, 133, 115, 107, 102, 90, 81, 78, 59, 63, 48, 51]
    parts = gdist.assign_chromosomes(w, 8)
    assert sorted(c for p in parts for c in p) == list(range(22))
    loads = [sum(w[c] for c in p) for p in parts]
    assert max(loads) <= 1.08 * sum(w) / 8


def test_assign_locus_ranges_balances_and_covers():
    n_loci = [86555, 84425, 68745, 66366, 62807, 59403, 55236, 50808, 49019, 47040, 46866, 46467, 39972, 37246, 35579, 31360, 28165, 27088, 20507, 21861, 16687, 17798]
    total = sum((n + 127) // 128 for n in n_loci)
    for world in (1, 2, 3, 8, 22, 40):
        for piece_cost in (0, gdist.piece_cost_chunks(total / world)):
            parts = gdist.assign_locus_ranges(n_loci, world, piece_cost=piece_cost)
            seen = [[] for _ in n_loci]
            loads, costs = [], []
            for pieces in parts:
                loads.append(sum((s1 - s0 + 127) // 128 for _, s0, s1 in pieces))
                costs.append(loads[-1] + piece_cost * len(pieces))
                for c, s0, s1 in pieces:
                    assert s0 % 128 == 0 and s0 < s1 <= n_loci[c]
                    seen[c].append((s0, s1))
            if piece_cost == 0:
                assert max(loads) == -(-total // world)              # chunks of 16 bytes: no rank above the even share
            else:                                                    # the most expensive rank is within one piece of the even share
                assert max(costs) <= (total + piece_cost * (len(n_loci) + world - 1)) / world + piece_cost
            for c, iv in enumerate(seen):                            # every locus on exactly one rank
                iv.sort()
                assert iv[0][0] == 0 and iv[-1][1] == n_loci[c] and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
    eight = gdist.assign_locus_ranges(n_loci, 8)                     # the rank with the six short chromosomes holds fewer loci than the rank with two long ones
    assert len(eight[0]) < len(eight[7]) and sum(s1 - s0 for _, s0, s1 in eight[0]) > sum(s1 - s0 for _, s0, s1 in eight[7])
    assert [len(p) for p in gdist.assign_locus_ranges([500000], 4)] == [1, 1, 1, 1]   # one chromosome spreads over every rank


@pytest.mark.parametrize("split", ["chromosomes", "locus-tiles"])
@pytest.mark.parametrize("name", ["B_rm_mut", "A_am_pois", "D_two_pops"])
def test_two_ranks_match_single_rank(name, split):
    G = Golden(name)
    n_gen = min(G.G, 3)
    single = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX))
    G.configure(single)
    ref = run_generations(G, single, n_gen)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, name, n_gen, q, split)) for r in range(2)]
    [p.start() for p in procs]
    results = [q.get(timeout=300) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    covered = []
    for rank, mine, out in results:
        for k in ("pos_male", "pos_female", "inbreed", "num_offspring"):
            assert np.array_equal(out["couples"][k], ref["couples"][k]), f"rank {rank}: couples differ"
        assert np.array_equal(out["ind"]["ids"], ref["ind"]["ids"]) and np.array_equal(out["ind"]["sex"], ref["ind"]["sex"])
        for k in "ADGCEFP":
            np.testing.assert_allclose(out["ind"][k], ref["ind"][k], rtol=1e-10, atol=1e-12)
        for k, (c, s0, s1) in enumerate(mine):
            assert np.array_equal(out["hap"][k], ref["hap"][c][:, s0:s1]), f"rank {rank}: chromosome {c} loci [{s0}, {s1}) differ"
            covered.append((c, s0, s1))
    for c in range(G.n_chr):   # the pieces tile every chromosome
        iv = sorted((s0, s1) for cc, s0, s1 in covered if cc == c)
        assert iv[0][0] == 0 and iv[-1][1] == ref["hap"][c].shape[1] and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
