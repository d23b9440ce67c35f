"""GPU-RNG mode (-m gpu): the device Philox streams, the skip-sampler for crossovers/mutations and the
device mating kernels must reproduce the CPU oracle's restatement of the same streams bit for bit —
couples, crossovers, start haplotypes, mutation hits, sex, haplotypes, causal-variant alleles — and the fp64
columns to 1e-9, over every generation of every single-population golden configuration."""
import numpy as np
import pytest

from geneevolve_b200 import capi
from golden_util import SCENARIOS, Golden
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu
FLOAT_KEYS = ["A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]


def covered_hits(G, d, pop):
    """Mutation hits per (offspring, chromosome) restricted to the covered range, as sorted tuples."""
    out = []
    mo = d["mut_off"]
    for s in range(len(mo) - 1):
        c = s % G.n_chr
        bp = G[f"in.p{pop}.c{c}.rmap_bp"]
        hits = [(int(b), int(g)) for b, g in zip(d["mut_bp"][mo[s]:mo[s + 1]], d["mut_gam"][mo[s]:mo[s + 1]]) if bp[0] <= b < bp[-1]]
        out.append(sorted(hits))
    return out


@pytest.mark.parametrize("name", SCENARIOS)
def test_philox_generation_matches_oracle(cuda_lib, name):
    G = Golden(name)
    gpu = capi.Engine(cuda_lib, **G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS, capacity=G.philox_capacity()))
    cpu = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX))
    for e in (gpu, cpu):
        G.configure(e)
        e.init_generation0()
    for gen in range(0, G.G + 1):
        if gen:
            try:
                gpu.step_generation(gen, G.all_params(gen), G.migration_row(gen))
            except capi.GeneEvolveError as e:
                # tiny inbreeding-avoiding populations can run out of marriageable couples: the oracle must
                # stop in the same generation with the same error
                with pytest.raises(capi.GeneEvolveError) as e2:
                    cpu.step_generation(gen, G.all_params(gen), G.migration_row(gen))
                assert e2.value.code == e.code
                return
            cpu.step_generation(gen, G.all_params(gen), G.migration_row(gen))
            for p in range(G.n_pop):
                a, b = gpu.get_couples(p), cpu.get_couples(p)
                for k in a:
                    assert np.array_equal(a[k], b[k]), f"{name} gen {gen}: couples {k}"
                da, db = gpu.draws(p), cpu.draws(p)
                for k in ("father", "mother", "xo_off", "xo_bp", "start_hap") + (("sex",) if G.n_pop == 1 else ()):
                    assert np.array_equal(da[k], db[k]), f"{name} gen {gen}: draw {k}"
                assert covered_hits(G, da, p) == covered_hits(G, db, p)
        for p in range(G.n_pop):
            a, b = gpu.individuals(p), cpu.individuals(p)
            assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["sex"], b["sex"])
            for k in FLOAT_KEYS:
                np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-11, err_msg=f"{name} gen {gen} {k}")
            for c in range(G.n_chr):
                assert np.array_equal(gpu.haplotypes(p, c), cpu.haplotypes(p, c)), f"{name} gen {gen} chr {c}"
                for f in range(G.n_phen):
                    assert np.array_equal(gpu.cv_alleles(p, f, c), cpu.cv_alleles(p, f, c))


@pytest.mark.parametrize("name,rep,seg_capacity", [("A_am_pois", capi.GE_REP_BITS, 0), ("G_bundled_example_chr1", capi.GE_REP_BITS, 0),
                                                   ("A_am_pois", capi.GE_REP_SEGMENTS, 400_000), ("C_fixed_mm", capi.GE_REP_BITS | capi.GE_REP_SEGMENTS, 400_000)])
def test_graph_replay_equals_plain_launches(cuda_lib, name, rep, seg_capacity):
    """The control chain of a generation replayed as a captured CUDA graph (the default once a buffer-parity key is warm: generation 5
    onwards) against the same context created with GE_FLAG_NO_GRAPH, which queues every kernel: identical couples, draws, haplotypes,
    segments and columns over 14 generations — the graphs really replayed (far fewer launches counted by the host is not the check:
    ge_get_launch_count counts a graph's kernels too) — and against the oracle at the end."""
    G = Golden(name)
    kw = G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=rep, seg_capacity=seg_capacity, capacity=G.philox_capacity())
    a, b = capi.Engine(cuda_lib, **kw), capi.Engine(cuda_lib, flags=capi.GE_FLAG_NO_GRAPH, **kw)
    cpu = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX))
    for e in (a, b, cpu):
        G.configure(e)
        e.init_generation0()
    gp = G.all_params(1)
    for gen in range(1, 15):
        for e in (a, b, cpu):
            e.step_generation(gen, gp)
        ca, cb = a.get_couples(0), b.get_couples(0)
        for k in ca:
            assert np.array_equal(ca[k], cb[k]), f"gen {gen}: couples {k}"
        da, db = a.draws(0), b.draws(0)
        for k in da:
            assert np.array_equal(da[k], db[k]), f"gen {gen}: draw {k}"
        ia, ib = a.individuals(0), b.individuals(0)
        for k in ia:
            assert np.array_equal(ia[k], ib[k]), f"gen {gen}: column {k}"
        for c in range(G.n_chr):
            assert np.array_equal(a.haplotypes(0, c), b.haplotypes(0, c)), f"gen {gen} chr {c}"
            if rep & capi.GE_REP_SEGMENTS:
                sa, sb = a.segments(0, c), b.segments(0, c)
                assert np.array_equal(sa["seg_off"], sb["seg_off"]) and np.array_equal(sa["seg"], sb["seg"])
    assert a.launch_count() == b.launch_count()
    ic = cpu.individuals(0)
    assert np.array_equal(ia["ids"], ic["ids"]) and np.array_equal(ia["sex"], ic["sex"])
    for k in FLOAT_KEYS:
        np.testing.assert_allclose(ia[k], ic[k], rtol=1e-9, atol=1e-11, err_msg=k)
    for c in range(G.n_chr):
        assert np.array_equal(a.haplotypes(0, c), cpu.haplotypes(0, c))
