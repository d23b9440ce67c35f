"""Re-basing the founder panel (-m gpu; SURVEY.md §8f-3, ge_rebase_founders): a CUDA run that re-bases twice must stay on the
ORACLE's trajectory — the oracle, like the reference, never re-bases (src/Simulation.cpp:2903-2958 only appends) — in haplotypes
(materialised from the re-based lists and the re-based panel), causal-variant alleles (carried planes and ras_find_cv on the lists),
every per-individual column, and, composed back through the kept history, the reference's own segment lists part for part."""
import numpy as np
import pytest

from geneevolve_b200 import capi
from golden_util import Golden
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu
FLOAT_KEYS = ["A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]
REPS = {"segments": capi.GE_REP_SEGMENTS, "bits+segments": capi.GE_REP_BITS | capi.GE_REP_SEGMENTS}


@pytest.mark.parametrize("rep", sorted(REPS))
@pytest.mark.parametrize("name", ["A_am_pois", "B_rm_mut", "D_two_pops", "F_three_pops_ring"])
def test_rebased_run_stays_on_the_oracle_trajectory(cuda_lib, name, rep):
    G = Golden(name)
    gpu = capi.Engine(cuda_lib, **G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=REPS[rep], capacity=G.philox_capacity()))
    cpu = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX))
    for e in (gpu, cpu):
        G.configure(e)
        e.init_generation0()
    n_gen = min(G.G, 6)
    for gen in range(1, n_gen + 1):
        gp = G.all_params(gen)
        gpu.step_generation(gen, gp, G.migration_row(gen))
        cpu.step_generation(gen, gp, G.migration_row(gen))
        if gen in (2, 4):
            gpu.rebase_founders(keep_history=True)
            for p in range(G.n_pop):
                s = gpu.segments(p, 0)   # every haplotype: one part that names itself
                n = gpu.population_size(p)
                assert len(s["seg"]) == 2 * n and np.array_equal(s["seg"][:, 2], np.arange(2 * n)) and np.all(s["seg"][:, 3] == p)
        for p in range(G.n_pop):
            a, b = gpu.individuals(p), cpu.individuals(p)
            assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["sex"], b["sex"])
            for k in FLOAT_KEYS:
                np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-11, err_msg=f"{name} gen {gen} pop {p} {k}")
            before = {(f, c): gpu.cv_alleles(p, f, c) for f in range(G.n_phen) for c in range(G.n_chr)}
            for c in range(G.n_chr):
                want = cpu.haplotypes(p, c)
                assert np.array_equal(gpu.haplotypes(p, c), want), f"{name} gen {gen} pop {p} chr {c}: haplotypes"
                assert np.array_equal(gpu.haplotypes_from_segments(p, c), want), f"{name} gen {gen} pop {p} chr {c}: haplotypes from the re-based lists"
                for f in range(G.n_phen):
                    assert np.array_equal(before[(f, c)], cpu.cv_alleles(p, f, c)), f"{name} gen {gen}: CV alleles"
                g0, ref = gpu.segments_gen0(p, c), cpu.segments(p, c)
                assert np.array_equal(g0["seg_off"], ref["seg_off"]) and np.array_equal(g0["seg"], ref["seg"]), f"{name} gen {gen} pop {p} chr {c}: lists against generation 0"
            gpu.recompute_cv_from_segments(p)   # ras_find_cv on the re-based lists and the re-based CV panel
            for (f, c), v in before.items():
                assert np.array_equal(gpu.cv_alleles(p, f, c), v), f"{name} gen {gen}: carried planes differ from ras_find_cv after a re-base"


def test_rebase_without_history_keeps_the_lists_short(cuda_lib):
    G = Golden("A_am_pois")
    kw = G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_SEGMENTS, capacity=G.philox_capacity())
    a, b = capi.Engine(cuda_lib, **kw), capi.Engine(cuda_lib, **kw)
    for e in (a, b):
        G.configure(e)
        e.init_generation0()
    gp = G.all_params(1)
    grown = []
    for gen in range(1, 13):
        a.step_generation(gen, gp)
        b.step_generation(gen, gp)
        if gen % 4 == 0:
            a.rebase_founders(keep_history=False)
        grown.append((sum(len(a.segments(0, c)["seg"]) for c in range(G.n_chr)), sum(len(b.segments(0, c)["seg"]) for c in range(G.n_chr))))
        for c in range(G.n_chr):
            assert np.array_equal(a.haplotypes(0, c), b.haplotypes(0, c))
        ia, ib = a.individuals(0), b.individuals(0)
        for k in FLOAT_KEYS:
            assert np.array_equal(ia[k], ib[k]), k
    assert grown[-1][1] > 2 * grown[2][1]                       # the reference's lists keep growing ...
    assert grown[11][0] == 2 * a.population_size(0) * G.n_chr   # ... a re-based run is back to one part per haplotype every fourth generation
    assert max(g[0] for g in grown) <= max(g[1] for g in grown[:4]) * 1.5
    with pytest.raises(capi.GeneEvolveError):
        a.segments_gen0(0, 0)   # the lineage to generation 0 was dropped
