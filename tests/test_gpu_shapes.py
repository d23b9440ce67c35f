"""CUDA vs oracle at the SHAPES of the BASELINE configurations (-m gpu): the golden scenarios are toy-sized (<= 1 000
individuals x 1 000 loci on one or two chromosomes), so every kernel is also checked bit for bit against the CPU oracle
on inputs with the structure of configs 2-5 — 22 autosomes with the b37-shaped 50 kb map (55 657 rows), hundreds of
thousands of loci, 1 000 causal variants, assortative mating + logit selection, three populations with ring
migration and two phenotypes, founder segments over ten generations — at population sizes the oracle follows in
seconds.  Same Philox streams on both sides (src/Simulation.cpp:1890-2082 is the loop being compared)."""
import numpy as np
import pytest

from geneevolve_b200 import capi, workloads
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu
FLOAT_KEYS = ["A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]


def compare_population(gpu, cpu, p, what, haplotypes=True, segments=False):
    a, b = gpu.individuals(p), cpu.individuals(p)
    assert np.array_equal(a["ids"], b["ids"]), f"{what}: ids"
    assert np.array_equal(a["sex"], b["sex"]), f"{what}: sex"
    for k in FLOAT_KEYS:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-11, err_msg=f"{what}: {k}")
    for c in range(gpu.n_chr):
        if haplotypes:
            assert np.array_equal(gpu.haplotypes(p, c), cpu.haplotypes(p, c)), f"{what}: haplotypes of chromosome {c}"
        if segments:
            sa, sb = gpu.segments(p, c), cpu.segments(p, c)
            assert np.array_equal(sa["seg_off"], sb["seg_off"]) and np.array_equal(sa["seg"], sb["seg"]), f"{what}: segment lists of chromosome {c}"
        for f in range(gpu.n_phen):
            assert np.array_equal(gpu.cv_alleles(p, f, c), cpu.cv_alleles_bits(p, f, c)), f"{what}: CV alleles ({f}, {c})"


def compare_draws(gpu, cpu, p, what, sex=True):
    a, b = gpu.get_couples(p), cpu.get_couples(p)
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{what}: couples {k}"
    da, db = gpu.draws(p), cpu.draws(p)
    for k in ("father", "mother", "xo_off", "xo_bp", "start_hap") + (("sex",) if sex else ()):
        assert np.array_equal(da[k], db[k]), f"{what}: draw {k}"


def test_config3_shape_22_autosomes_assortative_selection(cuda_lib):
    """2 000 individuals x 22 autosomes x 200 000 loci, rho = 0.4, logit selection, Poisson family sizes, 3 generations."""
    cfg = workloads.make_workload("config3_100k_x_1M", n_override=2000, loci_override=200000, founders_override=1500)
    kw = dict(n_pop=1, n_chr=22, n_phen=1, rng_mode=capi.GE_RNG_PHILOX, seed=12345, capacity=3000, representation=capi.GE_REP_BITS)
    gpu, cpu = capi.Engine(cuda_lib, **kw), OracleEngine(**kw)
    for e in (gpu, cpu):
        workloads.configure_engine(e, cfg)
        e.init_generation0()
    compare_population(gpu, cpu, 0, "generation 0")
    gp = [capi.gen_params(2000, cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    for gen in range(1, 4):
        gpu.step_generation(gen, gp)
        cpu.step_generation(gen, gp)
        compare_draws(gpu, cpu, 0, f"generation {gen}")
        compare_population(gpu, cpu, 0, f"generation {gen}")


def test_config2_shape_chr22_half_a_million_loci(cuda_lib):
    """chr22 alone (684 map rows), 500 000 loci (rows of 62.5 KB, 4 tiles per gamete), 1 000 individuals, 2 generations."""
    cfg = workloads.make_workload("config2_chr22_10k", n_override=1000, founders_override=600)
    kw = dict(n_pop=1, n_chr=1, n_phen=1, rng_mode=capi.GE_RNG_PHILOX, seed=777, capacity=1600, representation=capi.GE_REP_BITS)
    gpu, cpu = capi.Engine(cuda_lib, **kw), OracleEngine(**kw)
    for e in (gpu, cpu):
        workloads.configure_engine(e, cfg)
        e.init_generation0()
    gp = [capi.gen_params(1000, 0.0, "p", "logit", 0.0, 1.0)]
    for gen in range(1, 3):
        gpu.step_generation(gen, gp)
        cpu.step_generation(gen, gp)
        compare_draws(gpu, cpu, 0, f"generation {gen}")
        compare_population(gpu, cpu, 0, f"generation {gen}")


def test_config4_shape_three_populations_ring_migration_two_phenotypes(cuda_lib):
    """Three populations of different sizes (1 500 / 1 000 / 500), ring migration 2 %, two phenotypes, 22 autosomes x 100 000
    loci, 3 generations: the row map after a migration, the per-population founder panels and the second phenotype's CV
    blocks all at once."""
    cfg = workloads.make_workload("config4_3pop_300k_x_2M", n_override=1500, loci_override=100000, founders_override=800)
    cfg["pops"] = [1500, 1000, 500]
    kw = dict(n_pop=3, n_chr=22, n_phen=2, rng_mode=capi.GE_RNG_PHILOX, seed=4242, capacity=2200, representation=capi.GE_REP_BITS)
    gpu, cpu = capi.Engine(cuda_lib, **kw), OracleEngine(**kw)
    for e in (gpu, cpu):
        workloads.configure_engine_multipop(e, cfg)
        e.init_generation0()
    gp = [capi.gen_params(n, cfg["mat_cor"], "p", "logit", 0.0, 1.0) for n in cfg["pops"]]
    for gen in range(1, 4):
        gpu.step_generation(gen, gp, cfg["migration"])
        cpu.step_generation(gen, gp, cfg["migration"])
        for p in range(3):
            assert gpu.population_size(p) == cpu.population_size(p)
            compare_draws(gpu, cpu, p, f"generation {gen} population {p}", sex=False)   # `sex` of the draws is post-migration on one side only
            compare_population(gpu, cpu, p, f"generation {gen} population {p}")


@pytest.mark.parametrize("seg_capacity", [0, 6_000_000])
def test_config5_shape_founder_segments_ten_generations(cuda_lib, seg_capacity):
    """The segment representation at config 5's shape (22 autosomes, no founder panel, loci nominal), 1 200 individuals, 10
    generations: every list of every chromosome must equal the oracle's restatement of `recombine` (:2903-2958) — with the
    output buffer growing (seg_capacity 0, control stream) and sized once (bulk stream)."""
    cfg = workloads.make_workload("config5_1M_x_10M_segments", n_override=1200, founders_override=700)
    kw = dict(n_pop=1, n_chr=22, n_phen=1, rng_mode=capi.GE_RNG_PHILOX, seed=99, capacity=1800, representation=capi.GE_REP_SEGMENTS)
    gpu, cpu = capi.Engine(cuda_lib, seg_capacity=seg_capacity, **kw), OracleEngine(**kw)
    for e in (gpu, cpu):
        workloads.configure_engine(e, cfg)
        e.init_generation0()
    gp = [capi.gen_params(1200, cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    for gen in range(1, 11):
        gpu.step_generation(gen, gp)
        cpu.step_generation(gen, gp)
        if gen in (1, 2, 5, 10):
            compare_draws(gpu, cpu, 0, f"generation {gen}")
            compare_population(gpu, cpu, 0, f"generation {gen}", haplotypes=False, segments=True)
    # ras_find_cv literally (scan every part for every CV) agrees with the planes carried by crossover parity
    before = [gpu.cv_alleles(0, 0, c) for c in range(22)]
    gpu.recompute_cv_from_segments(0)
    for c in range(22):
        assert np.array_equal(gpu.cv_alleles(0, 0, c), before[c])
        assert np.array_equal(before[c], cpu.cv_alleles(0, 0, c))
