"""Fixed-draw parity of the CUDA path (-m gpu): given the REAL reference's exported parent pairs,
crossovers, start haplotypes, mutation hits, sex and N(0,1) draws (tests/golden/*.npz), the library must give
bit-exact haplotypes / causal-variant alleles / pedigree and genetic values + phenotypes within 1e-9
relative (north star: bit-exact integers, 1e-6 relative on genetic values) — every generation, through the
C-ABI.  The CPU oracle runs alongside as a second checker."""
import numpy as np
import pytest

from geneevolve_b200 import capi
from golden_util import SCENARIOS, Golden
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-9, 1e-11
FLOAT_KEYS = ["A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]


def step_replay(G, eng, gen):
    G.step_replay(eng, gen)


def compare_to_golden(G, eng, gen, pops=None):
    for p in pops if pops is not None else range(G.n_pop):
        ind = eng.individuals(p)
        assert eng.population_size(p) == int(G.g(gen, p, "n"))
        assert np.array_equal(ind["ids"], G.g(gen, p, "ids"))
        assert np.array_equal(ind["sex"], G.g(gen, p, "sex"))
        for k in FLOAT_KEYS:
            np.testing.assert_allclose(ind[k], G.g(gen, p, k), rtol=RTOL, atol=ATOL, err_msg=f"{G.name} gen {gen} pop {p} {k}")
        for c in range(G.n_chr):
            assert np.array_equal(eng.haplotypes(p, c), G.g(gen, p, f"c{c}.hap")), f"{G.name} gen {gen} chr {c}: haplotypes"


@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("rep", [capi.GE_REP_BITS, capi.GE_REP_BITS | capi.GE_REP_SEGMENTS, capi.GE_REP_SEGMENTS, "segments-verbatim-walk",
                                 "segments-wide-parts", "segments-bulk-stream", "both-bulk-stream"])
def test_replay_matches_reference(cuda_lib, name, rep):
    extra, fmt = {}, 8   # sorted genetic maps: packed 8-byte parts, unless the verbatim walk (which needs `en` in memory) or wide parts are asked for
    if isinstance(rep, str):
        # the segment path's default is plan + gather (seg_plan_kernel, seg_gather_kernel); the reference's loop verbatim (one thread per
        # gamete, what maps with unsorted crossover lists use) and plan + gather on the reference's 16-byte parts are checked the same way
        if rep == "segments-verbatim-walk":
            extra["flags"] = capi.GE_FLAG_SEG_VERBATIM
            fmt = 16
        if rep == "segments-wide-parts":
            extra["flags"] = capi.GE_FLAG_SEG_WIDE_PARTS
            fmt = 16
        if rep.endswith("bulk-stream"):   # with seg_capacity the whole chain is queued on the bulk stream, n_seg is read back later
            extra["seg_capacity"] = 400000
        rep = capi.GE_REP_SEGMENTS | (capi.GE_REP_BITS if rep.startswith("both") else 0)
    G = Golden(name)
    gpu = capi.Engine(cuda_lib, **G.engine_kwargs(rng_mode=capi.GE_RNG_REPLAY, representation=rep, **extra))
    cpu = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_REPLAY))
    for e in (gpu, cpu):
        G.configure(e)
        e.init_generation0([G.draws0(p) for p in range(G.n_pop)])
    compare_to_golden(G, gpu, 0)
    if rep & capi.GE_REP_SEGMENTS:
        assert gpu.segment_format() == fmt
    for p in range(G.n_pop):
        for f in range(G.n_phen):
            a, b = gpu.gen0_constants(p, f), cpu.gen0_constants(p, f)
            for k in a:
                assert a[k] == pytest.approx(b[k], rel=1e-12, abs=1e-14)
    for gen in range(1, G.G + 1):
        # the device mating kernels under the reference's own mating draws must arrive at the reference's couples
        # (src/Simulation.cpp:2090-2157 / :2167-2360); the generation itself then replays the reference's per-offspring draws
        for p in range(G.n_pop):
            gpu.mate_replay(p, gen, G.params(gen, p), **G.mate_draws(gen, p))
            c = gpu.get_couples(p)
            for k, key in (("pos_male", "couple_male"), ("pos_female", "couple_female"), ("inbreed", "couple_inbreed"), ("num_offspring", "couple_noff")):
                assert np.array_equal(c[k].astype(np.int64), G.g(gen, p, key).astype(np.int64)), f"{G.name} gen {gen} pop {p}: {key} under the reference's mating draws"
        step_replay(G, gpu, gen)
        step_replay(G, cpu, gen)
        compare_to_golden(G, gpu, gen)
        for p in range(G.n_pop):
            for c in range(G.n_chr):
                for f in range(G.n_phen):
                    assert np.array_equal(gpu.cv_alleles(p, f, c), cpu.cv_alleles(p, f, c)), "causal-variant alleles"
                if rep & capi.GE_REP_SEGMENTS:
                    s, r = gpu.segments(p, c), cpu.segments(p, c)
                    assert np.array_equal(s["seg_off"], r["seg_off"]) and np.array_equal(s["seg"], r["seg"]), "segments"
                    assert np.array_equal(s["seg"], G.g(gen, p, f"c{c}.seg"))
                    assert np.array_equal(s["mut_off"], r["mut_off"])
                    for k in range(len(s["mut_off"]) - 1):
                        assert np.array_equal(np.sort(s["mut_bp"][s["mut_off"][k]:s["mut_off"][k + 1]]), np.sort(r["mut_bp"][r["mut_off"][k]:r["mut_off"][k + 1]]))
            if rep & capi.GE_REP_SEGMENTS:
                # the planes the hot path carries forward by crossover parity == ras_find_cv (:2752-2815) on the parts
                before = [[gpu.cv_alleles(p, f, c) for c in range(G.n_chr)] for f in range(G.n_phen)]
                gpu.recompute_cv_from_segments(p)
                for f in range(G.n_phen):
                    for c in range(G.n_chr):
                        assert np.array_equal(gpu.cv_alleles(p, f, c), before[f][c]), "propagated CV planes differ from ras_find_cv on the segments"
            for f in range(G.n_phen):
                m, r = gpu.moments(p, f), cpu.moments(p, f)
                for k in m:
                    assert m[k] == pytest.approx(r[k], rel=1e-9, abs=1e-12)
