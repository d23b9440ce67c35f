"""Pins the CPU oracle (oracle/ge_oracle.cpp) against the REAL reference.

The oracle runs in reference-stream mode from `--seed` alone and must reproduce, bit for bit, everything
the reference exported into tests/golden/*.npz: couples, crossovers, start haplotypes, mutation hits, sex,
per-individual values, founder segments and the materialised haplotype matrix — every generation, every
population.  It also proves the representation claim the CUDA path rests on: the bit-packed propagation
equals ras_convert_interval_to_hap_matrix (src/Simulation.cpp:1186-1230) applied to the reference's segments.
"""
import numpy as np
import pytest

from golden_util import SCENARIOS, Golden
from oracle.oracle import GO_RNG_REF, OracleEngine

STATE_KEYS = ["ids", "sex", "A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]


def check_state(G, eng, gen):
    for p in range(G.n_pop):
        ind = eng.individuals(p)
        assert eng.population_size(p) == int(G.g(gen, p, "n"))
        for k in STATE_KEYS:
            assert np.array_equal(ind[k], G.g(gen, p, k)), f"{G.name} gen {gen} pop {p}: {k} differs"
        for f in range(G.n_phen):
            c = eng.gen0_constants(p, f)
            ref = G.g(gen, p, "var_a0_var_d0_beta")[f]
            assert (c["var_a0"], c["var_d0"], c["beta"]) == tuple(ref)
        for c in range(G.n_chr):
            s = eng.segments(p, c)
            assert np.array_equal(s["seg_off"], G.g(gen, p, f"c{c}.seg_off"))
            assert np.array_equal(s["seg"], G.g(gen, p, f"c{c}.seg"))
            # mutation lists: the reference keeps one list per part; compare per haplotype as multisets
            ref_off, ref_bp = G.g(gen, p, f"c{c}.segmut_off"), G.g(gen, p, f"c{c}.segmut_bp")
            seg_off = s["seg_off"]
            for r in range(len(seg_off) - 1):
                a = np.sort(s["mut_bp"][s["mut_off"][r]:s["mut_off"][r + 1]])
                b = np.sort(ref_bp[ref_off[seg_off[r]]:ref_off[seg_off[r + 1]]])
                assert np.array_equal(a, b)
            hap_ref = G.g(gen, p, f"c{c}.hap")
            assert np.array_equal(eng.haplotypes_from_segments(p, c), hap_ref), "segment materialisation differs"
            assert np.array_equal(eng.haplotypes(p, c), hap_ref), "bit-packed haplotypes differ"
            for f in range(G.n_phen):
                assert np.array_equal(eng.cv_alleles(p, f, c), eng.cv_alleles_bits(p, f, c))


def check_draws(G, eng, gen):
    for p in range(G.n_pop):
        cp = eng.get_couples(p)
        assert np.array_equal(cp["pos_male"], G.g(gen, p, "couple_male"))
        assert np.array_equal(cp["pos_female"], G.g(gen, p, "couple_female"))
        assert np.array_equal(cp["inbreed"], G.g(gen, p, "couple_inbreed"))
        assert np.array_equal(cp["num_offspring"], G.g(gen, p, "couple_noff").astype(np.int32))
        d = eng.draws(p)
        for k, gk in [("father", "off_father"), ("mother", "off_mother"), ("sex", "off_sex"), ("xo_off", "xo_off"),
                      ("xo_bp", "xo_bp"), ("start_hap", "start_hap"), ("mut_off", "mut_off")]:
            assert np.array_equal(d[k], G.g(gen, p, gk)), f"{G.name} gen {gen} pop {p}: draw {k} differs"
        mo = d["mut_off"]
        for s in range(len(mo) - 1):  # the reference lists hits part by part, the oracle in draw order
            a = sorted(zip(d["mut_bp"][mo[s]:mo[s + 1]], d["mut_gam"][mo[s]:mo[s + 1]]))
            b = sorted(zip(G.g(gen, p, "mut_bp")[mo[s]:mo[s + 1]], G.g(gen, p, "mut_gam")[mo[s]:mo[s + 1]]))
            assert a == b
        assert np.array_equal(eng.e_raw(p), G.g(gen, p, "e_raw"))


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_stream_reproduces_reference(name):
    G = Golden(name)
    eng = OracleEngine(**G.engine_kwargs(rng_mode=GO_RNG_REF))
    G.configure(eng)
    eng.init_generation0()
    check_state(G, eng, 0)
    for gen in range(1, G.G + 1):
        if G.n_pop == 1:
            eng.step_generation(gen, G.all_params(gen), G.migration_row(gen))
            check_draws(G, eng, gen)
        else:
            # step by hand so the draws can be read before migration reorders the populations
            for p in range(G.n_pop):
                eng.mate(p, gen, G.params(gen, p))
                eng.reproduce(p, gen)
                eng.compute_AD(p, gen)
                for f in range(G.n_phen):
                    eng.scale_AD_compute_GEF(p, gen, f)
                check_draws_pop = p
            check_draws(G, eng, gen)
            for f in range(G.n_phen):
                eng.environmental_effects_specific_to_each_population(f)
            for p in range(G.n_pop):
                eng.compute_mating_value_selection_value(p, gen, G.params(gen, p))
            eng.do_migration(gen, G.migration_row(gen))
            for p in range(G.n_pop):
                eng.save_human_info_to_Pop_info_prev_gen(p)
        check_state(G, eng, gen)
