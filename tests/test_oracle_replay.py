"""The CPU oracle in fixed-draw mode: fed the reference's exported draws it must land on the reference's
state again (this is the plumbing the GPU parity tests rely on, checked here without a GPU)."""
import numpy as np
import pytest

from geneevolve_b200 import capi
from golden_util import SCENARIOS, Golden
from oracle.oracle import OracleEngine


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_replay_reproduces_reference(name):
    G = Golden(name)
    eng = OracleEngine(**G.engine_kwargs(rng_mode=capi.GE_RNG_REPLAY))
    G.configure(eng)
    eng.init_generation0([G.draws0(p) for p in range(G.n_pop)])
    for gen in range(0, G.G + 1):
        if gen:
            G.step_replay(eng, gen)
        for p in range(G.n_pop):
            ind = eng.individuals(p)
            for k in ["ids", "sex", "A", "D", "G", "C", "E", "F", "P", "mv", "sv", "svf"]:
                assert np.array_equal(ind[k], G.g(gen, p, k)), f"{name} gen {gen} pop {p} {k}"
            for c in range(G.n_chr):
                assert np.array_equal(eng.haplotypes(p, c), G.g(gen, p, f"c{c}.hap"))
                assert np.array_equal(eng.segments(p, c)["seg"], G.g(gen, p, f"c{c}.seg"))
