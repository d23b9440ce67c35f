"""Statistical parity under the GPU's own RNG (north star, second correctness mode).

The library's Philox streams cannot reproduce the reference's minstd_rand0/rand() draws, so in GE_RNG_PHILOX mode
agreement is statistical: over 16 replicate runs per scenario the means of
  * the heterozygosity trajectory and the squared allele-frequency drift (doc §3.2-3.3),
  * the decay of LD with genetic distance (regression slope of D_t on D_0 in four recombination-fraction bins, §3.4),
  * realised heritability var(A)/var(P), var(A), var(P) (§3.5: inflation under assortative mating rho = 0.4),
  * the realised spouse correlation of mating values, the number of couples and of offspring,
must agree with the same statistics of 16 replicate runs of the REAL reference (tests/golden/stats_reference.json,
made by tests/golden/make_stats_reference.py) within 4.5 standard errors of the difference plus 2 % relative +
1e-3 absolute slack (tests/stats_util.compare).  Heterozygosity is also held to theory, (1 - 1/2N)^t.

The CUDA library is checked with -m gpu; the CPU oracle's restatement of the same Philox streams is checked
without a GPU, which pins the stream design itself (sampler laws, mating scheme) on every CPU run of the suite.
"""
import json
import os

import numpy as np
import pytest

import stats_util as su
from geneevolve_b200 import capi
from golden_util import Golden

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "stats_reference.json")) as f:
    REFERENCE = json.load(f)


def run_replicate(make_engine, name, seed):
    sc = su.SCENARIOS[name]
    G = Golden("stats_inputs_" + name)
    cap = int(max(r[0] for r in sc["gens"]) * 1.4) + 64
    eng = make_engine(n_pop=1, n_chr=G.n_chr, n_phen=1, vt_type=G.vt_type, seed=seed, capacity=cap, rng_mode=capi.GE_RNG_PHILOX,
                      representation=capi.GE_REP_BITS)
    G.configure(eng)
    eng.init_generation0()
    T = su.Trajectory(sc)
    for gen in range(len(sc["gens"]) + 1):
        if gen:
            mv = eng.individuals(0)["mv"]
            eng.step_generation(gen, G.all_params(gen), None)
            cp = eng.get_couples(0)
            T.add_couples(mv, cp["pos_male"], cp["pos_female"], cp["num_offspring"])
        ind = eng.individuals(0)
        T.add_generation(gen, [eng.haplotypes(0, c) for c in range(G.n_chr)], ind["A"][0], ind["P"][0])
    return T.vector()


def check_scenario(make_engine, name):
    assert REFERENCE["seeds"] == su.SEEDS
    ref_runs = REFERENCE["scenarios"][name]
    runs = [run_replicate(make_engine, name, seed) for seed in su.SEEDS]
    for k in runs[0]:
        su.compare([r[k] for r in ref_runs], [r[k] for r in runs], f"{name}:{k}")
    sc = su.SCENARIOS[name]
    if sc["rm"]:  # drift theory: H_t = H_0 * prod (1 - 1/(2 N_{t-1})), N_0 = founders
        sizes = [sc["n_founders"]] + [r[0] for r in sc["gens"]]
        theory = np.cumprod([1.0] + [1 - 1 / (2 * n) for n in sizes[:-1]])
        het = np.array([r["het"] for r in runs])
        se = het.std(axis=0, ddof=1) / np.sqrt(het.shape[0])
        assert np.all(np.abs(het.mean(axis=0) - theory) <= 4.5 * se + 0.01), (het.mean(axis=0), theory)
    if sc["ld_gens"]:  # LD decays faster at larger genetic distance and with time
        s = np.array([[r[f"ld_slope_gen{g}"] for g in sc["ld_gens"]] for r in runs]).mean(axis=0)  # [gen][bin]
        assert np.all(np.diff(s, axis=0) < 0.02) and np.all(np.diff(s, axis=1) < 0.02), s


@pytest.mark.parametrize("name", sorted(su.SCENARIOS))
def test_oracle_philox_statistics_match_reference(name):
    from oracle.oracle import OracleEngine
    check_scenario(lambda **kw: OracleEngine(**kw), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(su.SCENARIOS))
def test_gpu_philox_statistics_match_reference(cuda_lib, name):
    check_scenario(lambda **kw: capi.Engine(cuda_lib, **kw), name)
