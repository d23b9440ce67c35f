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
    if sc["rm"]:  # drift theory: every generation draws 2 N_t gametes from the parental gene pool, H_t = H_0 * prod (1 - 1/(2 N_t))
        theory = np.cumprod([1.0] + [1 - 1 / (2 * r[0]) for r in sc["gens"]])
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


@pytest.mark.gpu
def test_gpu_heterozygosity_decay_at_the_documented_scale(cuda_lib):
    """The reference's own validation recipe (GeneEvolveDocumentation.pdf §3.3, Table 3.2): N = 33 253 per generation,
    random mating, chr22, 100 generations; mean heterozygosity must follow h(t) = h(0) * prod(1 - 1/2N) and the per-SNP
    squared deviation from that curve must be what Wright-Fisher drift predicts (the documentation reports a mean MSE of
    1.11e-4 in its setting)."""
    from geneevolve_b200 import workloads
    N, G, n_snp, n_founders = 33253, 100, 5000, 2000
    rng = np.random.default_rng(33253)
    (bp, cm, p), = workloads.genetic_map([22])
    pos = np.sort(rng.choice(np.arange(int(bp[0]), int(bp[-1])), size=n_snp, replace=False)).astype(np.uint64)
    freq = rng.uniform(0.1, 0.9, n_snp)
    panel = (rng.uniform(size=(2 * n_founders, n_snp)) < freq[None, :]).astype(np.uint8)
    cv_idx = np.sort(rng.choice(n_snp, size=50, replace=False))
    eng = capi.Engine(cuda_lib, n_pop=1, n_chr=1, n_phen=1, seed=2018, capacity=N + 64, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS)
    eng.set_loci(0, pos)
    eng.set_population(0, False, True, 0.0)
    eng.set_genetic_map(0, 0, bp, p, int(bp[1] - bp[0]))
    eng.set_founder_panel(0, 0, panel)
    eng.set_cv(0, 0, 0, pos[cv_idx], rng.normal(size=50), np.zeros(50), panel[:, cv_idx])
    eng.set_pheno_scheme(0, 0, va=0.5, vd=0.0, ve=0.5)
    eng.init_generation0()

    def het():
        w = eng.haplotypes_packed(0, 0)
        bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, :n_snp]
        q = bits.mean(axis=0)
        return 2 * q * (1 - q)

    h0 = het()
    q0 = (1 - np.sqrt(1 - 2 * h0)) / 2                      # allele frequencies at generation 0 (minor allele)
    gp = [capi.gen_params(N, 0.0, "p", "thr", 1, 1)]
    theory, drift = 1.0, 0.0
    for gen in range(1, G + 1):
        theory *= 1 - 1 / (2 * N)                            # 2N gametes are drawn from the parental gene pool
        drift += 1 / (2 * N)                                # Var(p_t) ~ p q * sum 1/(2 N_k)
        eng.step_generation(gen, gp)
        assert eng.population_size(0) == N
        if gen in (25, 50, 100):
            h = het()
            assert abs(h.mean() / h0.mean() - theory) < 0.002, (gen, h.mean() / h0.mean(), theory)
            # per-SNP squared deviation from the curve (the documentation's Table 3.2 statistic, 1.11e-4 in its setting):
            # Wright-Fisher drift predicts Var(H_t) ~ (dH/dp)^2 Var(p_t) = 4 (1 - 2p)^2 p q * sum 1/(2 N_k)
            mse, expect = np.mean((h - theory * h0) ** 2), np.mean(4 * (1 - 2 * q0) ** 2 * q0 * (1 - q0)) * drift
            assert 0.7 * expect < mse < 1.3 * expect and mse < 3e-4, (gen, mse, expect)


@pytest.mark.gpu
def test_gpu_ld_decay_follows_the_genetic_map(cuda_lib):
    """LD decay versus genetic distance against theory (doc §3.4 recipe, here with a closed form instead of PLINK): under
    random mating D_t(x, y) = D_0(x, y) (1 - c_xy)^t (1 - 1/2N)^t, where c_xy is the recombination fraction between the two
    SNPs under the law the sampler implements — row j of the map fires with probability p_j and its crossover lands
    uniformly in [bp_j, bp_j + bp_dist) (the reference's position/probability offset).  The b37-shaped
    chr22 map has hot and cold spots, so rates AND positions of the crossovers have to be right for the binned regression
    slopes of D_t on D_0 to follow the prediction."""
    from geneevolve_b200 import workloads
    N, n_snp, n_founders = 20000, 1500, 2000
    rng = np.random.default_rng(2204)
    (bp, cm, p), = workloads.genetic_map([22])
    dist = int(bp[1] - bp[0])
    pos = np.sort(rng.choice(np.arange(int(bp[1]), int(bp[-2])), size=n_snp, replace=False)).astype(np.uint64)
    anc = (rng.uniform(size=(8, n_snp)) < rng.uniform(0.25, 0.75, n_snp)[None, :]).astype(np.uint8)
    panel = np.zeros((2 * n_founders, n_snp), np.uint8)
    for h in range(2 * n_founders):                      # mosaics of 8 ancestral haplotypes in blocks of 100-200 SNPs: long-range LD
        s = 0
        while s < n_snp:
            L = int(rng.integers(100, 200))
            panel[h, s:s + L] = anc[rng.integers(0, 8), s:s + L]
            s += L
    cv_idx = np.sort(rng.choice(n_snp, size=40, replace=False))
    eng = capi.Engine(cuda_lib, n_pop=1, n_chr=1, n_phen=1, seed=414, capacity=N + 64, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS)
    eng.set_loci(0, pos)
    eng.set_population(0, False, True, 0.0)
    eng.set_genetic_map(0, 0, bp, p, dist)
    eng.set_founder_panel(0, 0, panel)
    eng.set_cv(0, 0, 0, pos[cv_idx], rng.normal(size=40), np.zeros(40), panel[:, cv_idx])
    eng.set_pheno_scheme(0, 0, va=0.5, vd=0.0, ve=0.5)
    eng.init_generation0()

    def ld():
        w = eng.haplotypes_packed(0, 0)
        x = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, :n_snp].astype(np.float32)
        x -= x.mean(axis=0, keepdims=True)
        return (x.T @ x / x.shape[0]).astype(np.float64)

    # expected crossovers between consecutive SNPs: row j contributes p_j * |[bp_j, bp_j + dist) ∩ (x, y]| / dist
    edges = np.concatenate([[0.0], np.cumsum(p)])         # cumulative rate at the START of each row's landing interval, spread uniformly over it
    def cum_rate(x):                                      # expected number of crossovers at positions <= x
        j = np.clip((x.astype(np.int64) - int(bp[0])) // dist, 0, len(bp) - 1)
        frac = np.clip((x.astype(np.float64) - bp[j].astype(np.float64)) / dist, 0.0, 1.0)
        return edges[j] + p[j] * frac
    lam = np.abs(cum_rate(pos)[:, None] - cum_rate(pos)[None, :])
    c = 0.5 * (1.0 - np.exp(-2.0 * lam))                  # Haldane: odd number of (near-Poisson) crossovers
    iu = np.triu_indices(n_snp, k=1)
    D0 = ld()[iu]
    gp = [capi.gen_params(N, 0.0, "p", "thr", 1, 1)]
    bins = [0.0, 0.005, 0.02, 0.05, 0.1]
    for gen in range(1, 21):
        eng.step_generation(gen, gp)
        if gen in (10, 20):
            Dt = ld()[iu]
            for lo, hi in zip(bins[:-1], bins[1:]):
                m = (c[iu] >= lo) & (c[iu] < hi) & (np.abs(D0) > 0.02)
                assert m.sum() > 500
                slope = np.sum(Dt[m] * D0[m]) / np.sum(D0[m] ** 2)
                expect = np.sum(D0[m] ** 2 * (1 - c[iu][m]) ** gen) / np.sum(D0[m] ** 2) * (1 - 1 / (2 * N)) ** gen
                assert abs(slope - expect) < 0.02, (gen, lo, hi, slope, expect)
