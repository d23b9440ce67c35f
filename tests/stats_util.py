"""Statistical-parity scenarios and statistics (north star, second correctness mode): under the GPU's own Philox
streams the allele-frequency trajectories, LD decay versus genetic distance, realised heritability and the
assortative-mating / selection responses must agree with replicate runs of the REAL reference.

This module defines the scenarios and computes the statistics from per-generation state; it is shared by
tests/golden/make_stats_reference.py (which runs the reference here and commits tests/golden/stats_*.{npz,json})
and by tests/test_gpu_statistics.py (which runs the CUDA library on the GPU box).  The recipes follow the
reference's own manual validation (GeneEvolveDocumentation.pdf ch. 3: MAF drift §3.2, heterozygosity decay
h(t) = (1-1/2N)^t h(0) §3.3, LD after t generations §3.4, var(A) under assortative mating §3.5).
"""
import numpy as np

N_REPLICATES = 16
SEEDS = [1000 + 17 * r for r in range(N_REPLICATES)]

# name -> scenario.  gen rows: (pop_size, mat_cor, offspring_dist, selection_func, par1, par2)
SCENARIOS = {
    # random mating, constant size, founder panel with strong block LD: drift + LD decay + h2
    "S_drift_ld": dict(rm=True, chrs=[1], n_founders=150, n_snp=240, n_cv=30, map_rows=81, map_step=1000, p_row=0.004,
                       ld_blocks=True, gens=[(300, 0.0, "p", "thr", 1, 1)] * 16, va=0.5, ve=0.5, ld_gens=[4, 8, 16]),
    # assortative mating rho = 0.4, Poisson families, no selection: var(A) inflation, spouse correlation
    "S_assort": dict(rm=False, chrs=[1, 2], n_founders=300, n_snp=60, n_cv=25, map_rows=21, map_step=1000, p_row=0.02,
                     ld_blocks=False, gens=[(600, 0.4, "p", "thr", 1, 1)] * 10, va=0.5, ve=0.5, ld_gens=[]),
    # directional selection logit(0, 2) on the phenotype: response of the mean additive value
    "S_select": dict(rm=False, chrs=[1, 2], n_founders=300, n_snp=60, n_cv=25, map_rows=21, map_step=1000, p_row=0.02,
                     ld_blocks=False, gens=[(600, 0.0, "p", "logit", 0, 2)] * 8, va=0.5, ve=0.5, ld_gens=[]),
}
C_BINS = [0.0, 0.01, 0.03, 0.08, 0.2]  # recombination-fraction bins of the LD statistic


def make_panel(sc, seed=20261018):
    """Founder panel, CV subset and effects of a scenario (deterministic).  Returns per chromosome
    dict(pos, hap[n_snp][nh], cv_idx, a)."""
    rng = np.random.default_rng(seed)
    nh = 2 * sc["n_founders"]
    out = []
    for c in sc["chrs"]:
        bp0 = 1000 * c
        span = (sc["map_rows"] - 1) * sc["map_step"]
        pos = np.sort(rng.choice(np.arange(bp0 + 1, bp0 + span - 1), size=sc["n_snp"], replace=False))
        if sc["ld_blocks"]:
            # founders are mosaics of 6 ancestral haplotypes in blocks of ~20 SNPs, 3 % noise: strong local LD
            anc = (rng.uniform(size=(6, sc["n_snp"])) < rng.uniform(0.2, 0.8, size=sc["n_snp"])[None, :]).astype(np.uint8)
            hap = np.zeros((sc["n_snp"], nh), np.uint8)
            for h in range(nh):
                s = 0
                while s < sc["n_snp"]:
                    L = int(rng.integers(10, 30))
                    hap[s:s + L, h] = anc[rng.integers(0, 6), s:s + L]
                    s += L
            hap ^= (rng.uniform(size=hap.shape) < 0.03).astype(np.uint8)
        else:
            freq = rng.uniform(0.15, 0.85, size=sc["n_snp"])
            hap = (rng.uniform(size=(sc["n_snp"], nh)) < freq[:, None]).astype(np.uint8)
        cv_idx = np.sort(rng.choice(sc["n_snp"], size=sc["n_cv"], replace=False))
        out.append(dict(chr=c, pos=pos, hap=hap, cv_idx=cv_idx, a=rng.normal(size=sc["n_cv"])))
    return out


def write_reference_inputs(sc, d, tag="s"):
    """The scenario in the reference's text formats; returns its CLI arguments (without --seed/--prefix)."""
    panel = make_panel(sc)
    with open(f"{d}/{tag}.rmap", "w") as f:
        f.write("chr bp cM\n")
        for c in sc["chrs"]:
            for j in range(sc["map_rows"]):
                f.write(f"{c} {1000 * c + j * sc['map_step']} {0.5 + 100.0 * sc['p_row'] * j:.12g}\n")
    with open(f"{d}/{tag}.indv", "w") as f:
        f.write("".join(f"id{i + 1}\n" for i in range(sc["n_founders"])))
    with open(f"{d}/{tag}.hapaddr", "w") as f:
        f.write("chr hap legend sample\n")
        for c in sc["chrs"]:
            f.write(f"{c} {d}/{tag}.chr{c}.hap {d}/{tag}.chr{c}.legend {d}/{tag}.indv\n")
    with open(f"{d}/{tag}.cvinfo", "w") as fi, open(f"{d}/{tag}.cvs", "w") as fc:
        fi.write("chr pos a d\n")
        for P in panel:
            c = P["chr"]
            with open(f"{d}/{tag}.chr{c}.legend", "w") as f:
                f.write("id pos allele0 allele1\n" + "".join(f"rs{c}_{k} {p} A C\n" for k, p in enumerate(P["pos"])))
            with open(f"{d}/{tag}.chr{c}.hap", "w") as f:
                for k in range(len(P["pos"])):
                    f.write(" ".join(map(str, P["hap"][k])) + " \n")
            with open(f"{d}/{tag}.cv.chr{c}.hap", "w") as f:
                for k in P["cv_idx"]:
                    f.write(" ".join(map(str, P["hap"][k])) + " \n")
            fi.write("".join(f"{c} {P['pos'][k]} {a:.10g} 0\n" for k, a in zip(P["cv_idx"], P["a"])))
            fc.write(f"{c} {d}/{tag}.cv.chr{c}.hap\n")
    with open(f"{d}/{tag}.gen", "w") as f:
        f.write("pop_size mat_cor offspring_dist selection_func selection_func_par1 selection_func_par2\n")
        for r in sc["gens"]:
            f.write(" ".join(str(x) for x in r) + "\n")
    args = ["--file_gen_info", f"{d}/{tag}.gen", "--file_hap_name", f"{d}/{tag}.hapaddr", "--file_recom_map", f"{d}/{tag}.rmap",
            "--file_cv_info", f"{d}/{tag}.cvinfo", "--file_cvs", f"{d}/{tag}.cvs", "--va", str(sc["va"]), "--vd", "0", "--ve", str(sc["ve"])]
    if sc["rm"]:
        args.append("--RM")
    return args


def recomb_fraction(sc, pos):
    """Recombination fraction between SNP pairs of one chromosome under the scenario's uniform map: a crossover in
    map row j lands in [bp_j, bp_j + step), so the expected number between two positions is p_row per step; the
    fraction is Haldane's (1 - exp(-2d))/2."""
    d = np.abs(pos[:, None].astype(float) - pos[None, :].astype(float)) / sc["map_step"] * sc["p_row"]
    return 0.5 * (1.0 - np.exp(-2.0 * d))


class Trajectory:
    """Accumulates the per-generation statistics of ONE run from its state, generation by generation."""

    def __init__(self, sc):
        self.sc = sc
        self.pos0 = make_panel(sc)[0]["pos"]
        self.p0 = None        # allele frequencies at generation 0 per chromosome
        self.D0 = None        # LD matrix (covariance of alleles) at generation 0, chromosome 0
        self.out = {k: [] for k in ("het", "dp2", "h2", "varA", "varP", "meanA", "n")}
        self.ld = {}          # gen -> slope of D_t on D_0 per recombination-fraction bin
        self.spouse_cor, self.n_couples = [], []

    @staticmethod
    def _ld(h):  # h: [n_hap][n_snp] 0/1 -> covariance matrix
        x = h.astype(float) - h.mean(axis=0, keepdims=True)
        return x.T @ x / h.shape[0]

    def add_generation(self, gen, haps, A, P):
        """haps: list over chromosomes of [2n][n_snp] uint8 (Hap_SNP layout); A, P: [n] of phenotype 0."""
        p = [h.mean(axis=0) for h in haps]
        if gen == 0:
            self.p0 = p
            self.D0 = self._ld(haps[0])
        het = np.mean(np.concatenate([2 * q * (1 - q) for q in p]))
        het0 = np.mean(np.concatenate([2 * q * (1 - q) for q in self.p0]))
        w = np.concatenate([q0 * (1 - q0) for q0 in self.p0])
        dp = np.concatenate([(q - q0) ** 2 for q, q0 in zip(p, self.p0)])
        self.out["het"].append(het / het0)
        self.out["dp2"].append(float(np.sum(dp) / np.sum(w)))
        self.out["h2"].append(float(np.var(A, ddof=1) / np.var(P, ddof=1)))
        self.out["varA"].append(float(np.var(A, ddof=1)))
        self.out["varP"].append(float(np.var(P, ddof=1)))
        self.out["meanA"].append(float(np.mean(A)))
        self.out["n"].append(int(len(A)))
        if gen in self.sc["ld_gens"]:
            Dt = self._ld(haps[0])
            c = recomb_fraction(self.sc, self.pos0)
            iu = np.triu_indices_from(c, k=1)
            slopes = []
            for lo, hi in zip(C_BINS[:-1], C_BINS[1:]):
                m = (c[iu] >= lo) & (c[iu] < hi)
                slopes.append(float(np.sum(Dt[iu][m] * self.D0[iu][m]) / np.sum(self.D0[iu][m] ** 2)))
            self.ld[gen] = slopes

    def add_couples(self, mv_parents, pos_male, pos_female, num_offspring):
        k = np.asarray(num_offspring) > 0
        a, b = np.asarray(mv_parents)[np.asarray(pos_male)[k]], np.asarray(mv_parents)[np.asarray(pos_female)[k]]
        self.spouse_cor.append(float(np.corrcoef(a, b)[0, 1]) if len(a) > 2 else 0.0)
        self.n_couples.append(int(len(pos_male)))

    def vector(self):
        """name -> 1-D array; the statistics compared between the GPU and the reference."""
        v = {k: np.asarray(x, float) for k, x in self.out.items()}
        for g, s in self.ld.items():
            v[f"ld_slope_gen{g}"] = np.asarray(s, float)
        if self.spouse_cor:
            v["spouse_cor"] = np.asarray(self.spouse_cor, float)
            v["n_couples"] = np.asarray(self.n_couples, float)
        return v


def compare(ref, gpu, what, n_sigma=4.5, rel=0.02, abs_tol=1e-3):
    """ref, gpu: [replicates][len] arrays of one statistic.  The replicate means must agree within
    n_sigma standard errors of their difference plus a small relative/absolute slack (the slack covers statistics
    whose replicate variance is ~0, e.g. generation-0 values fixed by construction)."""
    ref, gpu = np.asarray(ref, float), np.asarray(gpu, float)
    m_r, m_g = ref.mean(axis=0), gpu.mean(axis=0)
    se = np.sqrt(ref.var(axis=0, ddof=1) / ref.shape[0] + gpu.var(axis=0, ddof=1) / gpu.shape[0])
    tol = n_sigma * se + rel * np.abs(m_r) + abs_tol
    bad = np.abs(m_r - m_g) > tol
    assert not bad.any(), f"{what}: GPU mean {m_g[bad]} vs reference mean {m_r[bad]} (tolerance {tol[bad]}) at {np.nonzero(bad)[0]}"
