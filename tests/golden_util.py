"""Helpers shared by the parity tests: load a golden fixture (tests/golden/*.npz, produced by the REAL
reference through oracle/ref_driver.cpp) and configure an engine (oracle or CUDA) from it."""
import glob
import os

import numpy as np

from geneevolve_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
SCENARIOS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                   if not os.path.basename(p).startswith("stats_inputs_"))  # those hold inputs only (tests/test_statistics.py)


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = dict(np.load(os.path.join(HERE, "golden", name + ".npz")))
        self.n_pop = int(self.z["in.n_pop"])
        self.G = int(self.z["in.tot_gen"])
        self.n_chr = int(self.z["in.p0.nchr"])
        self.n_phen = int(self.z["in.p0.nphen"])
        self.vt_type = int(self.z["in.vt_type"])
        self.seed = int(self.z["in.seed"])

    def __getitem__(self, k):
        return self.z[k]

    def g(self, gen, pop, name):
        return self.z[f"g{gen}.p{pop}.{name}"]

    def params(self, gen, pop):
        """Row gen-1 of the generation table of `pop`."""
        z, pre, k = self.z, f"in.p{pop}.", gen - 1
        return capi.ge_gen_params(int(z[pre + "pop_size"][k]), float(z[pre + "mat_cor"][k]), int(z[pre + "offspring_dist"][k]),
                                  int(z[pre + "selection_func"][k]), float(z[pre + "selection_par1"][k]), float(z[pre + "selection_par2"][k]))

    def all_params(self, gen):
        return [self.params(gen, p) for p in range(self.n_pop)]

    def migration_row(self, gen):
        return self.z["in.migration"][gen - 1] if "in.migration" in self.z else None

    def philox_capacity(self):
        """Room for Poisson family sizes and migration under the library's own draws."""
        return int(1.5 * max(int(self.z[f"g{g}.p{p}.n"]) for g in range(self.G + 1) for p in range(self.n_pop))) + 300

    def engine_kwargs(self, **over):
        cap = max(int(self.z[f"g{g}.p{p}.n"]) for g in range(self.G + 1) for p in range(self.n_pop)) + 64
        kw = dict(n_pop=self.n_pop, n_chr=self.n_chr, n_phen=self.n_phen, vt_type=self.vt_type, seed=self.seed, capacity=cap)
        kw.update(over)
        return kw

    def configure(self, eng):
        z = self.z
        for c in range(self.n_chr):
            eng.set_loci(c, z[f"in.p0.c{c}.panel_pos"])
        for p in range(self.n_pop):
            pre = f"in.p{p}."
            eng.set_population(p, bool(z[pre + "avoid_inbreeding"]), bool(z[pre + "RM"]), float(z[pre + "MM_percent"]))
            for c in range(self.n_chr):
                cp = pre + f"c{c}."
                eng.set_genetic_map(p, c, z[cp + "rmap_bp"], z[cp + "recom_prob"], int(z[cp + "bp_dist"]))
                if int(z[pre + "has_mutation_map"]):
                    eng.set_mutation_map(p, c, z[cp + "mut_bp"], z[cp + "mut_rate"])
                eng.set_founder_panel(p, c, z[cp + "panel"])
                for f in range(self.n_phen):
                    fp = cp + f"f{f}."
                    eng.set_cv(p, f, c, z[fp + "cv_bp"], z[fp + "cv_a"], z[fp + "cv_d"], z[fp + "cv_val"])
            for f in range(self.n_phen):
                s = z[pre + "scheme"][f]
                eng.set_pheno_scheme(p, f, va=s[0], vd=s[1], ve=s[2], vc=s[3], vf=s[4], omega=s[5], beta=s[6], lam=s[7])
        if len(z["in.gamma"]):
            eng.set_gamma(z["in.gamma"])

    def draws0(self, pop):
        """Generation-0 draws (sex, N(0,1) environment draws, sibling-common and parental effects)."""
        n = int(self.g(0, pop, "n"))
        return capi.Draws(n, sex=self.g(0, pop, "sex"), e_raw=self.g(0, pop, "e_raw"), common=self.g(0, pop, "C"),
                          parental0=self.g(0, pop, "F"))

    def draws(self, gen, pop):
        g = lambda k: self.g(gen, pop, k)  # noqa: E731
        n = len(g("off_sex"))
        kw = dict(father=g("off_father"), mother=g("off_mother"), sex=g("off_sex"), xo_off=g("xo_off"), xo_bp=g("xo_bp"),
                  start_hap=g("start_hap"), e_raw=g("e_raw"), common=self.birth_order(gen, pop, "C"))
        if int(self.z[f"in.p{pop}.has_mutation_map"]):
            kw.update(mut_off=g("mut_off"), mut_bp=g("mut_bp"), mut_gam=g("mut_gam"))
        return capi.Draws(n, **kw)

    def mate_draws(self, gen, pop):
        """Keyword arguments of Engine.mate_replay: the draws the reference's random_mate / assort_mate consumed in this generation."""
        m = lambda k: self.g(gen, pop, "mate." + k)  # noqa: E731
        if int(self.z[f"in.p{pop}.RM"]):
            return dict(thin_u=m("thin_u"), rm_father_idx=m("rm_father_idx"), rm_mother_idx=m("rm_mother_idx"))
        kw = dict(thin_u=m("thin_u"), mm_u=np.nan_to_num(m("mm_u"), nan=2.0), t1=m("t1"), t2=m("t2"))
        if len(m("trim_order")):
            kw["trim_order"] = m("trim_order")
        if len(m("family")):
            kw["family"] = m("family")
        if len(m("remainder_order")):
            kw["remainder_order"] = m("remainder_order")
        return kw

    def birth_order(self, gen, pop, name):
        """Per-offspring [n_phen][n] array `name` in birth order.  Without migration that is the exported
        state; with migration the children are found again by (ID, father ID, mother ID)."""
        if self.n_pop == 1:
            return self.g(gen, pop, name)
        n = len(self.g(gen, pop, "off_sex"))
        out = np.zeros((self.n_phen, n))
        found = np.zeros(n, bool)
        fa, mo = self.g(gen, pop, "off_father"), self.g(gen, pop, "off_mother")
        par_ids = self.g(gen - 1, pop, "ids")
        for q in range(self.n_pop):
            ids, val = self.g(gen, q, "ids"), self.g(gen, q, name)
            for j in range(ids.shape[0]):
                i = int(ids[j, 0])
                if i < n and not found[i] and ids[j, 1] == par_ids[fa[i], 0] and ids[j, 2] == par_ids[mo[i], 0]:
                    out[:, i] = val[:, j]
                    found[i] = True
        assert found.all()
        return out

    def migration_sample(self, gen, pop):
        """Positions (in the pre-migration population `pop`) of the individuals the reference moved out in
        generation `gen`: the stayers head the post-migration population in their original order."""
        n_pre = len(self.g(gen, pop, "premig_ids"))
        row = self.migration_row(gen).reshape(self.n_pop, self.n_pop)
        s = sum(int(round(row[pop, j] * n_pre)) for j in range(self.n_pop) if j != pop)
        stay = self.g(gen, pop, "ids")[:n_pre - s, 0]
        return np.array(sorted(set(range(n_pre)) - set(int(x) for x in stay), reverse=True), dtype=np.uint64)

    def step_replay(self, eng, gen):
        """One generation with the reference's draws through the composite entry point."""
        if self.n_pop > 1:
            for p in range(self.n_pop):
                eng.set_migration_sample(p, self.migration_sample(gen, p))
        eng.step_generation(gen, self.all_params(gen), self.migration_row(gen), [self.draws(gen, p) for p in range(self.n_pop)])
