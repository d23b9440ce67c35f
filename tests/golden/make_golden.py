#!/usr/bin/env python
"""Regenerates the golden fixtures under tests/golden/ by running the REAL reference.

Needs oracle/_ref/ge_ref_export (built by `make -C oracle ref`, which needs /root/reference), so it only
runs in the build container; the resulting *.npz files are committed and travel to the GPU box.

Each scenario writes tiny synthetic inputs in the reference's own text formats (SURVEY.md §5 "config"
row), runs the reference through oracle/ref_driver.cpp (which self-checks its replay against the untouched
Simulation::reproduce), and stores every exported array (inputs as the reference parsed them, couples,
crossovers, start haplotypes, mutation hits, sex, N(0,1) draws, per-generation individual state, segment
lists and the materialised haplotype matrix) in one compressed npz.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.gex import read_gex  # noqa: E402

EXPORT = os.path.join(ROOT, "oracle", "_ref", "ge_ref_export")


def write_inputs(d, rng, *, tag, chrs, n_founders, n_snp, n_cv, n_phen, map_rows, map_step, p_row,
                 mut_step=None, mut_rate=0.0, cv_scale=1.0):
    """Writes one population's input files; returns the CLI fragment naming them."""
    nh = 2 * n_founders
    # genetic map: uniform bp step (like Recom.Map.b37.*KbDiff), cM increments random around p_row
    with open(f"{d}/{tag}.rmap", "w") as f:
        f.write("chr bp cM\n")
        for c in chrs:
            bp0 = 1000 * c
            cm = 0.5
            for j in range(map_rows):
                if j > 0:
                    cm += 100.0 * p_row * rng.uniform(0.2, 1.8)
                f.write(f"{c} {bp0 + j * map_step} {cm:.12g}\n")
    if mut_step:
        with open(f"{d}/{tag}.mutmap", "w") as f:
            f.write("chr bp mutation_rate\n")
            for c in chrs:
                bp0 = 1000 * c
                nrow = (map_rows - 1) * map_step // mut_step + 1
                for j in range(nrow):
                    f.write(f"{c} {bp0 + j * mut_step} {mut_rate}\n")
    with open(f"{d}/{tag}.hapaddr", "w") as f:
        f.write("chr hap legend sample\n")
        for c in chrs:
            f.write(f"{c} {d}/{tag}.chr{c}.hap {d}/{tag}.chr{c}.legend {d}/{tag}.indv\n")
    with open(f"{d}/{tag}.indv", "w") as f:
        for i in range(n_founders):
            f.write(f"{tag}_id{i + 1}\n")
    span = (map_rows - 1) * map_step
    for c in chrs:
        bp0 = 1000 * c
        # SNP positions: a few before the first map row and a few at/after the last one (never covered by
        # any segment, :3029-3034), the rest inside
        lo, hi = bp0 - 3 * max(1, map_step // 4), bp0 + span + 3 * max(1, map_step // 4)
        pos = np.sort(rng.choice(np.arange(lo, hi), size=n_snp, replace=False))
        freq = rng.uniform(0.1, 0.9, size=n_snp)
        hap = (rng.uniform(size=(n_snp, nh)) < freq[:, None]).astype(int)
        with open(f"{d}/{tag}.chr{c}.legend", "w") as f:
            f.write("id pos allele0 allele1\n")
            for k, p in enumerate(pos):
                f.write(f"rs{c}_{k} {p} A C\n")
        with open(f"{d}/{tag}.chr{c}.hap", "w") as f:
            for k in range(n_snp):
                f.write(" ".join(map(str, hap[k])) + " \n")
        # CVs are a subset of the SNPs (their founder alleles must agree with the panel)
        for ph in range(n_phen):
            sel = np.sort(rng.choice(n_snp, size=n_cv, replace=False))
            with open(f"{d}/{tag}.ph{ph}.cv.chr{c}.hap", "w") as f:
                for k in sel:
                    f.write(" ".join(map(str, hap[k])) + " \n")
            mode = "a" if c != chrs[0] else "w"
            with open(f"{d}/{tag}.ph{ph}.cvinfo", mode) as f:
                if c == chrs[0]:
                    f.write("chr pos a d\n")
                for k in sel:
                    f.write(f"{c} {pos[k]} {rng.normal() * cv_scale:.10g} {rng.normal() * 0.5 * cv_scale:.10g}\n")
            with open(f"{d}/{tag}.ph{ph}.cvs", mode) as f:
                f.write(f"{c} {d}/{tag}.ph{ph}.cv.chr{c}.hap\n")
    args = ["--file_hap_name", f"{d}/{tag}.hapaddr", "--file_recom_map", f"{d}/{tag}.rmap"]
    if mut_step:
        args += ["--file_mutation_map", f"{d}/{tag}.mutmap"]
    for ph in range(n_phen):
        args += ["--file_cv_info", f"{d}/{tag}.ph{ph}.cvinfo", "--file_cvs", f"{d}/{tag}.ph{ph}.cvs"]
    return args


def write_geninfo(path, rows):
    with open(path, "w") as f:
        f.write("pop_size mat_cor offspring_dist selection_func selection_func_par1 selection_func_par2\n")
        for r in rows:
            f.write(" ".join(str(x) for x in r) + "\n")


def run(name, args, seed):
    with tempfile.TemporaryDirectory() as t:
        gex = os.path.join(t, "out.gex")
        cmd = [EXPORT, "--export", gex, "--export_hap"] + args + ["--seed", str(seed), "--prefix", os.path.join(t, "o")]
        subprocess.run(cmd, check=True)
        arrs = read_gex(gex)
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, **arrs)
    print(name, len(arrs), "arrays", os.path.getsize(out) // 1024, "KiB")


def main():
    with tempfile.TemporaryDirectory() as d:
        # A. assortative mating + Poisson families + logit selection, dominance, two phenotypes
        rng = np.random.default_rng(101)
        a = write_inputs(d, rng, tag="A", chrs=[1, 2], n_founders=40, n_snp=150, n_cv=9, n_phen=2,
                         map_rows=30, map_step=50, p_row=0.04)
        write_geninfo(f"{d}/A.gen", [(60, 0.4, "p", "logit", 0, 1), (70, 0.4, "p", "logit", 0.5, 1.5),
                                      (64, -0.3, "p", "logit", 0, 1), (60, 0.6, "p", "logit", 0, 1)])
        run("A_am_pois", ["--file_gen_info", f"{d}/A.gen"] + a +
            ["--va", "0.5", "--vd", "0.2", "--ve", "0.3", "--omega", "1", "--lambda", "1",
             "--va", "0.6", "--vd", "0", "--ve", "0.4", "--omega", "0.5", "--lambda", "0"], 12345)

        # B. random mating + mutation map (dense so that mutations hit SNP/CV sites), 3 chromosomes, odd sizes
        rng = np.random.default_rng(202)
        b = write_inputs(d, rng, tag="B", chrs=[3, 5, 7], n_founders=33, n_snp=131, n_cv=7, n_phen=1,
                         map_rows=41, map_step=8, p_row=0.05, mut_step=4, mut_rate=0.06)
        write_geninfo(f"{d}/B.gen", [(50, 0, "p", "thr", 1, 1)] * 3 + [(45, 0, "p", "probit", 0, 1)] * 2)
        run("B_rm_mut", ["--file_gen_info", f"{d}/B.gen"] + b + ["--RM", "--va", "1", "--vd", "0", "--ve", "1"], 777)

        # C. fixed family size, --MM, sibling-common and vertical-transmission effects (vt_type 2), every
        #    selection function.  (No --avoid_inbreeding here: with 'f' the reference indexes an empty
        #    pos_couple_can_marry, src/Simulation.cpp:2314-2353, and segfaults.)
        rng = np.random.default_rng(303)
        c = write_inputs(d, rng, tag="C", chrs=[1, 2], n_founders=50, n_snp=100, n_cv=10, n_phen=1,
                         map_rows=25, map_step=40, p_row=0.03)
        write_geninfo(f"{d}/C.gen", [(80, 0.3, "f", "logit", 0, 1), (80, 0.3, "f", "probit", 0.1, 1.2),
                                      (90, 0.5, "f", "stab", 0, 1), (85, 0.2, "f", "thr", 0.5, -0.3),
                                      (80, 0.0, "p", "logit", 0, 2)])
        run("C_fixed_mm", ["--file_gen_info", f"{d}/C.gen"] + c +
            ["--va", "0.4", "--vd", "0.1", "--ve", "0.3", "--vc", "0.1", "--vf", "0.1", "--vt_type", "2",
             "--MM", "0.3"], 4242)

        # E. inbreeding avoidance with Poisson families over enough generations for cousins to meet,
        #    sibling-common effect, vt_type 1
        rng = np.random.default_rng(505)
        e = write_inputs(d, rng, tag="E", chrs=[4], n_founders=24, n_snp=70, n_cv=8, n_phen=1,
                         map_rows=16, map_step=64, p_row=0.06)
        write_geninfo(f"{d}/E.gen", [(40, 0.5, "p", "logit", 0, 1)] * 6)
        run("E_avoid_inbreeding", ["--file_gen_info", f"{d}/E.gen"] + e +
            ["--va", "0.5", "--vd", "0", "--ve", "0.4", "--vc", "0.1", "--vf", "0.05", "--avoid_inbreeding"], 31337)

        # D. two populations with migration (≤ 1 non-zero off-diagonal entry per row),
        #    population-specific environmental shift (--gamma).  vf = 0: with migration the reference reads
        #    _Pop_info_prev_gen out of bounds for the parental effect (parent ID vs. position, :3118-3133)
        rng = np.random.default_rng(404)
        d1 = write_inputs(d, rng, tag="D1", chrs=[1, 2], n_founders=30, n_snp=90, n_cv=6, n_phen=1,
                          map_rows=20, map_step=50, p_row=0.05)
        # the second population shares SNP/CV positions with the first (required by :1186-1230) but has
        # its own founders and effect sizes: reuse the rng-independent layout by regenerating with same seed
        rng2 = np.random.default_rng(404)
        d2 = write_inputs(d, rng2, tag="D2", chrs=[1, 2], n_founders=30, n_snp=90, n_cv=6, n_phen=1,
                          map_rows=20, map_step=50, p_row=0.05, cv_scale=1.7)
        # different founder alleles for population 2: rewrite hap + cv hap files with flipped random bits
        rng3 = np.random.default_rng(405)
        for cc in [1, 2]:
            hap = np.loadtxt(f"{d}/D2.chr{cc}.hap", dtype=int)
            flip = rng3.uniform(size=hap.shape) < 0.35
            hap2 = hap ^ flip.astype(int)
            pos = [int(l.split()[1]) for l in open(f"{d}/D2.chr{cc}.legend").read().splitlines()[1:]]
            with open(f"{d}/D2.chr{cc}.hap", "w") as f:
                for k in range(hap2.shape[0]):
                    f.write(" ".join(map(str, hap2[k])) + " \n")
            cvpos = [int(l.split()[1]) for l in open(f"{d}/D2.ph0.cvinfo").read().splitlines()[1:] if int(l.split()[0]) == cc]
            with open(f"{d}/D2.ph0.cv.chr{cc}.hap", "w") as f:
                for p in cvpos:
                    f.write(" ".join(map(str, hap2[pos.index(p)])) + " \n")
        write_geninfo(f"{d}/D1.gen", [(50, 0.2, "p", "logit", 0, 1)] * 4)
        write_geninfo(f"{d}/D2.gen", [(40, 0.0, "p", "logit", 0, 1)] * 4)
        with open(f"{d}/D.mig", "w") as f:
            for _ in range(4):
                f.write("0.9 0.1 0.2 0.8\n")
        run("D_two_pops", ["--file_gen_info", f"{d}/D1.gen"] + d1 + ["--va", "0.5", "--vd", "0", "--ve", "0.5"] +
            ["--next_population", "--file_gen_info", f"{d}/D2.gen"] + d2 + ["--va", "0.5", "--vd", "0", "--ve", "0.5"] +
            ["--file_migration", f"{d}/D.mig", "--gamma", "0.2"], 99)


def three_populations():
    """F. BASELINE config 4 in miniature: three populations of different sizes, ring migration (the reference only
    survives matrices with at most one non-zero off-diagonal entry per row), two phenotypes with
    population-specific effect sizes, assortative mating in one population, random mating in another."""
    with tempfile.TemporaryDirectory() as d:
        frags = []
        for k, (tag, scale) in enumerate([("F1", 1.0), ("F2", 1.4), ("F3", 0.7)]):
            rng = np.random.default_rng(606)   # same SNP / CV positions in every population (required by :1186-1230, :2762)
            frag = write_inputs(d, rng, tag=tag, chrs=[1, 2], n_founders=26, n_snp=80, n_cv=5, n_phen=2,
                                map_rows=18, map_step=50, p_row=0.05, cv_scale=scale)
            rngk = np.random.default_rng(607 + k)   # population-specific founder alleles
            for cc in [1, 2]:
                hap = np.loadtxt(f"{d}/{tag}.chr{cc}.hap", dtype=int)
                hap2 = hap ^ (rngk.uniform(size=hap.shape) < 0.3).astype(int)
                pos = [int(l.split()[1]) for l in open(f"{d}/{tag}.chr{cc}.legend").read().splitlines()[1:]]
                with open(f"{d}/{tag}.chr{cc}.hap", "w") as f:
                    for r in range(hap2.shape[0]):
                        f.write(" ".join(map(str, hap2[r])) + " \n")
                for ph in range(2):
                    cvpos = [int(l.split()[1]) for l in open(f"{d}/{tag}.ph{ph}.cvinfo").read().splitlines()[1:] if int(l.split()[0]) == cc]
                    with open(f"{d}/{tag}.ph{ph}.cv.chr{cc}.hap", "w") as f:
                        for p in cvpos:
                            f.write(" ".join(map(str, hap2[pos.index(p)])) + " \n")
            frags.append(frag)
        write_geninfo(f"{d}/F1.gen", [(50, 0.3, "p", "logit", 0, 1)] * 3)
        write_geninfo(f"{d}/F2.gen", [(40, 0.0, "p", "probit", 0, 1)] * 3)
        write_geninfo(f"{d}/F3.gen", [(30, 0.0, "f", "logit", 0.2, 1.5)] * 3)
        with open(f"{d}/F.mig", "w") as f:
            for _ in range(3):
                f.write("0.9 0.1 0 0 0.85 0.15 0.2 0 0.8\n")
        ph = ["--va", "0.5", "--vd", "0", "--ve", "0.5", "--omega", "1", "--lambda", "1", "--va", "0.4", "--vd", "0.1", "--ve", "0.5", "--omega", "0.5", "--lambda", "0"]
        run("F_three_pops_ring", ["--file_gen_info", f"{d}/F1.gen"] + frags[0] + ph +
            ["--next_population", "--file_gen_info", f"{d}/F2.gen"] + frags[1] + ph + ["--RM"] +
            ["--next_population", "--file_gen_info", f"{d}/F3.gen"] + frags[2] + ph +
            ["--file_migration", f"{d}/F.mig"], 2026)


def bundled_example():
    """G. BASELINE config 1: the reference's own bundled example (Examples.zip: 2 000 founders, b37 50 kb map, 100 CVs per
    chromosome) filtered to chromosome 1, random mating, 1 000 individuals per generation, three generations,
    `--seed 12345` — the reference's real inputs and real genetic map (4 971 rows on chr1) rather than synthetic ones."""
    import zipfile
    with tempfile.TemporaryDirectory() as d:
        zipfile.ZipFile("/root/reference/Examples.zip").extractall(d)
        e = os.path.join(d, "Examples")
        with open(f"{e}/g.hapaddr", "w") as f:
            f.write(f"chr hap legend sample\n1 {e}/ref.chr1.hap {e}/ref.chr1.legend {e}/ref.chr1.indv\n")
        with open(f"{e}/g.cvinfo", "w") as f:   # the loader rejects CV rows of inactive chromosomes (src/Population.cpp:250-254)
            lines = open(f"{e}/cv.info").read().splitlines()
            f.write(lines[0] + "\n" + "".join(l + "\n" for l in lines[1:] if l.split()[0] == "1"))
        with open(f"{e}/g.cvs", "w") as f:
            f.write(f"1 {e}/cv.chr1.hap\n")
        write_geninfo(f"{e}/g.gen", [(1000, 0, "p", "thr", 1, 1)] * 3)
        run("G_bundled_example_chr1", ["--file_gen_info", f"{e}/g.gen", "--file_hap_name", f"{e}/g.hapaddr", "--file_recom_map", f"{e}/Recom.Map.b37.50KbDiff",
                                       "--file_cv_info", f"{e}/g.cvinfo", "--file_cvs", f"{e}/g.cvs", "--RM"], 12345)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "F":
        three_populations()     # added later; the other fixtures are not regenerated
    elif which == "G":
        bundled_example()
    else:
        main()
        three_populations()
        bundled_example()
