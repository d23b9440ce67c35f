"""Synthetic inputs of the named BASELINE.json configurations.

Everything here is host-side input generation (numpy): genetic maps with the shape of the reference's
Recom.Map.b37.50KbDiff (22 autosomes, uniform 50 kb rows, ~35.9 Morgans), locus positions apportioned by
chromosome length, causal variants and founder alleles.  The same arrays can be written out in the
reference's own text formats (write_reference_inputs) so the reference binary runs on identical inputs.
"""
import os

import numpy as np

# GRCh37 autosome lengths (bp) and sex-averaged genetic lengths (cM) — the shape of the bundled b37 map
CHR_BP = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431,
          135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248,
          59128983, 63025520, 48129895, 51304566]
CHR_CM = [286.3, 268.8, 223.4, 214.7, 204.1, 192.0, 187.2, 168.0, 166.4, 181.1, 158.4, 174.7, 125.7, 120.2, 141.9,
          134.0, 128.5, 117.7, 107.7, 108.3, 62.8, 74.1]

CONFIGS = {
    # name: individuals per generation, loci, chromosomes, founders, causal variants, map step
    "config1_bundled_1chr": dict(n=1000, loci=1000, chrs=[1], founders=2000, n_cv=100, rm=True, mat_cor=0.0),
    "config2_chr22_10k": dict(n=10000, loci=500000, chrs=[22], founders=2000, n_cv=1000, rm=False, mat_cor=0.0),
    "config3_100k_x_1M": dict(n=100000, loci=1000000, chrs=list(range(1, 23)), founders=4000, n_cv=1000, rm=False, mat_cor=0.4),
    # config 4: three populations of different sizes, ring migration (a full matrix crashes the reference),
    # two phenotypes; 150 GB of bit-packed rows per generation -> needs >= 4 GPUs (chromosome shards)
    "config4_3pop_300k_x_2M": dict(n=150000, pops=[150000, 100000, 50000], loci=2000000, chrs=list(range(1, 23)), founders=4000, n_cv=1000, n_phen=2,
                                   rm=False, mat_cor=0.4, migration=[0.98, 0.02, 0.0, 0.0, 0.98, 0.02, 0.02, 0.0, 0.98]),
    # config 5 (whole-genome sequence): 2.5 TB per generation bit-packed does not fit, so it runs on founder segments
    # like the reference; the 10M loci are nominal (the segment path never touches non-causal loci)
    "config5_1M_x_10M_segments": dict(n=1000000, loci=10000000, chrs=list(range(1, 23)), founders=4000, n_cv=1000, rm=False, mat_cor=0.4,
                                      segments=True),
}


def genetic_map(chrs, step=50000, seed=20261018):
    """Per chromosome: (bp[R], cM[R], recom_prob[R]) like Population::ras_read_rmap + ras_compute_recom_prob
    (src/Population.cpp:349-414, 471-507): p[0] = 0, p[k] = (cM[k]-cM[k-1]) * 0.01."""
    rng = np.random.default_rng(seed)
    out = []
    for c in chrs:
        L, cm_tot = CHR_BP[c - 1], CHR_CM[c - 1]
        R = L // step
        bp = 700000 + step * np.arange(R, dtype=np.uint64)
        w = rng.gamma(0.5, 1.0, size=R - 1)  # hot and cold spots like a real map
        cm = np.concatenate([[0.0], np.cumsum(w / w.sum() * cm_tot)]) + 0.5
        p = np.zeros(R)
        p[1:] = (cm[1:] - cm[:-1]) * 0.01
        out.append((bp, cm, p))
    return out


def make_workload(name, seed=20261018, n_override=None, loci_override=None, founders_override=None):
    cfg = dict(CONFIGS[name])
    if founders_override:
        cfg["founders"] = founders_override
    if n_override:
        cfg["n"] = n_override
    if loci_override:
        cfg["loci"] = loci_override
    rng = np.random.default_rng(seed)
    chrs = cfg["chrs"]
    maps = genetic_map(chrs, seed=seed)
    spans = np.array([float(m[0][-1] - m[0][0]) for m in maps])
    n_loci = np.maximum(1, np.floor(cfg["loci"] * spans / spans.sum()).astype(int))
    n_loci[0] += cfg["loci"] - n_loci.sum()
    loci = []
    for (bp, _, _), nl in zip(maps, n_loci):
        lo, hi = int(bp[0]), int(bp[-1])
        if cfg.get("segments"):  # only the causal variants need positions
            nl = max(1000, cfg["n_cv"])
        pos = np.sort(rng.choice(hi - lo, size=nl, replace=False).astype(np.uint64) + np.uint64(lo)) if nl < (hi - lo) // 4 else \
            np.unique(rng.integers(lo, hi, size=nl * 2, dtype=np.uint64))[:nl]
        loci.append(pos)
    # causal variants: drawn from the loci, a ~ N(0,1), d = 0
    n_cv = np.maximum(1, np.floor(cfg["n_cv"] * n_loci / n_loci.sum()).astype(int))
    n_cv[0] += cfg["n_cv"] - n_cv.sum()
    cvs = []
    nh = 2 * cfg["founders"]
    for pos, k in zip(loci, n_cv):
        sel = np.sort(rng.choice(len(pos), size=k, replace=False))
        f = np.clip(rng.beta(0.5, 0.5, size=k), 0.01, 0.99)
        val = (rng.random((nh, k)) < f[None, :]).astype(np.uint8)
        cvs.append(dict(bp=pos[sel], a=rng.normal(size=k), d=np.zeros(k), val=val, idx=sel))
    cfg.update(maps=maps, loci=loci, cvs=cvs, n_loci=[int(x) for x in n_loci] if cfg.get("segments") else [len(p) for p in loci], seed=seed)
    if "pops" in cfg:
        # further phenotypes: their own CV sets (positions shared by all populations, as ras_find_cv requires); every
        # population has its own founders, hence its own CV alleles (index [phenotype][chromosome]['val'][population])
        all_cvs = []
        for f in range(cfg["n_phen"]):
            per_chr = []
            for c, (pos, k) in enumerate(zip(loci, n_cv)):
                r = np.random.default_rng([seed, 1000 + f, c])
                sel = np.sort(r.choice(len(pos), size=k, replace=False))
                fr = np.clip(r.beta(0.5, 0.5, size=k), 0.01, 0.99)
                val = [(np.random.default_rng([seed, 2000 + f, c, q]).random((nh, k)) < fr[None, :]).astype(np.uint8) for q in range(len(cfg["pops"]))]
                per_chr.append(dict(bp=pos[sel], a=r.normal(size=k), d=np.zeros(k), val=val, idx=sel))
            all_cvs.append(per_chr)
        cfg["cvs_multi"] = all_cvs
    return cfg


def founder_words(cfg, c, rng):
    """Bit-packed founder panel of chromosome index c: random alleles, causal-variant columns forced to
    agree with the CV panel (like cv.chr*.hap being rows of ref.chr*.hap in the bundled examples)."""
    nh, nl = 2 * cfg["founders"], cfg["n_loci"][c]
    nw = (nl + 31) // 32
    w = rng.integers(0, 2 ** 32, size=(nh, nw), dtype=np.uint32)
    if nl % 32:
        w[:, -1] &= np.uint32((1 << (nl % 32)) - 1)
    cv = cfg["cvs"][c]
    for k, s in enumerate(cv["idx"]):
        word, bit = int(s) // 32, np.uint32(1 << (int(s) % 32))
        w[:, word] = (w[:, word] & ~bit) | (cv["val"][:, k].astype(np.uint32) * bit)
    return w


def founder_words_multi(cfg, c, pop, rng):
    """Like founder_words for population `pop` of a multi-population workload: the CV columns of every phenotype agree with
    that population's CV panels."""
    nh, nl = 2 * cfg["founders"], cfg["n_loci"][c]
    nw = (nl + 31) // 32
    w = rng.integers(0, 2 ** 32, size=(nh, nw), dtype=np.uint32)
    if nl % 32:
        w[:, -1] &= np.uint32((1 << (nl % 32)) - 1)
    for f in range(cfg["n_phen"]):
        cv = cfg["cvs_multi"][f][c]
        for k, s in enumerate(cv["idx"]):
            word, bit = int(s) // 32, np.uint32(1 << (int(s) % 32))
            w[:, word] = (w[:, word] & ~bit) | (cv["val"][pop][:, k].astype(np.uint32) * bit)
    return w


def shard_pieces(cfg, chrs_local=None, pieces=None):
    """The (chromosome index, first locus, end locus) pieces a context owns: whole chromosomes (chrs_local, or everything) or
    locus ranges from dist.assign_locus_ranges (a chromosome then appears on several ranks, each holding a slice of its loci,
    all of them with the whole genetic map — crossovers are drawn per chromosome from Philox counters keyed by its global id)."""
    if pieces is not None:
        return [tuple(int(x) for x in pc) for pc in pieces]
    chrs_local = list(range(len(cfg["chrs"]))) if chrs_local is None else list(chrs_local)
    return [(c, 0, cfg["n_loci"][c]) for c in chrs_local]


def _cv_slice(cv, s0, s1, val):
    keep = (cv["idx"] >= s0) & (cv["idx"] < s1)
    return cv["bp"][keep], cv["a"][keep], cv["d"][keep], np.ascontiguousarray(val[:, keep])


def configure_engine_multipop(eng, cfg, chrs_local=None, panel_seed=7, pieces=None):
    """Multi-population workloads (config 4): every population its own founders, sizes and (here identical) effect sizes;
    phenotype f has omega = 1 / (1 + f), lambda = 1 for the first phenotype only."""
    pcs = shard_pieces(cfg, chrs_local, pieces)
    if pcs != [(c, 0, cfg["n_loci"][c]) for c in range(len(cfg["chrs"]))]:
        eng.set_chromosome_ids([c for c, _, _ in pcs])
    for k, (c, s0, s1) in enumerate(pcs):
        eng.set_loci(k, cfg["loci"][c][s0:s1])
    for p in range(len(cfg["pops"])):
        eng.set_population(p, avoid_inbreeding=False, random_mating=cfg["rm"], mm_percent=0.0)
        for k, (c, s0, s1) in enumerate(pcs):
            bp, cm, pr = cfg["maps"][c]
            eng.set_genetic_map(p, k, bp, pr, int(bp[1] - bp[0]))
            w = founder_words_multi(cfg, c, p, np.random.default_rng([panel_seed, c, p]))
            eng.set_founder_panel_packed(p, k, w[:, s0 // 32:(s1 + 31) // 32])
            for f in range(cfg["n_phen"]):
                cv = cfg["cvs_multi"][f][c]
                eng.set_cv(p, f, k, *_cv_slice(cv, s0, s1, cv["val"][p]))
        for f in range(cfg["n_phen"]):
            eng.set_pheno_scheme(p, f, va=0.5, vd=0.0, ve=0.5, vc=0.0, vf=0.0, omega=1.0 / (1 + f), beta=0.0, lam=1.0 if f == 0 else 0.0)


def configure_engine(eng, cfg, va=0.5, vd=0.0, ve=0.5, panel_seed=7, chrs_local=None, pieces=None):
    """Feeds the workload to an engine.  chrs_local: indices (into cfg['chrs']) of the chromosomes this context owns
    (chromosome sharding, what the segment representation uses); pieces: locus ranges (dist.assign_locus_ranges, bit-packed
    rows).  The founder panel of a chromosome does not depend on the sharding."""
    pcs = shard_pieces(cfg, chrs_local, pieces)
    if pcs != [(c, 0, cfg["n_loci"][c]) for c in range(len(cfg["chrs"]))]:
        eng.set_chromosome_ids([c for c, _, _ in pcs])
    segments_only = bool(cfg.get("segments"))
    for k, (c, s0, s1) in enumerate(pcs):
        if not segments_only:
            eng.set_loci(k, cfg["loci"][c][s0:s1])
    eng.set_population(0, avoid_inbreeding=False, random_mating=cfg["rm"], mm_percent=0.0)
    for k, (c, s0, s1) in enumerate(pcs):
        bp, cm, p = cfg["maps"][c]
        eng.set_genetic_map(0, k, bp, p, int(bp[1] - bp[0]))
        if not segments_only:
            w = founder_words(cfg, c, np.random.default_rng([panel_seed, c]))
            eng.set_founder_panel_packed(0, k, w[:, s0 // 32:(s1 + 31) // 32])
        cv = cfg["cvs"][c]
        if segments_only:
            eng.set_cv(0, 0, k, cv["bp"], cv["a"], cv["d"], cv["val"])
        else:
            eng.set_cv(0, 0, k, *_cv_slice(cv, s0, s1, cv["val"]))
    eng.set_pheno_scheme(0, 0, va=va, vd=vd, ve=ve, vc=0.0, vf=0.0, omega=1.0, beta=0.0, lam=1.0)


def write_reference_inputs(cfg, d, n_gen, tag="w", selection=("logit", 0, 1)):
    """The same workload in the reference's text formats; returns the CLI arguments.  The founder .hap files
    hold ONE row (the reference only counts their columns during the loop, src/Simulation.cpp:307-311)."""
    os.makedirs(d, exist_ok=True)
    chrs, nh = cfg["chrs"], 2 * cfg["founders"]
    with open(f"{d}/{tag}.rmap", "w") as f:
        f.write("chr bp cM\n")
        for c, (bp, cm, _) in zip(chrs, cfg["maps"]):
            f.write("".join(f"{c} {int(b)} {x:.12g}\n" for b, x in zip(bp, cm)))
    with open(f"{d}/{tag}.indv", "w") as f:
        f.write("".join(f"id{i + 1}\n" for i in range(cfg["founders"])))
    with open(f"{d}/{tag}.hapaddr", "w") as f:
        f.write("chr hap legend sample\n")
        for c in chrs:
            f.write(f"{c} {d}/{tag}.chr{c}.hap {d}/{tag}.chr{c}.legend {d}/{tag}.indv\n")
    with open(f"{d}/{tag}.cvinfo", "w") as fi, open(f"{d}/{tag}.cvs", "w") as fc:
        fi.write("chr pos a d\n")
        for c, cv in zip(chrs, cfg["cvs"]):
            with open(f"{d}/{tag}.chr{c}.hap", "w") as f:
                f.write("0 " * nh + "\n")
            with open(f"{d}/{tag}.chr{c}.legend", "w") as f:
                f.write("id pos allele0 allele1\nrs1 1 A C\n")
            with open(f"{d}/{tag}.cv.chr{c}.hap", "w") as f:
                for k in range(len(cv["bp"])):
                    f.write(" ".join(map(str, cv["val"][:, k])) + " \n")
            fi.write("".join(f"{c} {int(b)} {a:.10g} {dd:.10g}\n" for b, a, dd in zip(cv["bp"], cv["a"], cv["d"])))
            fc.write(f"{c} {d}/{tag}.cv.chr{c}.hap\n")
    with open(f"{d}/{tag}.gen", "w") as f:
        f.write("pop_size mat_cor offspring_dist selection_func selection_func_par1 selection_func_par2\n")
        for _ in range(n_gen):
            f.write(f"{cfg['n']} {cfg['mat_cor']} p {selection[0]} {selection[1]} {selection[2]}\n")
    args = ["--file_gen_info", f"{d}/{tag}.gen", "--file_hap_name", f"{d}/{tag}.hapaddr", "--file_recom_map", f"{d}/{tag}.rmap",
            "--file_cv_info", f"{d}/{tag}.cvinfo", "--file_cvs", f"{d}/{tag}.cvs", "--va", "0.5", "--vd", "0", "--ve", "0.5"]
    if cfg["rm"]:
        args.append("--RM")
    return args
