"""Edge cases and size-independent properties of the CUDA path (-m gpu), through the C-ABI.

Part 1 — adversarial small cases against the (pinned) CPU oracle under caller-supplied draws: chromosome lengths
around the 32-bit word / 128-bit chunk / 8 KB tile boundaries (1, 31, 32, 33, 127, 128, 129, 4097 … loci), crossovers
exactly on locus positions, on the first and last map rows, duplicated, hundreds per gamete (beyond the shared-memory
staging of propagate_bits_kernel), empty crossover lists, a single offspring, couples without offspring.
Part 2 — error behaviour of the boundary (capacity, empty mating lists, bad migration rows, bad indices).
Part 3 — properties at the bench's full size (100 000 individuals x 1 000 000 loci), where no CPU oracle can follow:
the causal-variant bit planes and the bit-packed haplotype rows are propagated by two independent kernels from the
same draws, so the rows read at the CV loci must equal the CV planes; with recombination switched off every
offspring row must be a verbatim copy of the drawn parental row.
"""
import numpy as np
import pytest

from geneevolve_b200 import capi, workloads
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu


class Case:
    """A random single-population input with chosen chromosome sizes, plus caller-supplied draws."""

    def __init__(self, seed, n_loci, n_founders=12, map_rows=9, step=64, n_cv=5):
        rng = np.random.default_rng(seed)
        self.rng, self.n_chr, self.nf = rng, len(n_loci), n_founders
        self.maps, self.loci, self.panel, self.cv = [], [], [], []
        for c, nl in enumerate(n_loci):
            bp = 1000 + step * np.arange(map_rows, dtype=np.uint64)
            p = np.concatenate([[0.0], rng.uniform(0.0, 0.2, map_rows - 1)])
            lo, hi = int(bp[0]) - 20, int(bp[-1]) + 20          # some loci outside the covered range
            if nl <= hi - lo:
                pos = np.sort(rng.choice(np.arange(lo, hi), size=nl, replace=False)).astype(np.uint64)
            else:                                               # more loci than bp: duplicates are legal positions
                pos = np.sort(rng.integers(lo, hi, size=nl)).astype(np.uint64)
            hap = (rng.uniform(size=(2 * n_founders, nl)) < 0.5).astype(np.uint8)
            k = min(n_cv, nl)
            idx = np.sort(rng.choice(nl, size=k, replace=False))
            self.maps.append((bp, p, step)); self.loci.append(pos); self.panel.append(hap)
            self.cv.append(dict(bp=pos[idx], a=rng.normal(size=k), d=rng.normal(size=k) * 0.3, val=hap[:, idx]))

    def kwargs(self, cap, **over):
        kw = dict(n_pop=1, n_chr=self.n_chr, n_phen=1, seed=11, capacity=cap, rng_mode=capi.GE_RNG_REPLAY)
        kw.update(over)
        return kw

    def configure(self, e):
        for c in range(self.n_chr):
            e.set_loci(c, self.loci[c])
        e.set_population(0, False, False, 0.0)
        for c in range(self.n_chr):
            bp, p, step = self.maps[c]
            e.set_genetic_map(0, c, bp, p, step)
            e.set_founder_panel(0, c, self.panel[c])
            e.set_cv(0, 0, c, self.cv[c]["bp"], self.cv[c]["a"], self.cv[c]["d"], self.cv[c]["val"])
        e.set_pheno_scheme(0, 0, va=0.5, vd=0.1, ve=0.4)

    def draws0(self):
        n = self.nf
        return capi.Draws(n, sex=(np.arange(n) % 2 + 1).astype(np.uint8), e_raw=self.rng.normal(size=(1, n)))

    def draws(self, n_par, n_off, xo_lists):
        """xo_lists(slot, chromosome) -> sorted crossover positions of that gamete."""
        rng, C = self.rng, self.n_chr
        off, bp = [0], []
        for s in range(n_off * C * 2):
            x = list(xo_lists(s, (s // 2) % C))
            bp += x
            off.append(len(bp))
        return capi.Draws(n_off, father=rng.integers(0, n_par, n_off), mother=rng.integers(0, n_par, n_off),
                          sex=rng.integers(1, 3, n_off).astype(np.uint8), xo_off=np.array(off, np.uint64), xo_bp=np.array(bp, np.uint64),
                          start_hap=rng.integers(0, 2, n_off * C * 2).astype(np.uint8), e_raw=rng.normal(size=(1, n_off)))


def run_pair(cuda_lib, case, gens, rep=capi.GE_REP_BITS | capi.GE_REP_SEGMENTS, cap=64, flags=0):
    """gens: list of (n_off, xo_lists).  Runs the library and the oracle on identical draws, compares everything."""
    gpu = capi.Engine(cuda_lib, **case.kwargs(cap, representation=rep, flags=flags))
    cpu = OracleEngine(**case.kwargs(cap))
    d0 = case.draws0()
    for e in (gpu, cpu):
        case.configure(e)
        e.init_generation0([d0])
    gp = [capi.gen_params(10)]
    n_par = case.nf
    for g, (n_off, xo) in enumerate(gens, 1):
        d = case.draws(n_par, n_off, xo)
        for e in (gpu, cpu):
            e.step_generation(g, gp, None, [d])
        for c in range(case.n_chr):
            assert np.array_equal(gpu.haplotypes(0, c), cpu.haplotypes(0, c)), f"gen {g} chr {c}: haplotypes"
            assert np.array_equal(gpu.haplotypes(0, c), cpu.haplotypes_from_segments(0, c)), f"gen {g} chr {c}: bits vs segments"
            assert np.array_equal(gpu.cv_alleles(0, 0, c), cpu.cv_alleles(0, 0, c)), f"gen {g} chr {c}: CV alleles"
            if rep & capi.GE_REP_SEGMENTS:
                assert np.array_equal(gpu.segments(0, c)["seg"], cpu.segments(0, c)["seg"])
        a, b = gpu.individuals(0), cpu.individuals(0)
        assert np.array_equal(a["ids"], b["ids"])
        for k in "ADGEP":
            np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-11)
        n_par = n_off


@pytest.mark.parametrize("n_loci", [[1], [31, 32, 33], [127, 128, 129], [255, 257, 1], [4097, 5], [65536 + 1, 200]])
def test_word_chunk_and_tile_boundaries(cuda_lib, n_loci):
    case = Case(1000 + sum(n_loci), n_loci)

    def xo(slot, c):
        bp, _, step = case.maps[c]
        r = case.rng
        k = int(r.integers(0, 4))
        return np.sort(r.integers(int(bp[0]), int(bp[-1]) + step, size=k))

    run_pair(cuda_lib, case, [(9, xo), (1, xo), (17, xo)])


def test_crossovers_on_loci_map_rows_and_duplicates(cuda_lib):
    case = Case(7, [100, 40])

    def xo(slot, c):
        bp, _, step = case.maps[c]
        pos = case.loci[c]
        choice = slot % 6
        if choice == 0:
            return []                                            # the chosen parental haplotype unchanged (:2910)
        if choice == 1:
            return [int(bp[0])]                                  # on the first map row: flips everything
        if choice == 2:
            return [int(bp[-1]), int(bp[-1]) + step - 1]         # last-row quirk: positions at/after the covered end
        if choice == 3:
            p = int(pos[len(pos) // 2])
            return [p, p]                                        # duplicate on a locus: two flips, no net change
        if choice == 4:
            return sorted(int(x) for x in pos[[3, 4, 5]])        # on three consecutive loci
        return sorted(int(x) for x in case.rng.choice(pos, size=7))

    run_pair(cuda_lib, case, [(24, xo), (24, xo)])


def test_hundreds_of_crossovers_per_gamete(cuda_lib):
    """More flips per offspring than propagate_bits_kernel stages in shared memory (global-memory fallback)."""
    case = Case(9, [3000], map_rows=40, step=100)

    def xo(slot, c):
        bp, _, step = case.maps[c]
        return np.sort(case.rng.integers(int(bp[0]), int(bp[-1]), size=260))

    run_pair(cuda_lib, case, [(6, xo), (5, xo)], rep=capi.GE_REP_BITS)


@pytest.mark.parametrize("order", ["sorted", "reversed", "shuffled"])
def test_causal_variant_blocks_of_several_words_in_any_row_order(cuda_lib, order):
    """70 and 40 causal variants on two chromosomes (blocks of 3 and 2 words of the CV bit planes), listed in ascending
    position (the warp-per-row kernel with its prefix-XOR masks), reversed and shuffled (cv.info rows need not be sorted:
    the thread-per-word kernel without binary search), many crossovers per gamete so that words see several flips."""
    case = Case(31, [300, 40], n_founders=10, map_rows=12, n_cv=70)
    if order != "sorted":
        for cv in case.cv:
            k = len(cv["bp"])
            perm = np.arange(k)[::-1] if order == "reversed" else np.random.default_rng(5).permutation(k)
            cv["bp"], cv["a"], cv["d"], cv["val"] = cv["bp"][perm], cv["a"][perm], cv["d"][perm], cv["val"][:, perm]
    lo, hi = 1000, 1000 + 64 * 11

    def xo(slot, c):
        r = np.random.default_rng([slot, c])
        return np.sort(r.integers(lo - 5, hi + 5, size=int(r.integers(0, 9))))
    run_pair(cuda_lib, case, [(14, xo), (9, xo), (11, xo)])


@pytest.mark.parametrize("flags", [0, capi.GE_FLAG_SEG_WIDE_PARTS, capi.GE_FLAG_SEG_VERBATIM])
def test_segment_kernels_long_lists_and_many_crossovers(cuda_lib, flags):
    """The kernels of the segment path (seg_plan_kernel + seg_gather_kernel on packed and on 16-byte parts, and the reference's loop
    verbatim) against the oracle's `recombine`: lists that grow past several 32-part chunks, slots without crossovers, crossovers on
    map rows / loci / duplicated, beyond the covered end, and slots with more than 30 crossovers."""
    case = Case(31, [900, 60, 7], map_rows=50, step=128)

    def xo(slot, c):
        bp, _, step = case.maps[c]
        r = case.rng
        choice = slot % 7
        if choice == 0:
            return []
        if choice == 1:
            return [int(bp[0])]
        if choice == 2:
            p = int(case.loci[c][len(case.loci[c]) // 2])
            return [p, p, int(bp[-1])]
        if choice == 3:
            return np.sort(r.integers(int(bp[0]), int(bp[-1]) + step, size=45))      # > 30 crossovers: verbatim fallback
        return np.sort(r.integers(int(bp[0]), int(bp[-1]) + step, size=int(r.integers(1, 14))))

    run_pair(cuda_lib, case, [(20, xo)] * 9, cap=64, flags=flags)


def ibd_numpy(seg, off, i, j, min_bp):
    """Restatement of ge_ibd_sharing on the downloaded lists: per haplotype the covered positions are labelled with their
    founder (hap_index, root), then runs of equal labels are measured position by position (small genomes only)."""
    tot = runs = 0
    for ha in range(2):
        for hb in range(2):
            A, B = seg[off[2 * i + ha]:off[2 * i + ha + 1]], seg[off[2 * j + hb]:off[2 * j + hb + 1]]
            lo, hi = int(min(A[0, 0], B[0, 0])), int(max(A[-1, 1], B[-1, 1]))
            la, lb = np.full(hi - lo, -1, np.int64), np.full(hi - lo, -2, np.int64)
            for L, q in ((la, A), (lb, B)):
                for st, en, z, w in q:
                    L[int(st) - lo:int(en) - lo] = int(z) * 64 + int(w)
            same = np.concatenate([[0], (la == lb).astype(np.int8), [0]])
            d = np.diff(same)
            for s0, s1 in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1)):   # maximal stretches on which the two agree
                if s1 - s0 >= min_bp:
                    tot += int(s1 - s0)
                    runs += 1
    return tot, runs


@pytest.mark.parametrize("fmt", ["packed", "16"])
def test_ibd_sharing_matches_a_position_by_position_count(cuda_lib, fmt):
    case = Case(77, [120, 30], n_founders=10, map_rows=12, step=40)
    gpu = capi.Engine(cuda_lib, **case.kwargs(64, representation=capi.GE_REP_SEGMENTS, flags=capi.GE_FLAG_SEG_WIDE_PARTS if fmt == "16" else 0))
    case.configure(gpu)
    gpu.init_generation0([case.draws0()])
    assert gpu.segment_format() == (8 if fmt == "packed" else 16)

    def xo(slot, c):
        bp, _, step = case.maps[c]
        return np.sort(case.rng.integers(int(bp[0]), int(bp[-1]), size=int(case.rng.integers(0, 4))))

    n_par = case.nf
    for g in range(1, 5):      # small populations: plenty of shared ancestry after four generations
        gpu.step_generation(g, [capi.gen_params(10)], None, [case.draws(n_par, 14, xo)])
        n_par = 14
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.integers(0, 14, 40), np.arange(14)])
    b = np.concatenate([rng.integers(0, 14, 40), np.arange(14)])      # the last 14 pairs are individuals with themselves
    for c in range(2):
        s = gpu.segments(0, c)
        for min_bp in (0, 25):
            tot, runs = gpu.ibd_sharing(0, c, a, b, min_bp)
            want = [ibd_numpy(s["seg"].astype(np.int64), s["seg_off"].astype(np.int64), int(i), int(j), min_bp) for i, j in zip(a, b)]
            assert [int(t) for t in tot] == [w[0] for w in want] and [int(r) for r in runs] == [w[1] for w in want], (c, min_bp)
        assert tot[:40].max() > 0   # unrelated pairs share something, too, after four generations of ten founders


def test_segment_capacity_is_enforced_on_both_paths(cuda_lib):
    """seg_capacity smaller than the lists: GE_ERR_CAPACITY from the generation itself or, when the chain is queued on the bulk
    stream and finishes after the generation's read-back, from the first call after it — never a write beyond the buffer."""
    for flags, cap_parts in ((capi.GE_FLAG_SEG_VERBATIM, 40), (0, 40), (capi.GE_FLAG_SERIAL, 40)):
        case = Case(5, [200, 50])
        gpu = capi.Engine(cuda_lib, **case.kwargs(64, representation=capi.GE_REP_SEGMENTS, seg_capacity=cap_parts, flags=flags))
        case.configure(gpu)
        gpu.init_generation0([case.draws0()])

        def xo(slot, c):
            bp, _, step = case.maps[c]
            return np.sort(case.rng.integers(int(bp[0]), int(bp[-1]), size=3))

        with pytest.raises(capi.GeneEvolveError) as e:
            gpu.step_generation(1, [capi.gen_params(10)], None, [case.draws(case.nf, 20, xo)])
            gpu.segments(0, 0)
        assert "seg_capacity" in str(e.value)


def test_philox_many_crossovers_matches_oracle(cuda_lib):
    """Recombination rates high enough that most gametes exceed the per-slot stash of sample_xo_kernel."""
    case = Case(21, [300, 90], map_rows=60, step=64)
    case.maps = [(bp, np.concatenate([[0.0], np.full(len(bp) - 1, 0.35)]), step) for bp, _, step in case.maps]
    kw = case.kwargs(80, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS)
    gpu, cpu = capi.Engine(cuda_lib, **kw), OracleEngine(**kw)
    for e in (gpu, cpu):
        case.configure(e)
        e.set_population(0, False, True, 0.0)   # random mating
        e.init_generation0()
    for g in range(1, 4):
        gp = [capi.gen_params(40, 0.0, "p", "thr", 1, 1)]
        for e in (gpu, cpu):
            e.step_generation(g, gp)
        da, db = gpu.draws(0), cpu.draws(0)
        assert len(da["xo_bp"]) / (40 * 2 * 2) > 8          # mean crossovers per gamete-chromosome
        for k in ("father", "mother", "xo_off", "xo_bp", "start_hap", "sex"):
            assert np.array_equal(da[k], db[k]), k
        for c in range(2):
            assert np.array_equal(gpu.haplotypes(0, c), cpu.haplotypes(0, c))


def test_boundary_errors(cuda_lib):
    case = Case(3, [50])
    e = capi.Engine(cuda_lib, **case.kwargs(16, representation=capi.GE_REP_BITS))
    case.configure(e)
    e.init_generation0([case.draws0()])
    gp = [capi.gen_params(10)]
    none = lambda s, c: []  # noqa: E731
    with pytest.raises(capi.GeneEvolveError) as ei:          # more offspring than the capacity given at ge_create
        e.step_generation(1, gp, None, [case.draws(case.nf, 17, none)])
    assert ei.value.code == -3
    d = case.draws(case.nf, 4, none)
    d.arrays["father"][2] = case.nf                            # parent index outside the parent generation
    with pytest.raises(capi.GeneEvolveError) as ei:
        e.step_generation(1, gp, None, [d])
    assert ei.value.code == -1
    with pytest.raises(capi.GeneEvolveError):                  # replay context cannot mate on its own
        e.mate(0, 1, gp[0])
    # Philox context: nobody may mate when the selection function is 0 for everyone (thr with p1 = 0)
    kw = case.kwargs(64, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS)
    p = capi.Engine(cuda_lib, **kw)
    case.configure(p)
    p.init_generation0()
    p.step_generation(1, [capi.gen_params(20, 0.0, "p", "thr", 0.0, 1e9)])   # evaluated at the END of generation 1
    with pytest.raises(capi.GeneEvolveError) as ei:
        p.step_generation(2, [capi.gen_params(20)])
    assert ei.value.code == -4 and "couples=0" in str(ei.value)
    with pytest.raises(capi.GeneEvolveError):
        capi.Engine(cuda_lib, n_pop=1, n_chr=1, n_phen=1, capacity=0)


def test_migration_row_must_sum_to_one(cuda_lib):
    case = Case(5, [40])
    kw = case.kwargs(64, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_BITS, n_pop=2)
    e = capi.Engine(cuda_lib, **kw)
    for c in range(case.n_chr):
        e.set_loci(c, case.loci[c])
    for p in range(2):
        e.set_population(p, False, True, 0.0)
        bp, pr, step = case.maps[0]
        e.set_genetic_map(p, 0, bp, pr, step)
        e.set_founder_panel(p, 0, case.panel[0])
        e.set_cv(p, 0, 0, case.cv[0]["bp"], case.cv[0]["a"], case.cv[0]["d"], case.cv[0]["val"])
        e.set_pheno_scheme(p, 0, va=0.5, vd=0.0, ve=0.5)
    e.init_generation0()
    gp = [capi.gen_params(20, 0.0, "p", "thr", 1, 1)] * 2
    with pytest.raises(capi.GeneEvolveError) as ei:
        e.step_generation(1, gp, [0.9, 0.2, 0.1, 0.9])
    assert ei.value.code == -6
    e2 = capi.Engine(cuda_lib, **kw)  # a valid ring row moves round(0.1 * n) individuals each way
    for c in range(case.n_chr):
        e2.set_loci(c, case.loci[c])
    for p in range(2):
        e2.set_population(p, False, True, 0.0)
        bp, pr, step = case.maps[0]
        e2.set_genetic_map(p, 0, bp, pr, step)
        e2.set_founder_panel(p, 0, case.panel[0])
        e2.set_cv(p, 0, 0, case.cv[0]["bp"], case.cv[0]["a"], case.cv[0]["d"], case.cv[0]["val"])
        e2.set_pheno_scheme(p, 0, va=0.5, vd=0.0, ve=0.5)
    e2.init_generation0()
    e2.step_generation(1, gp, [0.9, 0.1, 0.2, 0.8])
    assert (e2.population_size(0), e2.population_size(1)) == (20 - 2 + 4, 20 - 4 + 2)
    # both populations share one effect table here, so the library carries no root-population plane and uses the
    # tabulated genetic values; the oracle always looks a, d up through the root population (:2776-2786): same numbers
    cpu = OracleEngine(**kw)
    for c in range(case.n_chr):
        cpu.set_loci(c, case.loci[c])
    for p in range(2):
        cpu.set_population(p, False, True, 0.0)
        bp, pr, step = case.maps[0]
        cpu.set_genetic_map(p, 0, bp, pr, step)
        cpu.set_founder_panel(p, 0, case.panel[0])
        cpu.set_cv(p, 0, 0, case.cv[0]["bp"], case.cv[0]["a"], case.cv[0]["d"], case.cv[0]["val"])
        cpu.set_pheno_scheme(p, 0, va=0.5, vd=0.0, ve=0.5)
    cpu.init_generation0()
    cpu.step_generation(1, gp, [0.9, 0.1, 0.2, 0.8])
    for g in range(2, 5):
        e2.step_generation(g, gp, [0.9, 0.1, 0.2, 0.8])
        cpu.step_generation(g, gp, [0.9, 0.1, 0.2, 0.8])
    for p in range(2):
        a, b = e2.individuals(p), cpu.individuals(p)
        assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["sex"], b["sex"])
        for k in "ADGEP":
            np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-11)
        assert np.array_equal(e2.haplotypes(p, 0), cpu.haplotypes(p, 0))


# ---------------------------------------------------------------------------------------------------------------
# full size
# ---------------------------------------------------------------------------------------------------------------
def full_size_engine(cuda_lib, zero_recombination=False, n=100000):
    cfg = workloads.make_workload("config3_100k_x_1M", n_override=n)
    if zero_recombination:
        cfg["maps"] = [(bp, cm, np.zeros_like(p)) for bp, cm, p in cfg["maps"]]
    cap = int(max(cfg["n"], cfg["founders"]) * 1.03) + 1024
    eng = capi.Engine(cuda_lib, n_pop=1, n_chr=len(cfg["chrs"]), n_phen=1, representation=capi.GE_REP_BITS, rng_mode=capi.GE_RNG_PHILOX,
                      seed=99, capacity=cap)
    workloads.configure_engine(eng, cfg)
    eng.init_generation0()
    return eng, cfg


def test_full_size_rows_agree_with_cv_planes(cuda_lib):
    """100k x 1M, three generations with assortative mating and selection: at every causal variant the bit-packed row
    (propagate_bits_kernel) and the CV plane (cv_propagate_bits_kernel) must hold the same allele."""
    eng, cfg = full_size_engine(cuda_lib)
    gp = [capi.gen_params(cfg["n"], cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    for g in range(1, 4):
        eng.step_generation(g, gp)
    n = eng.population_size(0)
    assert abs(n - cfg["n"]) < 6 * np.sqrt(cfg["n"])
    for c in (len(cfg["chrs"]) - 1, 0, 10):          # the shortest, the longest and a middle chromosome
        words = eng.haplotypes_packed(0, c)
        idx = cfg["cvs"][c]["idx"].astype(np.int64)
        at_cv = ((words[:, idx // 32] >> (idx % 32).astype(np.uint32)) & 1).astype(np.uint8)
        cv = eng.cv_alleles(0, 0, c)
        assert at_cv.shape == cv.shape == (2 * n, len(idx))
        assert np.array_equal(at_cv, cv), f"chromosome index {c}"
        f = words.view(np.uint8).reshape(2 * n, -1)
        assert 0.3 < np.unpackbits(f[:2000], axis=1).mean() < 0.7   # rows still look like random founder mosaics
    ind = eng.individuals(0)
    assert np.isfinite(ind["P"]).all() and 0.3 < np.var(ind["A"][0]) / np.var(ind["P"][0]) < 0.7


def test_north_star_target_hundred_generations(cuda_lib):
    """The north-star run itself — 100 000 individuals x 1 000 000 loci x 100 generations, assortative mating rho = 0.4 with
    logit selection — must finish in seconds and still satisfy the two-path invariant at generation 100: rows read at
    the causal loci equal the causal-variant planes (any single wrong run boundary in any of the 100 copies of any
    ancestor would break it for that lineage)."""
    import time
    eng, cfg = full_size_engine(cuda_lib)
    gp = [capi.gen_params(cfg["n"], cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    eng.synchronize()
    t0 = time.perf_counter()
    for g in range(1, 101):
        eng.step_generation(g, gp)
    eng.synchronize()
    wall = time.perf_counter() - t0
    assert wall < 5.0, f"100 generations took {wall:.2f} s"
    n = eng.population_size(0)
    assert abs(n - cfg["n"]) < 6 * np.sqrt(cfg["n"])
    for c in (len(cfg["chrs"]) - 1, 7):
        words = eng.haplotypes_packed(0, c)
        idx = cfg["cvs"][c]["idx"].astype(np.int64)
        at_cv = ((words[:, idx // 32] >> (idx % 32).astype(np.uint32)) & 1).astype(np.uint8)
        assert np.array_equal(at_cv, eng.cv_alleles(0, 0, c)), f"chromosome index {c} after 100 generations"
    ind = eng.individuals(0)
    assert np.isfinite(ind["P"]).all() and np.array_equal(ind["ids"][:, 0], np.arange(n, dtype=np.uint64))
    h2 = np.var(ind["A"][0]) / np.var(ind["P"][0])
    assert 0.2 < h2 < 0.8 and abs(np.var(ind["E"][0], ddof=1) - 0.5) < 1e-9     # var(E) is rescaled to ve exactly every generation (:3167-3180)
    print(f"north-star target: 100 generations of {cfg['n']} x {sum(cfg['n_loci'])} in {wall:.2f} s ({wall * 10:.2f} ms per generation), h2 = {h2:.3f}")


def test_full_width_rows_are_verbatim_parental_copies_without_recombination(cuda_lib):
    """20k x 1M with every recombination probability zero: each offspring row equals the drawn parental row."""
    eng, cfg = full_size_engine(cuda_lib, zero_recombination=True, n=20000)
    gp = [capi.gen_params(cfg["n"], cfg["mat_cor"], "p", "logit", 0.0, 1.0)]
    eng.step_generation(1, gp)
    C = len(cfg["chrs"])
    for c in (C - 1, 3):
        parents = eng.haplotypes_packed(0, c)
        eng.step_generation(2 if c == C - 1 else 3, gp)
        d = eng.draws(0)
        assert len(d["xo_bp"]) == 0
        kids = eng.haplotypes_packed(0, c)
        n_off = len(d["father"])
        slot = (np.arange(n_off) * C + c) * 2
        src_f = 2 * d["father"].astype(np.int64) + d["start_hap"][slot]
        src_m = 2 * d["mother"].astype(np.int64) + d["start_hap"][slot + 1]
        assert np.array_equal(kids[0::2], parents[src_f]) and np.array_equal(kids[1::2], parents[src_m])


def test_segment_compaction_preserves_haplotypes(cuda_lib):
    """ge_compact_segments (extension): two segment-only contexts on the same Philox streams, one compacted
    after every generation.  A small closed population drifts towards few founders, so adjacent same-founder parts and
    zero-length parts do occur; materialised haplotypes, CV alleles and phenotypes must stay identical, the lists shrink,
    and a second compaction changes nothing."""
    case = Case(77, [400, 150], n_founders=8, map_rows=30, step=64)
    case.maps = [(bp, np.concatenate([[0.0], np.full(len(bp) - 1, 0.08)]), step) for bp, _, step in case.maps]
    kw = case.kwargs(80, rng_mode=capi.GE_RNG_PHILOX, representation=capi.GE_REP_SEGMENTS)
    plain, packed = capi.Engine(cuda_lib, **kw), capi.Engine(cuda_lib, **kw)
    for e in (plain, packed):
        case.configure(e)
        e.set_population(0, False, True, 0.0)
        e.init_generation0()
    shrunk = 0
    for g in range(1, 13):
        gp = [capi.gen_params(30, 0.0, "p", "thr", 1, 1)]
        plain.step_generation(g, gp)
        packed.step_generation(g, gp)
        before, after = packed.compact_segments(0)
        assert after <= before and packed.compact_segments(0) == (after, after)
        shrunk += before - after
        for c in range(2):
            assert np.array_equal(plain.haplotypes(0, c), packed.haplotypes(0, c)), f"gen {g} chr {c}"
            assert np.array_equal(plain.cv_alleles(0, 0, c), packed.cv_alleles(0, 0, c))
            s = packed.segments(0, c)
            seg, off = s["seg"], s["seg_off"]
            for r in range(len(off) - 1):     # still a contiguous tiling, no mergeable neighbours left
                q = seg[off[r]:off[r + 1]]
                assert np.all(q[1:, 0] == q[:-1, 1]) and not np.any((q[1:, 2] == q[:-1, 2]) & (q[1:, 3] == q[:-1, 3]))
        a, b = plain.individuals(0), packed.individuals(0)
        for k in "ADGEP":
            assert np.array_equal(a[k], b[k])
    assert shrunk > 0


def test_caller_arrays_are_validated_before_any_kernel_reads_them(cuda_lib):
    """The C-ABI refuses bad caller input with GE_ERR_INVALID before any state changes (no device read through a bad index): couple
    positions outside the generation, negative family sizes, CSR offsets that do not start at 0 or decrease, crossover positions
    that do not ascend (bit-packed rows need them monotone), mutation hits without a mutation map, maps changed after generation 0
    and map rows with probability >= 1 (GE_ERR_UNSUPPORTED)."""
    case = Case(9, [70, 33])
    e = capi.Engine(cuda_lib, **case.kwargs(32, representation=capi.GE_REP_BITS))
    case.configure(e)
    e.init_generation0([case.draws0()])
    n = case.nf
    ok = dict(pos_male=[0, 1], pos_female=[1, 0], inbreed=[0, 0], num_offspring=[1, 2])
    for bad in (dict(pos_male=[0, n]), dict(pos_female=[n + 5, 0]), dict(num_offspring=[1, -1])):
        with pytest.raises(capi.GeneEvolveError) as ei:
            e.set_couples(0, **{**ok, **bad})
        assert ei.value.code == -1
    e.set_couples(0, **ok)
    none = lambda s, c: []  # noqa: E731
    gp = [capi.gen_params(10)]

    def broken(mutate):
        d = case.draws(n, 6, lambda s, c: [int(case.maps[c][0][1]) + 3, int(case.maps[c][0][2]) + 5])
        mutate(d.arrays)
        with pytest.raises(capi.GeneEvolveError) as ei:
            e.step_generation(1, gp, None, [d])
        assert ei.value.code == -1, str(ei.value)
        assert e.population_size(0) == n        # nothing happened

    broken(lambda a: a["xo_off"].__setitem__(0, 1))                                   # offsets must start at 0
    broken(lambda a: a["xo_off"].__setitem__(3, int(a["xo_off"][2]) - 1))             # ... and never decrease
    broken(lambda a: a["xo_bp"].__setitem__(slice(0, 2), a["xo_bp"][:2][::-1].copy()))  # positions of a gamete must ascend
    broken(lambda a: a.__setitem__("mut_off", np.zeros(6 * case.n_chr + 1, np.uint64)))   # hits without a mutation map
    e.step_generation(1, gp, None, [case.draws(n, 6, none)])                          # the context is still usable
    assert e.population_size(0) == 6
    bp, pr, step = case.maps[0]
    with pytest.raises(capi.GeneEvolveError) as ei:                                    # maps are frozen at generation 0
        e.set_genetic_map(0, 0, bp, pr, step)
    assert ei.value.code == -1
    f = capi.Engine(cuda_lib, **case.kwargs(32, representation=capi.GE_REP_BITS))
    p1 = np.array(pr, float)
    p1[2] = 1.0
    with pytest.raises(capi.GeneEvolveError) as ei:                                    # a 100 cM jump between two rows
        f.set_genetic_map(0, 0, bp, p1, step)
    assert ei.value.code == -7
