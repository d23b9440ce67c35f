"""CPU checks (-m "not gpu") of the two claims the segment path's plan + gather kernels rest on (geneevolve_b200/csrc/ge_segments.cuh):

1. The packed 8-byte part format drops `en` because the lists `Simulation::recombine` (src/Simulation.cpp:2903-2958) emits are
   contiguous tilings: en_i == st_{i+1}, the first part starts at the first map row, the last one ends at the last map row.
   Checked on every list the REAL reference exported into tests/golden/*.npz (every scenario, generation, population, chromosome).

2. For a sorted tiling the parts the reference's four clip branches (:2922-2954) emit for interval [L, R) on a haplotype are the
   index range [#(y <= L), max(#(y <= R), #(x < R))) of its list, each clipped to (max(x, L), min(y, R)) — also when the interval
   is reversed (a crossover below the first or beyond the last map row).  Checked against a line-by-line restatement of the
   reference's loop on random tilings with zero-length parts, duplicated crossovers, and crossovers on part boundaries and
   outside the covered range.
"""
import numpy as np
import pytest

from golden_util import SCENARIOS, Golden


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_lists_are_contiguous_tilings(name):
    G = Golden(name)
    n_lists = 0
    for gen in range(G.G + 1):
        for p in range(G.n_pop):
            for c in range(G.n_chr):
                off, seg = G.g(gen, p, f"c{c}.seg_off"), G.g(gen, p, f"c{c}.seg")
                seg = seg.reshape(-1, 4)
                first, last = off[:-1], off[1:] - 1
                assert np.all(off[1:] > off[:-1]), "empty list"
                inner = np.ones(len(seg), bool)
                inner[last] = False
                assert np.array_equal(seg[inner, 1], seg[np.flatnonzero(inner) + 1, 0]), f"{name} gen {gen}: en_i != st_(i+1)"
                assert np.all(seg[:, 0] <= seg[:, 1])
                # every list starts at the first row and ends at the last row of the genetic map of the population that owns it now;
                # populations of one scenario share the map range (checked), so the owner does not matter
                lo = {int(G[f"in.p{q}.c{c}.rmap_bp"][0]) for q in range(G.n_pop)}
                hi = {int(G[f"in.p{q}.c{c}.rmap_bp"][-1]) for q in range(G.n_pop)}
                assert len(lo) == 1 and len(hi) == 1
                assert np.all(seg[first, 0] == lo.pop()) and np.all(seg[last, 1] == hi.pop())
                n_lists += len(off) - 1
    assert n_lists > 100


def recombine_verbatim(H0, H1, xo, lo_c, hi_c, hi):
    """src/Simulation.cpp:2903-2958, branch for branch (recombination_locs = [cov_lo, crossovers..., cov_hi])."""
    if len(xo) == 0:
        return [tuple(q) for q in (H1 if hi else H0)]
    locs = [lo_c] + list(xo) + [hi_c]
    out = []
    for i1 in range(1, len(locs)):
        L, R = locs[i1 - 1], locs[i1]
        H = H1 if hi else H0
        i2 = 0
        while i2 < len(H) and H[i2][1] <= L:
            i2 += 1
        if i2 < len(H):
            x, y, z = H[i2]
            if x < L < y and R < y:
                out.append((L, R, z)); i2 += 1
        if i2 < len(H):
            x, y, z = H[i2]
            if x < L < y and R >= y:
                out.append((L, y, z)); i2 += 1
        while i2 < len(H):
            x, y, z = H[i2]
            if not (y <= R and L <= x):
                break
            out.append((x, y, z)); i2 += 1
        if i2 < len(H):
            x, y, z = H[i2]
            if x < R < y:
                out.append((x, R, z))
        hi ^= 1
    return out


def recombine_by_index_ranges(H0, H1, xo, lo_c, hi_c, hi):
    """What seg_plan_kernel + seg_gather_kernel compute."""
    if len(xo) == 0:
        return [tuple(q) for q in (H1 if hi else H0)]
    locs = [lo_c] + list(xo) + [hi_c]
    out = []
    for j in range(len(locs) - 1):
        L, R = locs[j], locs[j + 1]
        H = H1 if hi else H0
        i0 = sum(1 for q in H if q[1] <= L)
        i1 = max(i0, sum(1 for q in H if q[1] <= R), sum(1 for q in H if q[0] < R))
        # (the kernel reaches the same i1 by walking x forward from max(#(y <= R), i0))
        out += [(max(x, L), min(y, R), z) for x, y, z in H[i0:i1]]
        hi ^= 1
    return out


def random_tiling(rng, lo_c, hi_c, n):
    cuts = np.sort(rng.integers(lo_c, hi_c + 1, size=n - 1))
    if n > 2 and rng.random() < 0.5:                     # zero-length parts, also at both ends
        cuts[rng.integers(0, n - 1)] = cuts[rng.integers(0, n - 1)]
        cuts = np.sort(cuts)
    if n > 1 and rng.random() < 0.2:
        cuts[0] = lo_c
    if n > 1 and rng.random() < 0.2:
        cuts[-1] = hi_c
    b = [lo_c] + [int(v) for v in cuts] + [hi_c]
    return [(b[i], b[i + 1], int(rng.integers(0, 1000))) for i in range(n)]


def test_index_ranges_equal_the_reference_loop_on_sorted_tilings():
    rng = np.random.default_rng(20261018)
    lo_c, hi_c = 1000, 1400
    n_cases = n_reversed = 0
    for _ in range(6000):
        H0, H1 = random_tiling(rng, lo_c, hi_c, int(rng.integers(1, 9))), random_tiling(rng, lo_c, hi_c, int(rng.integers(1, 9)))
        k = int(rng.integers(0, 7))
        pool = [q[0] for q in H0 + H1] + [q[1] for q in H0 + H1] + list(range(lo_c - 3, hi_c + 40, 7))   # part boundaries, outside the range
        xo = sorted(int(v) for v in rng.choice(pool, size=k))
        if k and rng.random() < 0.3:
            xo[int(rng.integers(0, k))] = xo[0]            # duplicates
            xo.sort()
        hi = int(rng.integers(0, 2))
        a, b = recombine_verbatim(H0, H1, xo, lo_c, hi_c, hi), recombine_by_index_ranges(H0, H1, xo, lo_c, hi_c, hi)
        assert a == b, (H0, H1, xo, hi)
        n_cases += 1
        n_reversed += bool(xo) and (xo[0] < lo_c or xo[-1] > hi_c)
        # the pieces tile again (so the next generation may rely on it): contiguous, from cov_lo to cov_hi
        if a:
            assert all(p[1] == q[0] for p, q in zip(a, a[1:])) and a[0][0] == lo_c and a[-1][1] == hi_c, (H0, H1, xo, hi, a)
    assert n_cases == 6000 and n_reversed > 300
