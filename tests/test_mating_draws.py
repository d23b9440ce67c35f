"""The mating draws the reference consumed (tests/golden `g*.p*.mate.*`, exported by oracle/ref_driver.cpp's export_mating_draws) are
what they claim to be: a plain numpy restatement of random_mate (src/Simulation.cpp:2090-2157) and assort_mate (:2167-2360) fed with
them — thinning uniforms, the --MM uniforms, the shuffled trim order, the template normals, the Poisson family sizes or the shuffled
remainder order — gives back the reference's own `_couples_info`, couple for couple.  CPU only; this is the specification of
ge_mate_replay (the CUDA mating kernels under GE_RNG_REPLAY, checked against the same couples in tests/test_gpu_replay.py)."""
import numpy as np
import pytest

from golden_util import SCENARIOS, Golden


def ras_rank(x):
    """CommFunc::ras_rank (src/CommFunc.cpp:152-161): position in a stable ascending sort."""
    r = np.empty(len(x), np.int64)
    r[np.argsort(x, kind="stable")] = np.arange(len(x))
    return r


def couples_from_draws(G, gen, p):
    z = lambda k: G.g(gen, p, "mate." + k)  # noqa: E731
    par = lambda k: G.g(gen - 1, p, k)      # noqa: E731  (the parents are generation gen-1 as exported, i.e. after its migration)
    sex, svf, mv, ids = par("sex"), par("svf"), par("mv"), par("ids")
    pre = f"in.p{p}."
    keep = z("thin_u") < svf
    pop_size = int(G.z[pre + "pop_size"][gen - 1])
    if int(G.z[pre + "RM"]):
        males, females = np.flatnonzero(keep & (sex == 1)), np.flatnonzero(keep & (sex == 2))
        return dict(male=males[z("rm_father_idx")], female=females[z("rm_mother_idx")], inbreed=np.zeros(pop_size, np.uint8), noff=np.ones(pop_size, np.int64))
    mm = float(G.z[pre + "MM_percent"])
    lists = {1: [], 2: []}
    for i in np.flatnonzero(keep):
        if sex[i] in (1, 2):
            lists[int(sex[i])] += [i, i] if z("mm_u")[i] < mm else [i]
    lm, lf = np.array(lists[1], np.int64), np.array(lists[2], np.int64)
    n2 = min(len(lm), len(lf))
    if len(lm) != len(lf):   # std::random_shuffle + erase of the first n_remove (:2233-2246)
        order = z("trim_order").astype(np.int64)
        if len(lm) > len(lf):
            lm = lm[order][len(lm) - n2:]
        else:
            lf = lf[order][len(lf) - n2:]
    for lst in (lm, lf):     # std::sort is not stable: the restatement is only defined when distinct individuals never tie
        v = mv[lst]
        s = np.argsort(v, kind="stable")
        ties = v[s][1:] == v[s][:-1]
        assert np.all(lst[s][1:][ties] == lst[s][:-1][ties]), "distinct individuals with equal mating values"
    lm, lf = lm[np.argsort(mv[lm], kind="stable")], lf[np.argsort(mv[lf], kind="stable")]
    male, female = lm[ras_rank(z("t1"))], lf[ras_rank(z("t2"))]
    inbreed = np.zeros(n2, np.uint8)
    if int(G.z[pre + "avoid_inbreeding"]):
        a, b = ids[male], ids[female]
        sib = a[:, 1] == b[:, 1]
        cousin = (a[:, 3] == b[:, 3]) | (a[:, 3] == b[:, 5]) | (a[:, 5] == b[:, 3]) | (a[:, 5] == b[:, 5]) | \
                 (a[:, 4] == b[:, 4]) | (a[:, 4] == b[:, 6]) | (a[:, 6] == b[:, 4]) | (a[:, 6] == b[:, 6])
        inbreed = (sib | cousin).astype(np.uint8)
    if chr(int(G.z[pre + "offspring_dist"][gen - 1])) in "pP":
        noff = z("family").astype(np.int64)
    else:
        n_ok = n2 - int(inbreed.sum())
        nf = pop_size // n_ok
        noff = np.full(n2, nf, np.int64)
        noff[z("remainder_order").astype(np.int64)[:pop_size - nf * n_ok]] += 1
    return dict(male=male, female=female, inbreed=inbreed, noff=noff)


@pytest.mark.parametrize("name", SCENARIOS)
def test_exported_mating_draws_reproduce_the_reference_couples(name):
    G = Golden(name)
    for gen in range(1, G.G + 1):
        for p in range(G.n_pop):
            c = couples_from_draws(G, gen, p)
            for k, key in (("male", "couple_male"), ("female", "couple_female"), ("inbreed", "couple_inbreed"), ("noff", "couple_noff")):
                assert np.array_equal(c[k], G.g(gen, p, key).astype(c[k].dtype)), f"{name} gen {gen} pop {p}: {key}"
