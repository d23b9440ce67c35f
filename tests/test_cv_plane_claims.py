"""CPU checks (-m "not gpu") of the arithmetic two control-chain kernels rest on (geneevolve_b200/csrc/ge_kernels.cuh); the GPU
tests compare the kernels themselves with the oracle bit for bit, these pin the rules in numpy.

cv_propagate_rows_kernel: every crossover of a gamete sets ONE toggle bit — at idx = lower_bound(block positions, crossover) — and
the copy mask of a CV word is start ^ (parity of the toggle bits before it in its block) ^ prefix-XOR inside the word; that must be
the reference's rule "CV k comes from haplotype start ^ (#{crossovers <= position k} & 1)" (recombine :2903-2958 read at the CVs).

genetic_value_groups_kernel: the per-CV table {LA[k][t], LD[k][t]} summed per group of four CVs and indexed by the two allele
nibbles gives the same A and D as the per-CV sum (ras_compute_AD :2686-2746) up to rounding."""
import numpy as np


def prefix_xor32(m):
    m = np.uint32(m)
    for s in (1, 2, 4, 8, 16):
        m ^= np.uint32((int(m) << s) & 0xFFFFFFFF)
    return m


def masks_by_toggle_bits(cv_bp, start, crossovers):
    nk = len(cv_bp)
    n_words = (nk + 31) // 32
    toggle = np.zeros(n_words, dtype=np.uint32)
    for x in crossovers:
        idx = int(np.searchsorted(cv_bp, x, side="left"))     # CVs at or above the crossover flip
        if idx < nk:
            toggle[idx >> 5] ^= np.uint32(1 << (idx & 31))
    out = np.zeros(n_words, dtype=np.uint32)
    for w in range(n_words):
        m = prefix_xor32(toggle[w])
        par = start & 1
        for w2 in range(w):
            par ^= bin(int(toggle[w2])).count("1") & 1
        out[w] = ~m & np.uint32(0xFFFFFFFF) if par else m
    return out


def test_toggle_bits_and_prefix_xor_give_the_crossover_parity_masks():
    rng = np.random.default_rng(3)
    for nk, n_xo in [(1, 0), (1, 3), (31, 2), (32, 5), (33, 1), (70, 9), (100, 40), (64, 64)]:
        cv_bp = np.sort(rng.choice(np.arange(1000, 5000), size=nk, replace=False))
        for start in (0, 1):
            xo = np.sort(rng.integers(900, 5100, n_xo))
            if n_xo > 2:
                xo[1] = xo[0]                                  # a duplicated position flips twice
                xo[2] = cv_bp[min(nk - 1, 5)]                  # a crossover exactly on a CV flips that CV
            got = masks_by_toggle_bits(cv_bp, start, xo)
            for k in range(nk):
                want = (start ^ int((xo <= cv_bp[k]).sum())) & 1
                assert (int(got[k >> 5]) >> (k & 31)) & 1 == want, (nk, n_xo, start, k)


def test_group_tables_of_four_cvs_sum_to_the_per_cv_terms():
    rng = np.random.default_rng(4)
    for nk in (1, 3, 4, 5, 31, 32, 45):
        LA, LD = rng.normal(size=(nk, 3)), rng.normal(size=(nk, 3))
        n_groups = ((nk + 31) // 32) * 8
        LG = np.zeros((n_groups, 256, 2))
        for g in range(n_groups):
            for idx in range(256):
                for j in range(4):
                    k = g * 4 + j
                    if k < nk:
                        t = ((idx >> j) & 1) + ((idx >> (4 + j)) & 1)
                        LG[g, idx] += (LA[k, t], LD[k, t])
        for _ in range(20):
            a0, a1 = rng.integers(0, 2, nk), rng.integers(0, 2, nk)
            want = np.array([LA[np.arange(nk), a0 + a1].sum(), LD[np.arange(nk), a0 + a1].sum()])
            pad = n_groups * 4 - nk                            # absent CVs: bits 0 in both planes, no contribution
            b0, b1 = np.concatenate([a0, np.zeros(pad, int)]), np.concatenate([a1, np.zeros(pad, int)])
            got = np.zeros(2)
            for g in range(n_groups):
                n0 = sum(int(b0[g * 4 + j]) << j for j in range(4))
                n1 = sum(int(b1[g * 4 + j]) << j for j in range(4))
                got += LG[g, n0 | (n1 << 4)]
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
