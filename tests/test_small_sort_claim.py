"""CPU check (-m "not gpu") of the rule small_sort_merge_kernel (geneevolve_b200/csrc/ge_mating.cuh) uses to turn stably sorted
tiles into one stable sort: final rank = position inside the own tile + #(keys <= key) in every earlier tile + #(keys < key) in
every later tile.  The GPU tests check the kernels themselves (couples bit-exact against the oracle under the Philox streams);
this pins the arithmetic, with heavy ties and with the all-ones keys that the remainder draw gives inbred couples."""
import numpy as np


def tile_sort_then_rank(keys, vals, tile):
    n = len(keys)
    tk, tv = keys.copy(), vals.copy()
    for b in range(0, n, tile):                       # small_sort_tile_kernel: a stable sort of every tile
        o = np.argsort(keys[b:b + tile], kind="stable")
        tk[b:b + tile], tv[b:b + tile] = keys[b:b + tile][o], vals[b:b + tile][o]
    out_k, out_v = np.empty_like(tk), np.empty_like(tv)
    n_tiles = (n + tile - 1) // tile
    for e in range(n):
        mine = e // tile
        rank = e - mine * tile
        for b in range(n_tiles):
            if b == mine:
                continue
            t = tk[b * tile:(b + 1) * tile]
            rank += np.searchsorted(t, tk[e], side="right" if b < mine else "left")
        out_k[rank], out_v[rank] = tk[e], tv[e]
    return out_k, out_v


def test_rank_merge_of_sorted_tiles_is_a_stable_sort():
    rng = np.random.default_rng(11)
    for n, tile, spread in [(1, 8, 4), (7, 8, 3), (8, 8, 2), (9, 8, 2), (100, 16, 5), (257, 32, 1000), (300, 64, 2)]:
        keys = rng.integers(0, spread, n).astype(np.uint64)
        keys[rng.random(n) < 0.1] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vals = np.arange(n, dtype=np.uint32)
        k, v = tile_sort_then_rank(keys, vals, tile)
        o = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[o]) and np.array_equal(v, vals[o]), (n, tile)


def bitonic_tile(keys, vals):
    """small_sort_tile_kernel's network, index for index: compare-exchange on (key, position in the tile), the tile padded to a power
    of two (at least 64) with all-ones keys at later positions, values gathered by position at the end."""
    cnt = len(keys)
    m = 64
    while m < cnt:
        m <<= 1
    sk = np.full(m, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    sk[:cnt] = keys
    sp = np.arange(m)
    k = 2
    while k <= m:
        j = k >> 1
        while j > 0:
            p = np.arange(m >> 1)
            i = ((p & ~(j - 1)) << 1) | (p & (j - 1))
            l = i | j
            a, b, pa, pb = sk[i], sk[l], sp[i], sp[l]
            gt = (a > b) | ((a == b) & (pa > pb))
            swap = gt == ((i & k) == 0)
            sk[i], sk[l] = np.where(swap, b, a), np.where(swap, a, b)
            sp[i], sp[l] = np.where(swap, pb, pa), np.where(swap, pa, pb)
            j >>= 1
        k <<= 1
    return sk[:cnt], vals[sp[:cnt]]


def test_bitonic_network_on_key_and_position_is_a_stable_sort():
    rng = np.random.default_rng(12)
    for n, spread in [(1, 2), (2, 1), (63, 3), (64, 2), (65, 4), (500, 7), (1000, 2), (2048, 50), (2047, 1), (4096, 3), (3000, 2)]:
        keys = rng.integers(0, spread, n).astype(np.uint64)
        keys[rng.random(n) < 0.1] = np.uint64(0xFFFFFFFFFFFFFFFF)   # real all-ones keys must stay ahead of the padding
        vals = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        k, v = bitonic_tile(keys, vals)
        o = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[o]) and np.array_equal(v, vals[o]), n
