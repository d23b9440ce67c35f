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
