"""Chromosome-sharded CUDA contexts (-m gpu, one GPU): two contexts, each owning part of the chromosomes,
joined by a sum-allreduce hook, must reproduce the single-context run — couples, draws and haplotypes bit for
bit, fp64 columns to 1e-10.  The two "ranks" are threads in this process sharing the GPU (their collective is
a host-side barrier + sum), which exercises exactly the library code the NCCL path uses."""
import threading

import numpy as np
import pytest
import torch

from geneevolve_b200 import capi, dist as gdist
from golden_util import Golden
from test_sharding_gloo import configure_pieces, run_generations, shard_of

pytestmark = pytest.mark.gpu


REPS = {"bits": (capi.GE_REP_BITS, 0), "segments": (capi.GE_REP_SEGMENTS, 0), "segments-bulk-stream": (capi.GE_REP_SEGMENTS, 400_000),
        "bits+segments": (capi.GE_REP_BITS | capi.GE_REP_SEGMENTS, 0)}


@pytest.mark.parametrize("split", ["chromosomes", "locus-tiles"])
@pytest.mark.parametrize("rep", sorted(REPS))
@pytest.mark.parametrize("name", ["B_rm_mut", "A_am_pois", "D_two_pops"])
def test_two_sharded_contexts_match_one(cuda_lib, name, rep, split):
    if split == "locus-tiles" and rep != "bits":
        pytest.skip("founder segments are sharded by whole chromosomes (a part has no loci)")
    G = Golden(name)
    n_gen = min(G.G, 3)
    representation, seg_capacity = REPS[rep]
    segs = bool(representation & capi.GE_REP_SEGMENTS)
    single = capi.Engine(cuda_lib, **G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=representation, seg_capacity=seg_capacity, capacity=G.philox_capacity()))
    G.configure(single)
    ref = run_generations(G, single, n_gen, segments=segs)

    world = 2
    parts = shard_of(G, split, world)
    barrier = threading.Barrier(world)
    slots = [None] * world
    results, errors = [None] * world, []

    def make_hook(rank):
        def hook(ptr, count, stream):
            t = torch.as_tensor(gdist._DevPtr(ptr, count), device="cuda:0")
            torch.cuda.ExternalStream(stream, device=0).synchronize()
            slots[rank] = t.cpu()
            barrier.wait()
            total = slots[0] + slots[1]
            barrier.wait()
            t.copy_(total.cuda())
            torch.cuda.synchronize()
        return hook

    def run(rank):
        try:
            kw = G.engine_kwargs(rng_mode=capi.GE_RNG_PHILOX, representation=representation, seg_capacity=seg_capacity, capacity=G.philox_capacity())
            kw.update(n_chr=len(parts[rank]), rank=rank, world_size=world)
            eng = capi.Engine(cuda_lib, **kw)
            configure_pieces(G, eng, parts[rank])
            eng.set_allreduce(make_hook(rank))
            results[rank] = run_generations(G, eng, n_gen, segments=segs)
        except Exception as e:  # pragma: no cover
            errors.append(e)
            barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errors, errors
    for rank in range(world):
        out = results[rank]
        for k in ("pos_male", "pos_female", "inbreed", "num_offspring"):
            assert np.array_equal(out["couples"][k], ref["couples"][k])
        assert np.array_equal(out["ind"]["ids"], ref["ind"]["ids"]) and np.array_equal(out["ind"]["sex"], ref["ind"]["sex"])
        for k in "ADGCEFP":
            np.testing.assert_allclose(out["ind"][k], ref["ind"][k], rtol=1e-10, atol=1e-12)
        for k, (c, s0, s1) in enumerate(parts[rank]):
            assert np.array_equal(out["hap"][k], ref["hap"][c][:, s0:s1])
            if segs:
                for key in ("seg_off", "seg", "mut_off", "mut_bp"):
                    assert np.array_equal(out["seg"][k][key], ref["seg"][c][key]), f"rank {rank} chromosome {c}: {key}"
