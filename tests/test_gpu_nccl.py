"""Two ranks over NCCL (-m gpu, needs two GPUs: gpurun --gpus 2): bench.py under torch.distributed.run on a reduced config 3,
against the same workload on one GPU.  The ranks hold half of every row each; the library issues the one exchange of the path
itself (ge_set_allreduce_nccl: ncclAllReduce on the control stream), so the sharded generations must (a) replay as captured
CUDA graphs with the collective inside and (b) leave the same pedigree, sexes and couples as the single-GPU run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--steps", "4", "--warmup", "3", "--individuals", "6000", "--loci", "200000", "--no-cpu-baseline", "--no-other-workloads"]


def last_json(out):
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def test_two_ranks_replay_graphs_and_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    one = subprocess.run([sys.executable, "bench.py", "--gpus", "1"] + COMMON, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout + one.stderr
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                          "bench.py", "--gpus", "2"] + COMMON, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stdout + two.stderr
    a, b = last_json(one.stdout), last_json(two.stdout)
    assert b["n_gpus"] == 2 and b["e2e"]["state_hash_equal_on_all_ranks"]
    assert a["e2e"]["generations_simulated"] == b["e2e"]["generations_simulated"]
    assert a["e2e"]["state_hash"] == b["e2e"]["state_hash"]
    assert b["config"]["graph_replays_rank0"] > 0, "the sharded control chain did not replay as a CUDA graph"
