#!/usr/bin/env python
"""Runs the REAL reference (oracle/_ref/ge_ref_export, i.e. the reference's own objects driven in
sim_next_generation order) on the statistical scenarios of tests/stats_util.py for N_REPLICATES seeds each and
commits
  tests/golden/stats_inputs_<scenario>.npz   the inputs exactly as the reference parsed them (in.* arrays), and
  tests/golden/stats_reference.json          per-replicate statistics (tests/stats_util.Trajectory.vector()).
Only runs in the build container (needs /root/reference to have built oracle/_ref); the outputs travel.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.gex import read_gex  # noqa: E402
import stats_util as su  # noqa: E402

EXPORT = os.path.join(ROOT, "oracle", "_ref", "ge_ref_export")


def trajectory_from_export(sc, z):
    T = su.Trajectory(sc)
    n_chr, G = len(sc["chrs"]), len(sc["gens"])
    for gen in range(G + 1):
        if gen:
            T.add_couples(z[f"g{gen - 1}.p0.mv"], z[f"g{gen}.p0.couple_male"], z[f"g{gen}.p0.couple_female"], z[f"g{gen}.p0.couple_noff"])
        haps = [z[f"g{gen}.p0.c{c}.hap"] for c in range(n_chr)]
        T.add_generation(gen, haps, z[f"g{gen}.p0.A"][0], z[f"g{gen}.p0.P"][0])
    return T.vector()


def main():
    ref = {}
    for name, sc in su.SCENARIOS.items():
        runs = []
        with tempfile.TemporaryDirectory() as d:
            args = su.write_reference_inputs(sc, d)
            for r, seed in enumerate(su.SEEDS):
                gex = os.path.join(d, f"o{r}.gex")
                subprocess.run([EXPORT, "--export", gex, "--export_hap"] + args + ["--seed", str(seed), "--prefix", os.path.join(d, f"o{r}")],
                               check=True, stdout=subprocess.DEVNULL)
                z = read_gex(gex)
                if r == 0:
                    np.savez_compressed(os.path.join(HERE, f"stats_inputs_{name}.npz"), **{k: v for k, v in z.items() if k.startswith("in.")})
                runs.append({k: v.tolist() for k, v in trajectory_from_export(sc, z).items()})
                os.remove(gex)
        ref[name] = runs
        m = {k: np.mean([r[k] for r in runs], axis=0) for k in runs[0]}
        print(name, {k: np.round(v[-1] if v.ndim else v, 4) for k, v in m.items()})
    with open(os.path.join(HERE, "stats_reference.json"), "w") as f:
        json.dump({"seeds": su.SEEDS, "scenarios": ref}, f)
    print("wrote stats_reference.json", os.path.getsize(os.path.join(HERE, "stats_reference.json")) // 1024, "KiB")


if __name__ == "__main__":
    main()
