"""Drop-in check at the command line: the reference binary and host/geneevolve_b200_cli on the SAME input files
(config-3 shape at a size the reference finishes: N individuals x 22 autosomes, assortative mating + selection),
wall-clock for the whole run including text input and the per-generation .info output."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from geneevolve_b200 import workloads

n, gens = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = workloads.make_workload("config3_100k_x_1M", n_override=n)
cfg["founders"] = n
rng = np.random.default_rng(5)
for cv in cfg["cvs"]:
    cv["val"] = (rng.random((2 * n, len(cv["bp"]))) < 0.5).astype(np.uint8)
with tempfile.TemporaryDirectory() as d:
    args = workloads.write_reference_inputs(cfg, d, gens)
    out = {}
    for name, exe in (("reference", os.path.join(ROOT, "oracle", "_ref", "GeneEvolve_ref")), ("b200_cli", os.path.join(ROOT, "host", "geneevolve_b200_cli"))):
        t0 = time.perf_counter()
        r = subprocess.run([exe] + args + ["--seed", "7", "--prefix", os.path.join(d, name)], capture_output=True, text=True)
        out[name] = time.perf_counter() - t0
        assert r.returncode == 0, r.stdout[-2000:]
        if name == "b200_cli":
            print("".join(l + "\n" for l in r.stdout.splitlines() if "Time taken" in l), end="")
        files = sorted(f for f in os.listdir(d) if f.startswith(name))
        rows = sum(1 for _ in open(os.path.join(d, f"{name}.info.pop1.gen{gens}.txt"))) - 1
        h2 = open(os.path.join(d, f"{name}.pop1.summary")).read().splitlines()[-1].split()[8]
        print(f"{name}: {out[name]:.2f} s wall for {gens} generations of {n} individuals x 22 chr; {len(files)} output files; last generation {rows} individuals, h2 {h2}")
    print(f"speed-up of the whole command line run: {out['reference'] / out['b200_cli']:.1f}x")
