"""Timing aid: runs ONE rank's share of a locus-range-sharded job on a single GPU (the all-reduce hook is a no-op, so values are wrong
but the kernel work, launches and host read-backs are those of that rank).
usage: emulate_rank.py WORLD STEPS [FLAGS [RANK [WORKLOAD]]]   (FLAGS e.g. 1 = GE_FLAG_SERIAL: the copy alone, after the control chain)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geneevolve_b200 import capi, workloads, dist as gdist

world, steps = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
which = int(sys.argv[4]) if len(sys.argv) > 4 else world // 2          # which rank's share
name = sys.argv[5] if len(sys.argv) > 5 else "config3_100k_x_1M"
cfg = workloads.make_workload(name)
mine = gdist.assign_locus_ranges(cfg["n_loci"], world)[which]
pops = cfg.get("pops", [cfg["n"]])
n_phen = cfg.get("n_phen", 1)
cap = int(max(max(pops), cfg["founders"]) * (1.03 if len(pops) == 1 else 1.10)) + 1024
eng = capi.Engine(n_pop=len(pops), n_chr=len(mine), n_phen=n_phen, representation=capi.GE_REP_BITS, rng_mode=capi.GE_RNG_PHILOX, seed=12345,
                  capacity=cap, rank=0, world_size=world, flags=flags)
if len(pops) > 1:
    workloads.configure_engine_multipop(eng, cfg, pieces=mine)
else:
    workloads.configure_engine(eng, cfg, pieces=mine)
eng.set_allreduce(lambda ptr, count, stream: None)
eng.init_generation0()
gp = [capi.gen_params(q, cfg["mat_cor"], "p", "logit", 0.0, 1.0) for q in pops]
mig = cfg.get("migration")
for g in range(1, 6):
    eng.step_generation(g, gp, mig)
eng.set_profiling(1); eng.reset_kernel_times(); eng.synchronize()
t0 = time.perf_counter(); eng.timer_start()
for g in range(6, 6 + steps):
    eng.step_generation(g, gp, mig)
ms = eng.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
k_ms, k_n, _ = eng.kernel_time(capi.GE_KERNEL_PROPAGATE_BITS)
launches = eng.launch_count()
eng.set_profiling(2); eng.reset_kernel_times()
for g in range(6 + steps, 11 + steps):
    eng.step_generation(g, gp, mig)
phases = {name: round(eng.kernel_time(pid)[0] / 5, 3) for name, pid in capi.GE_PHASES.items()}
morgans = sum(float(cfg["maps"][c][2].sum()) for c, _, _ in mine)   # a chromosome a rank holds any part of is sampled whole
print(f"{name} world {world} rank {which}: {len(mine)} pieces, {sum(s1 - s0 for _, s0, s1 in mine)} loci, {morgans:.3f} Morgans: {ms / steps:.3f} ms/step "
      f"(host wall {wall / steps:.3f}), propagate {k_ms / max(k_n, 1):.3f} ms x {k_n / steps:.0f}, launches/step {launches / steps:.0f}, control chain {phases}")
