"""The C++ host (host/): the reference's command line and file formats over the C-ABI.

The same checks run twice: without a GPU on the host sources compiled against the CPU oracle (tests/host_oracle_shim.h,
built into tests/_build/ by this test) and with -m gpu on the product binary host/geneevolve_b200_cli, which links
libgeneevolve_b200.so.  Where the reference binary (oracle/_ref/GeneEvolve_ref) is present, it is run on the same
input files: the set of output files, every header and every draw-independent column (generation-0 IDs, additive
values, the generation-0 var_A/var_E of the summary) must be identical to what the reference writes.  The genotype
outputs are checked against each other: every part of the `.int` file, materialised from the founder panel
(ras_convert_interval_to_hap_matrix, src/Simulation.cpp:1186-1230), must reproduce the `.hap` file.
"""
import os
import subprocess

import numpy as np
import pytest

import stats_util as su

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "GeneEvolve_ref")
INFO_HEADER = ("ID ID_Father ID_Mother ID_Fathers_Father ID_Fathers_Mother ID_Mothers_Father ID_Mothers_Mother sex "
               "ph1_A ph1_D ph1_G ph1_C ph1_E ph1_F ph1_P MV SV SV_f")
SUMMARY_HEADER = ("gen ph1_var_A ph1_var_D ph1_var_G ph1_var_C ph1_var_E ph1_var_F ph1_var_P ph1_h2 ph1_var_G_std "
                  "var_mating_value var_selection_value")
INT_HEADER = "h_ID chr hap st en hap_index gen0_indv root_pop"


def oracle_cli():
    """host/*.cpp compiled against the CPU oracle (test-only build)."""
    from oracle import oracle
    oracle.lib()
    out = os.path.join(HERE, "_build", "host_oracle_cli")
    srcs = [os.path.join(ROOT, "host", f) for f in ("main.cpp", "ge_host.cpp")] + [os.path.join(HERE, "host_thread_collective.cpp")]
    deps = srcs + [os.path.join(ROOT, "host", "ge_host.hpp"), os.path.join(HERE, "host_oracle_shim.h"), oracle.LIB]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++14", "-I" + os.path.join(ROOT, "include"), "-include", os.path.join(HERE, "host_oracle_shim.h"),
                        "-o", out] + srcs + ["-L" + os.path.dirname(oracle.LIB), "-l:libge_oracle.so", "-pthread", "-Wl,-rpath," + os.path.dirname(oracle.LIB)], check=True)
    return out


def product_cli():
    out = os.path.join(ROOT, "host", "geneevolve_b200_cli")
    assert os.path.exists(out), "host/geneevolve_b200_cli is missing: run __graft_entry__.build()"
    return out


def scenario(tmp_path, n_gen=3):
    sc = dict(su.SCENARIOS["S_assort"])
    sc["gens"] = sc["gens"][:n_gen]
    return sc, su.write_reference_inputs(sc, str(tmp_path))


def check_cli(cli, tmp_path):
    sc, args = scenario(tmp_path)
    G = len(sc["gens"])
    pre = str(tmp_path / "ours")
    r = subprocess.run([cli] + args + ["--seed", "5", "--prefix", pre, "--out_hap", "--out_interval", "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    names = sorted(f[len("ours"):] for f in os.listdir(tmp_path) if f.startswith("ours"))
    expect = sorted([f".info.pop1.gen{g}.txt" for g in range(G + 1)] + [".pop1.summary"] +
                    [f".pop1.gen{G}.chr{c}.{e}" for c in sc["chrs"] for e in ("hap", "indv", "int")])
    assert names == expect
    # headers and shapes
    for g in range(G + 1):
        with open(f"{pre}.info.pop1.gen{g}.txt") as f:
            assert f.readline().rstrip("\n") == INFO_HEADER
    with open(f"{pre}.pop1.summary") as f:
        assert f.readline().rstrip("\n") == SUMMARY_HEADER
    summ = np.loadtxt(f"{pre}.pop1.summary", skiprows=1)
    assert summ.shape == (G + 1, 12) and np.array_equal(summ[:, 0], np.arange(G + 1))
    assert summ[0, 1] == pytest.approx(sc["va"], rel=1e-5) and np.allclose(summ[:, 5], sc["ve"], rtol=1e-5)   # var_A(0) = va, var_E = ve every generation
    info_last = np.loadtxt(f"{pre}.info.pop1.gen{G}.txt", skiprows=1)
    n_last = info_last.shape[0]
    assert abs(n_last - sc["gens"][-1][0]) < 6 * np.sqrt(sc["gens"][-1][0])     # Poisson families
    assert np.array_equal(info_last[:, 0], np.arange(1, n_last + 1)) and set(info_last[:, 7]) <= {1.0, 2.0}
    assert np.allclose(info_last[:, 10], info_last[:, 8] + info_last[:, 9], atol=1e-5)                         # G = A + D
    assert np.allclose(info_last[:, 14], info_last[:, 8:14].sum(axis=1) - info_last[:, 10], atol=2e-5)         # P = A + D + C + E + F
    # genotype outputs: .hap (rows = SNPs, columns = haplotypes), .indv, .int; segments materialise to the .hap file
    panel = su.make_panel(sc)
    for P in panel:
        c = P["chr"]
        base = f"{pre}.pop1.gen{G}.chr{c}"
        hap = np.loadtxt(base + ".hap", dtype=np.uint8)
        assert hap.shape == (sc["n_snp"], 2 * n_last)
        assert np.array_equal(np.loadtxt(base + ".indv", dtype=np.int64), np.arange(1, n_last + 1))
        with open(base + ".int") as f:
            assert f.readline().rstrip("\n") == INT_HEADER
            rows = [line.split() for line in f]
        lo, hi = 1000 * c, 1000 * c + (sc["map_rows"] - 1) * sc["map_step"]
        rebuilt = np.zeros_like(hap)
        cover = {}
        for h_id, chr_, ih, st, en, hidx, who, root in rows:
            assert int(chr_) == c and root == "1" and who == f"id{(int(hidx) - 1) // 2 + 1}.{(int(hidx) - 1) % 2 + 1}"
            col = 2 * (int(h_id) - 1) + int(ih)
            sel = (P["pos"] >= int(st)) & (P["pos"] < int(en))
            rebuilt[sel, col] = P["hap"][sel, int(hidx) - 1]
            cover.setdefault(col, []).append((int(st), int(en)))
        assert np.array_equal(rebuilt, hap), "the .int parts do not materialise to the .hap file"
        for col in range(2 * n_last):   # parts tile [first map row, last map row) in order, no gaps
            seg = cover[col]
            assert seg[0][0] == lo and seg[-1][1] == hi and all(a[1] == b[0] for a, b in zip(seg, seg[1:]))
    # against the reference binary on the same files
    if os.path.exists(REF_BIN):
        rp = str(tmp_path / "ref")
        rr = subprocess.run([REF_BIN] + args + ["--seed", "5", "--prefix", rp, "--out_hap", "--out_interval"], capture_output=True, text=True)
        assert rr.returncode == 0
        ref_names = sorted(f[len("ref"):] for f in os.listdir(tmp_path) if f.startswith("ref"))
        assert ref_names == names
        for suffix in [".info.pop1.gen0.txt", ".pop1.summary", f".pop1.gen{G}.chr{sc['chrs'][0]}.int"]:
            with open(pre + suffix) as a, open(rp + suffix) as b:
                assert a.readline() == b.readline(), suffix
        a, b = np.loadtxt(pre + ".info.pop1.gen0.txt", skiprows=1), np.loadtxt(rp + ".info.pop1.gen0.txt", skiprows=1)
        assert a.shape == b.shape and np.array_equal(a[:, :7], b[:, :7])
        for col in (8, 9, 10, 11, 13):   # A, D, G, C, F at generation 0 do not depend on any draw
            assert np.array_equal(a[:, col], b[:, col]), col
        rs = np.loadtxt(rp + ".pop1.summary", skiprows=1)
        assert rs.shape == summ.shape and np.array_equal(rs[0, [1, 2, 3, 4, 5, 6]], summ[0, [1, 2, 3, 4, 5, 6]])
        with open(pre + f".pop1.gen{G}.chr{sc['chrs'][0]}.hap") as f1, open(rp + f".pop1.gen{G}.chr{sc['chrs'][0]}.hap") as f2:
            l1, l2 = f1.readline(), f2.readline()
            assert set(l1) == set(l2) == set("01 \n") and l1[1] == l2[1] == " " and l1.endswith(" \n") and l2.endswith(" \n")


def check_errors(cli, tmp_path):
    sc, args = scenario(tmp_path, 1)
    bad = [a if a != str(tmp_path / "s.rmap") else str(tmp_path / "missing.rmap") for a in args]
    r = subprocess.run([cli] + bad + ["--prefix", str(tmp_path / "e")], capture_output=True, text=True)
    assert r.returncode == 255 and "Error: can not open the file [" in r.stdout          # exit code -1 like src/Main.cpp:84-88
    r = subprocess.run([cli] + args[2:] + ["--prefix", str(tmp_path / "e")], capture_output=True, text=True)
    assert r.returncode == 255 and "missing parameter [--file_gen_info]" in r.stdout
    r = subprocess.run([cli] + args + ["--out_vcf"], capture_output=True, text=True)
    assert r.returncode == 255 and "--out_vcf" in r.stdout and "can't convert to VCF output format" in r.stdout   # src/Simulation.cpp:1071-1075


def check_plink(cli, tmp_path):
    """--out_plink / --out_plink01 (ras_write_hap_to_plink_format, src/Simulation.cpp:1254-1303; format_plink::write_ped_map and
    write_ped01_map, src/format_plink.cpp:5-135): the .map file is draw-independent and must equal the reference's byte for
    byte; every .ped row is `father+1 ID+1 father+1 mother+1 sex -9` followed by the two alleles of every SNP, and must spell
    out exactly the haplotypes the same run writes with --out_hap."""
    sc, args = scenario(tmp_path, 2)
    G = len(sc["gens"])
    outs = {}
    for flag in ("--out_plink", "--out_plink01"):
        pre = str(tmp_path / ("o" + flag[6:]))
        r = subprocess.run([cli] + args + ["--seed", "5", "--prefix", pre, "--out_hap", flag, "--quiet"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[flag] = pre
    info = np.loadtxt(outs["--out_plink"] + f".info.pop1.gen{G}.txt", skiprows=1)
    n = info.shape[0]
    for c in sc["chrs"]:
        hap = np.loadtxt(outs["--out_plink"] + f".pop1.gen{G}.chr{c}.hap", dtype=np.uint8)          # [snp][2n]
        for flag, code in (("--out_plink", "AC"), ("--out_plink01", "01")):
            base = outs[flag] + f".pop1.gen{G}.chr{c}"
            rows = [line.rstrip("\n").split(" ") for line in open(base + ".ped")]
            assert len(rows) == n and all(len(r) == 6 + 2 * sc["n_snp"] for r in rows)
            ped = np.array([r[:6] for r in rows])
            assert np.array_equal(ped[:, 1].astype(np.int64), info[:, 0].astype(np.int64))            # IID = ID (already + 1 in .info)
            assert np.array_equal(ped[:, 0], ped[:, 2]) and np.array_equal(ped[:, 2].astype(np.int64), info[:, 1].astype(np.int64))
            assert np.array_equal(ped[:, 3].astype(np.int64), info[:, 2].astype(np.int64))
            assert np.array_equal(ped[:, 4].astype(np.int64), info[:, 7].astype(np.int64)) and set(ped[:, 5]) == {"-9"}
            al = np.array([r[6:] for r in rows]).reshape(n, sc["n_snp"], 2)                            # [ind][snp][hap]
            want = np.array(list(code))[hap.T.reshape(n, 2, sc["n_snp"]).transpose(0, 2, 1)]
            assert np.array_equal(al, want), (flag, c)
            lines = open(base + ".map").read().splitlines()
            assert len(lines) == sc["n_snp"] and lines[0].split(" ")[0] == str(c) and lines[0].split(" ")[2] == "0"
    if os.path.exists(REF_BIN):
        for flag in ("--out_plink", "--out_plink01"):
            rp = str(tmp_path / ("r" + flag[6:]))
            rr = subprocess.run([REF_BIN] + args + ["--seed", "5", "--prefix", rp, flag], capture_output=True, text=True)
            assert rr.returncode == 0
            for c in sc["chrs"]:
                a, b = outs[flag] + f".pop1.gen{G}.chr{c}", rp + f".pop1.gen{G}.chr{c}"
                assert open(a + ".map", "rb").read() == open(b + ".map", "rb").read()
                ra = open(a + ".ped").readline().rstrip("\n").split(" ")
                rb = open(b + ".ped").readline().rstrip("\n").split(" ")
                assert len(ra) == len(rb) and ra[5] == rb[5] == "-9" and set(ra[6:]) <= set(rb[6:]) | set("AC01")
                assert rb[0] == rb[2] and rb[1] == "1" and ra[1] == "1"                                  # same ID conventions


def test_host_cli_plink_on_oracle(tmp_path):
    check_plink(oracle_cli(), tmp_path)


@pytest.mark.gpu
def test_host_cli_plink_on_gpu(tmp_path):
    check_plink(product_cli(), tmp_path)


def check_two_populations(cli, tmp_path):
    """--next_population and --file_migration: under random mating the sizes after migration are draw-independent, so
    the row counts of every .info file must equal the reference's."""
    sc = dict(su.SCENARIOS["S_drift_ld"])
    sc["gens"] = [(120, 0.0, "p", "thr", 1, 1)] * 3
    a1 = su.write_reference_inputs(sc, str(tmp_path), tag="p1")
    sc2 = dict(sc)
    sc2["gens"] = [(80, 0.0, "p", "thr", 1, 1)] * 3
    a2 = su.write_reference_inputs(sc2, str(tmp_path), tag="p2")
    mig = tmp_path / "mig.txt"
    mig.write_text("0.9 0.1 0.25 0.75\n" * 3)
    args = a1 + ["--next_population"] + a2 + ["--file_migration", str(mig)]
    pre = str(tmp_path / "ours")
    r = subprocess.run([cli] + args + ["--seed", "9", "--prefix", pre, "--out_interval", "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = {(p, g): sum(1 for _ in open(f"{pre}.info.pop{p}.gen{g}.txt")) - 1 for p in (1, 2) for g in range(4)}
    n1, n2 = 120, 80
    m12, m21 = round(0.1 * n1), round(0.25 * n2)
    for g in range(1, 4):
        assert rows[(1, g)] == n1 - m12 + m21 and rows[(2, g)] == n2 - m21 + m12
    ints = sorted(f for f in os.listdir(tmp_path) if f.startswith("ours") and f.endswith(".int"))
    assert ints == ["ours.pop1.gen3.chr1.int", "ours.pop2.gen3.chr1.int"]
    roots = {line.split()[-1] for line in list(open(tmp_path / "ours.pop1.gen3.chr1.int"))[1:]}
    assert roots == {"1", "2"}      # after three generations of migration both founder panels contribute
    if os.path.exists(REF_BIN):
        rp = str(tmp_path / "ref")
        rr = subprocess.run([REF_BIN] + args + ["--seed", "9", "--prefix", rp, "--out_interval"], capture_output=True, text=True)
        assert rr.returncode == 0
        for (p, g), n in rows.items():
            assert sum(1 for _ in open(f"{rp}.info.pop{p}.gen{g}.txt")) - 1 == n, (p, g)
        with open(f"{rp}.pop2.summary") as f1, open(f"{pre}.pop2.summary") as f2:
            assert f1.readline() == f2.readline()


def test_host_cli_two_populations_on_oracle(tmp_path):
    check_two_populations(oracle_cli(), tmp_path)


@pytest.mark.gpu
def test_host_cli_two_populations_on_gpu(tmp_path):
    check_two_populations(product_cli(), tmp_path)


def check_multi_gpu(cli, tmp_path):
    """--gpus 2 (chromosome shards, one context per device, sum-allreduce of the partial genetic values) must write
    the same files as --gpus 1: genotype and segment files byte for byte, per-individual columns to 1e-9."""
    sc, args = scenario(tmp_path, 3)
    outs = {}
    for g in (1, 2):
        pre = str(tmp_path / f"g{g}")
        r = subprocess.run([cli] + args + ["--seed", "11", "--prefix", pre, "--out_hap", "--out_interval", "--quiet", "--gpus", str(g)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[g] = sorted(f[len(f"g{g}"):] for f in os.listdir(tmp_path) if f.startswith(f"g{g}."))
    assert outs[1] == outs[2] and len(outs[1]) == 4 + 1 + 3 * len(sc["chrs"])
    for suffix in outs[1]:
        a, b = str(tmp_path / ("g1" + suffix)), str(tmp_path / ("g2" + suffix))
        if suffix.endswith((".hap", ".int", ".indv")):
            assert open(a, "rb").read() == open(b, "rb").read(), suffix
        else:
            with open(a) as fa, open(b) as fb:
                assert fa.readline() == fb.readline()
            np.testing.assert_allclose(np.loadtxt(a, skiprows=1), np.loadtxt(b, skiprows=1), rtol=1e-5, atol=1e-6, err_msg=suffix)
    r = subprocess.run([cli] + args + ["--gpus", "7", "--prefix", str(tmp_path / "e")], capture_output=True, text=True)
    assert r.returncode == 255 and ("exceeds the number of chromosomes" in r.stdout or "cannot set up the collective" in r.stdout)


def test_host_cli_multi_gpu_on_oracle(tmp_path):
    check_multi_gpu(oracle_cli(), tmp_path)


@pytest.mark.gpu
def test_host_cli_multi_gpu_on_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    check_multi_gpu(product_cli(), tmp_path)


def test_host_cli_on_oracle(tmp_path):
    check_cli(oracle_cli(), tmp_path)


def test_host_cli_errors_on_oracle(tmp_path):
    check_errors(oracle_cli(), tmp_path)


@pytest.mark.gpu
def test_host_cli_on_gpu(tmp_path):
    check_cli(product_cli(), tmp_path)


@pytest.mark.gpu
def test_host_cli_errors_on_gpu(tmp_path):
    check_errors(product_cli(), tmp_path)
