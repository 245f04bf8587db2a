"""The host program (pangenomesim_b200/bin/pangenomesim-b200) against outputs of the unmodified
reference committed under tests/golden/: every host-side file must be byte-identical for the
same seed (rng-compat mode), and -- on the GPU -- the FASTA/MAF files must have the reference's
layout (record names, order, sizes), with sequences that follow the new engine's own oracle."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
EXE = ROOT / "pangenomesim_b200" / "bin" / "pangenomesim-b200"
RUNS = ["config1_seed42", "acc_seed1", "acc_seed2", "acc_seed3", "acc_seed4", "acc_seed5", "safeguard_like"]
HOST_FILES = ["coalescent.newick", "genefrequency.gfs", "distance.mat", "panmatrix.tsv", "reproducible.seed"]


def param_string(run):
    d = json.loads((GOLD / "ref_runs" / f"{run}.json").read_text())
    return ",".join(f"{k}={v}" for k, v in d.items()), d


def run_cli(out, params, *extra, check=True):
    return subprocess.run([str(EXE), "-o", str(out), "-p", params, *extra], check=check, timeout=300,
                          capture_output=True, text=True)


@pytest.mark.parametrize("run", RUNS)
def test_host_files_bit_exact(tmp_path, run):
    params, _ = param_string(run)
    run_cli(tmp_path / "o", params, "--no-sequences")
    for name in HOST_FILES:
        got = (tmp_path / "o" / name).read_bytes()
        want = (GOLD / "ref_runs" / run / name).read_bytes()
        assert got == want, f"{run}/{name}"


def test_config2_host_files_bit_exact(tmp_path):
    import hashlib
    fx = json.loads((GOLD / "config2.json").read_text())
    params = ",".join(f"{k}={v}" for k, v in fx["params"].items())
    run_cli(tmp_path / "o", params, "--no-sequences")
    for name in HOST_FILES:
        assert hashlib.md5((tmp_path / "o" / name).read_bytes()).hexdigest() == fx["md5"][name], name


def test_fast_mode_keeps_the_coalescent(tmp_path):
    params, _ = param_string("acc_seed3")
    run_cli(tmp_path / "o", params, "--no-sequences", "--rng-fast")
    assert (tmp_path / "o" / "coalescent.newick").read_bytes() == (GOLD / "ref_runs" / "acc_seed3" / "coalescent.newick").read_bytes()
    assert (tmp_path / "o" / "distance.mat").read_bytes() == (GOLD / "ref_runs" / "acc_seed3" / "distance.mat").read_bytes()


def test_cli_error_behaviour(tmp_path):
    r = run_cli(tmp_path / "o", "bogus=1", "--no-sequences", check=False)
    assert r.returncode == 1 and "unkown key bogus." in r.stderr  # the reference's own spelling
    r = run_cli(tmp_path / "o", "mut-rate=1e-3", "--no-sequences", check=False)
    assert r.returncode == 1 and "unkown key -3." in r.stderr   # value charset [a-z_0-9.]
    r = subprocess.run([str(EXE), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--param KEY=VALUE" in r.stdout
    r = run_cli(tmp_path / "deep" / "er" / "dir", "seed=3", "--no-sequences")
    assert (tmp_path / "deep" / "er" / "dir" / "coalescent.newick").exists()
    r = run_cli(tmp_path / "o", "num-genomes=0", "--no-sequences", check=False)
    assert r.returncode == 1 and "num-genomes must be at least 1" in r.stderr  # (the reference dies of an uncaught exception)
    r = run_cli(tmp_path / "o", "seed", "--no-sequences", check=False)
    assert r.returncode == 1 and "invalid value" in r.stderr


def test_more_than_999_genomes_host_side(tmp_path):
    """the reference aborts with std::length_error at genome1000 (util.cxx:21-22)"""
    run_cli(tmp_path / "o", "num-genomes=1003,core-size=1,gene-length=10,seed=5", "--no-sequences", "--rng-fast")
    lines = (tmp_path / "o" / "distance.mat").read_text().splitlines()
    assert lines[0] == "1003" and lines[1000].startswith("genome1000  ") and len(lines) == 1004
    assert "genome1003" in (tmp_path / "o" / "coalescent.newick").read_text()


def fasta_records(path):
    lines = Path(path).read_text().split("\n")
    assert lines[-1] == ""
    return [(lines[i][1:], lines[i + 1]) for i in range(0, len(lines) - 1, 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("run", RUNS)
def test_sequence_files_have_the_reference_layout(tmp_path, run, oracle):
    params, d = param_string(run)
    out = tmp_path / "o"
    run_cli(out, params)
    ref_dir = GOLD / "ref_runs" / run
    names = sorted(p.name for p in ref_dir.iterdir())
    assert sorted(p.name for p in out.iterdir()) == names
    L = d.get("gene-length", 100)
    seqs = {}
    for name in names:
        if not name.endswith(".fasta"):
            continue
        got, want = fasta_records(out / name), fasta_records(ref_dir / name)
        assert [n for n, _ in got] == [n for n, _ in want], name           # same records, same order
        assert (out / name).stat().st_size == (ref_dir / name).stat().st_size
        for n, s in got:
            assert len(s) == L and set(s) <= set("ACGT")
            assert seqs.setdefault(n, s) == s                                 # one sequence per (genome, gene)
    # alignment.maf: identical to the reference in everything but the nucleotides
    got = (out / "alignment.maf").read_text().split("\n")
    want = (ref_dir / "alignment.maf").read_text().split("\n")
    assert len(got) == len(want)
    for a, b in zip(got, want):
        if a.startswith("s "):
            fa, fb = a.split(" "), b.split(" ")
            assert fa[:6] == fb[:6], (a[:60], b[:60])
            assert fa[6] == seqs[fa[1]]                                       # MAF and FASTA agree
        else:
            assert a == b
    # ref.fasta = core.fasta + accessory.fasta
    assert (out / "ref.fasta").read_bytes() == (out / "core.fasta").read_bytes() + (out / "accessory.fasta").read_bytes()
    # and the sequences are the engine's: spot-check core gene 0 of genome 1 against the oracle
    core = d.get("core-size", 4)
    if core > 0:
        e = oracle.RefEngine(d["seed"])
        parent, left, right, time = e.coalescent(d.get("num-genomes", 3))
        rate = time * float(np.float32(d.get("mut-rate", 0.01)))
        row = oracle.row(d["seed"], parent, rate, 0, len(parent) - 1, 0, L)
        assert seqs["genome001.0"] == oracle.unpack_ascii(row, L).decode()


def test_host_library_coalescent_matches_reference_dump():
    from pangenomesim_b200.capi import host_coalescent
    trees = json.loads((GOLD / "trees.json").read_text())
    for key, t in trees.items():
        parent, time, leaf = host_coalescent(t["n"], t["seed"], return_time=True)
        assert [int(x) for x in parent] == t["parent"], key
        assert [float(x) for x in time] == [float.fromhex(h) for h in t["time_hex"]], key
        assert list(leaf[:t["n"]]) == list(range(t["n"])) and (leaf[t["n"]:] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("run,devices", [("acc_seed3", 2), ("safeguard_like", 3), ("config1_seed42", 8)])
def test_sharded_cli_is_byte_identical(tmp_path, run, devices):
    """--devices N shards the gene table over N engines (wrapping around the available GPUs);
    every output file must equal the single-engine run byte for byte"""
    params, _ = param_string(run)
    run_cli(tmp_path / "one", params, "--gpu-format")
    run_cli(tmp_path / "many", params, "--devices", str(devices))
    run_cli(tmp_path / "many_gpu", params, "--devices", str(devices), "--gpu-format")
    run_cli(tmp_path / "one_host", params)
    names = sorted(p.name for p in (tmp_path / "one").iterdir())
    for other in ("many", "many_gpu", "one_host"):  # host-expanded (default) and GPU-formatted text, one and several engines
        assert sorted(p.name for p in (tmp_path / other).iterdir()) == names
        for name in names:
            assert (tmp_path / other / name).read_bytes() == (tmp_path / "one" / name).read_bytes(), (other, name)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [1, 2, 3])
def test_cli_check_distances(tmp_path, devices):
    """--check-distances: K2 on every engine, partial counts summed (NCCL where every engine has its own
    GPU, on the host where engines share one), observed distances within the 99 % band of distance.mat"""
    import torch
    params = "num-genomes=40,core-size=300,gene-length=1000,mut-rate=0.05,seed=5"
    r = run_cli(tmp_path / "o", params, "--check-distances", "--devices", str(devices))
    assert "check-distances: 300000 sites, 780 pairs" in r.stderr and " 0 pairs outside" in r.stderr, r.stderr
    exp = (tmp_path / "o" / "distance.mat").read_text().split("\n")
    obs = (tmp_path / "o" / "distance.observed.mat").read_text().split("\n")
    assert exp[0] == obs[0] == "40" and len(exp) == len(obs)
    e = np.array([[float(x) for x in line.split()[1:]] for line in exp[1:41]])
    o = np.array([[float(x) for x in line.split()[1:]] for line in obs[1:41]])
    assert np.allclose(o, o.T) and (np.diag(o) == 0).all()
    assert np.abs(o - e).max() < 0.01 and np.abs(o - e).max() > 0  # sampling noise, not a copy


REF_EXE = ROOT / "oracle" / "_ref" / "pangenomesim"


@pytest.mark.skipif(not REF_EXE.exists(), reason="oracle/_ref/pangenomesim not built (make -C oracle)")
@pytest.mark.parametrize("params", [
    "num-genomes=520,core-size=2,gene-length=20,theta=5,rho=0.5,seed=11",       # n >= 512: threaded distance.mat
    "num-genomes=97,core-size=3,gene-length=33,theta=40,rho=2,mut-rate=0.2,seed=12",
    "num-genomes=2,core-size=1,gene-length=5,seed=13",
])
def test_host_files_against_the_reference_binary_live(tmp_path, params):
    """parameter sets outside tests/golden: the unmodified reference (compiled by oracle/Makefile)
    and the host program run side by side; every host-side file must be byte-identical"""
    subprocess.run([str(REF_EXE), "-o", str(tmp_path / "ref"), "-p", params], check=True, timeout=600,
                   capture_output=True)
    run_cli(tmp_path / "o", params, "--no-sequences", "--rng-compat")
    for name in HOST_FILES:
        assert (tmp_path / "o" / name).read_bytes() == (tmp_path / "ref" / name).read_bytes(), name


@pytest.mark.gpu
def test_cli_packed_side_output(tmp_path):
    """--packed: genomes.pgs2 beside the usual files; unpacked it equals every genomeNNN.fasta"""
    from pangenomesim_b200.packed import PackedGenomes
    out = tmp_path / "o"
    run_cli(out, "num-genomes=40,core-size=6,gene-length=130,theta=8,rho=0.5,seed=21", "--packed")
    pk = PackedGenomes(out / "genomes.pgs2")
    assert pk.n_genomes == 40 and pk.gene_length == 130 and pk.n_dense >= 6
    for g in range(40):
        assert pk.fasta(g) == (out / ("genome%03d.fasta" % (g + 1))).read_bytes(), g
    r = run_cli(tmp_path / "p", "num-genomes=5,seed=3", "--packed", "--devices", "2", check=False)
    assert r.returncode == 1 and "--packed needs a single engine" in r.stderr
