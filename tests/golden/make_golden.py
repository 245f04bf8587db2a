#!/usr/bin/env python3
"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref/pangenomesim and
oracle/_ref/ref_seqpath, built by oracle/Makefile from /root/reference/src with g++ 13.3 /
libstdc++ 13).  Run in the build container only: /root/reference does not exist on the GPU box,
which is why the outputs are committed.

    python tests/golden/make_golden.py

Fixtures
  ref_runs/<name>/        full output directory of a small reference run (all 12+ files)
  ref_runs/<name>.json    the parameters of that run
  config2.json            BASELINE config 2 (n=100, core=1000, L=1000, seed=1): md5 of every
                          output file, the small host-side files verbatim, first records of the
                          FASTA files (the 200 MB of sequence are not committed)
  trees.json              create_coalescent(n) dumps (parent, branch time) via ref_seqpath
  ref_distances.json      matched-parameter statistical fixture: pairwise p-distances of the
                          reference's own leaf sequences in the regime rate*L >> 1
"""
import hashlib
import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = ROOT / "oracle" / "_ref" / "pangenomesim"
SEQPATH = ROOT / "oracle" / "_ref" / "ref_seqpath"

SMALL_RUNS = {
    # name: parameters (order preserved on the command line)
    "config1_seed42": {"seed": 42},
    "acc_seed1": {"num-genomes": 10, "core-size": 5, "gene-length": 50, "theta": 5, "rho": 0.5, "seed": 1},
    "acc_seed2": {"num-genomes": 10, "core-size": 5, "gene-length": 50, "theta": 5, "rho": 0.5, "seed": 2},
    "acc_seed3": {"num-genomes": 12, "core-size": 3, "gene-length": 64, "theta": 20, "rho": 2, "seed": 3},
    "acc_seed4": {"num-genomes": 25, "core-size": 2, "gene-length": 33, "theta": 8, "rho": 0.25, "seed": 4},
    "acc_seed5": {"num-genomes": 7, "core-size": 0, "gene-length": 10, "theta": 30, "rho": 3, "mut-rate": 0.5, "seed": 5},
    "safeguard_like": {"gene-length": 100, "num-genomes": 10, "core-size": 20, "theta": 200, "rho": 0.2, "seed": 11},
}


def run_ref(out, params):
    arg = ",".join(f"{k}={v}" for k, v in params.items())
    subprocess.run([str(REF), "-o", str(out), "-p", arg], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def read_fasta(path):
    recs = []
    with open(path) as f:
        name = None
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                name = line[1:]
            else:
                recs.append((name, line))
    return recs


def main():
    if not REF.exists() or not SEQPATH.exists():
        sys.exit("build the reference first: make -C oracle")
    runs = HERE / "ref_runs"
    if runs.exists():
        shutil.rmtree(runs)
    runs.mkdir()
    for name, params in SMALL_RUNS.items():
        run_ref(runs / name, params)
        (runs / f"{name}.json").write_text(json.dumps(params, indent=1) + "\n")

    # ---- config 2
    with tempfile.TemporaryDirectory() as tmp:
        out = Path(tmp) / "c2"
        params = {"num-genomes": 100, "core-size": 1000, "gene-length": 1000, "seed": 1}
        run_ref(out, params)
        fx = {"params": params, "md5": {p.name: md5(p) for p in sorted(out.iterdir())}, "files": {}, "fasta_head": {}}
        for name in ("coalescent.newick", "genefrequency.gfs", "distance.mat", "reproducible.seed"):
            fx["files"][name] = (out / name).read_text()
        pan = (out / "panmatrix.tsv").read_text().splitlines()
        fx["panmatrix_header"] = pan[0]
        fx["panmatrix_rows_all_one"] = all(set(r.split("\t")[1:]) == {"1"} for r in pan[1:])
        for name in ("core.fasta", "ref.fasta", "genome001.fasta", "genome100.fasta"):
            fx["fasta_head"][name] = read_fasta(out / name)[:3]
        fx["genome_fasta_order"] = [n.split(".")[1] for n, _ in read_fasta(out / "genome001.fasta")]
        maf = (out / "alignment.maf").read_text().split("\n")
        fx["maf_head"] = [l[:60] for l in maf[:8]]
        fx["maf_first_block_names"] = [l.split(" ")[1] for l in maf[2:103]]
        fx["file_sizes"] = {p.name: p.stat().st_size for p in sorted(out.iterdir())}
        (HERE / "config2.json").write_text(json.dumps(fx, indent=1) + "\n")

    # ---- trees
    trees = {}
    with tempfile.TemporaryDirectory() as tmp:
        for n, seed in ((3, 42), (10, 1), (100, 1), (1000, 7)):
            dump = Path(tmp) / "t.txt"
            subprocess.run([str(SEQPATH), str(n), "10", "0", "0.01", str(seed), str(dump)], check=True,
                           stdout=subprocess.DEVNULL)
            parent, time = [], []
            for line in dump.read_text().splitlines():
                _, p, t, _ = line.split()
                parent.append(int(p))
                time.append(float(t).hex())
            trees[f"n{n}_seed{seed}"] = {"n": n, "seed": seed, "parent": parent, "time_hex": time}
    (HERE / "trees.json").write_text(json.dumps(trees) + "\n")

    # ---- matched-parameter distances (rate*L >> 1 on every branch)
    with tempfile.TemporaryDirectory() as tmp:
        out = Path(tmp) / "d"
        params = {"num-genomes": 6, "core-size": 2, "gene-length": 500000, "mut-rate": 0.2, "seed": 9}
        run_ref(out, params)
        seqs = {}
        for g in range(6):
            recs = dict(read_fasta(out / f"genome{g+1:03d}.fasta"))
            seqs[g] = "".join(recs[f"genome{g+1:03d}.{k}"] for k in ("0", "1"))
        S = len(seqs[0])
        pd = [[sum(a != b for a, b in zip(seqs[i], seqs[j])) / S for j in range(6)] for i in range(6)]
        dm = [l.split()[1:] for l in (out / "distance.mat").read_text().splitlines()[1:]]
        dump = Path(tmp) / "t.txt"
        subprocess.run([str(SEQPATH), "6", "10", "0", "0.2", "9", str(dump)], check=True, stdout=subprocess.DEVNULL)
        parent, time = [], []
        for line in dump.read_text().splitlines():
            _, p, t, _ = line.split()
            parent.append(int(p))
            time.append(float(t))
        (HERE / "ref_distances.json").write_text(json.dumps(
            {"params": params, "sites": S, "p_distance": pd, "distance_mat": [[float(x) for x in r] for r in dm],
             "parent": parent, "time": time, "newick": (out / "coalescent.newick").read_text()}, indent=1) + "\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
