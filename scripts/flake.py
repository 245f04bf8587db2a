#!/usr/bin/env python3
"""stress: sharded host-expanded CLI runs against the one-engine run; on a mismatch say WHICH records differ and how"""
import subprocess, sys, tempfile, os, shutil, collections
BIN = sys.argv[1]; RUNS = int(sys.argv[2]); TAG = sys.argv[3] if len(sys.argv) > 3 else ""
P = "gene-length=100,num-genomes=10,core-size=20,theta=200,rho=0.2,seed=11"
T = tempfile.mkdtemp()
def run(out, *flags):
    r = subprocess.run([BIN, "-o", out, "-p", P, "-v", *flags], capture_output=True, text=True)
    return r.stderr
run(T + "/one", "--gpu-format")
def records(path):
    out = {}
    with open(path) as f:
        lines = f.read().split("\n")
    for i in range(0, len(lines) - 1, 2):
        out[lines[i]] = lines[i + 1]
    return out
bad = 0
for i in range(RUNS):
    shutil.rmtree(T + "/many", ignore_errors=True)
    err = run(T + "/many", "--devices", "3")
    wrong = collections.OrderedDict()
    for g in range(1, 11):
        name = "genome%03d.fasta" % g
        a, b = records(f"{T}/one/{name}"), records(f"{T}/many/{name}")
        if list(a) != list(b):
            print(TAG, "run", i, name, "record NAMES differ"); bad += 1; continue
        for k in a:
            if a[k] != b[k]:
                wrong.setdefault(k.split(".")[1], []).append((g, a[k], b[k]))
    if wrong:
        bad += 1
        ids = [int(x) for x in wrong]
        print(TAG, "run", i, "wrong gene ids:", len(ids), "min", min(ids), "max", max(ids), "first", ids[:12])
        allseq = {}
        for g in range(1, 11):
            for k, v in records(f"{T}/one/genome%03d.fasta" % g).items(): allseq.setdefault(v, k)
        for gid, lst in list(wrong.items())[:4]:
            g, want, got = lst[0]
            nd = sum(1 for x, y in zip(want, got) if x != y)
            print(TAG, "  gene", gid, "genomes", [x[0] for x in lst], "sites differing", nd, "of", len(want), "got is", allseq.get(got, "no other record's sequence"), "| got", got[:40])
        print(TAG, "  device lines:", [l for l in err.split("\n") if l.startswith("device")])
print(TAG, BIN, bad, "bad runs of", RUNS)
shutil.rmtree(T, ignore_errors=True)
