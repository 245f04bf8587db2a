#!/usr/bin/env python3
"""bench.py -- headline benchmark of the pangenomesim sequence path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 4|3|5]   # this repo's CUDA engine
    python bench.py --impl reference [--steps K] [--warmup W]               # the reference's CPU code

Metric (BASELINE.json): simulated leaf nucleotides per second.  Workload: BASELINE config 4
(num-genomes=10000, core-size=5000, gene-length=1000, default mut-rate) -- 12.9 GB of 2-bit rows
rewritten every step, far larger than L2, so no flush is needed between steps.  One step = one
pgs_sweep (origin rows + every branch of the coalescent: k1_top_masks + k1_walk_dense, and
k1_walk_sparse for accessory genes) with tree and tables resident.
With N > 1 (torchrun, one rank per GPU) config 4 is taken AS WRITTEN: its 5000 genes are sharded N
ways ("scaling": "strong"; genes are independent, no collective in the sweep); the weak-scaling
leg (5000 genes per rank) is reported beside it.  --config 3 / 5 run the other BASELINE configs
(3: large accessory genome through the host model; 5: 50000 genomes at a high mutation rate).

Besides the contract keys the line carries
  roofline      K1's achieved algorithmic HBM GB/s against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference code (oracle/_ref/ref_seqpath) on one host core, bounded sample
  e2e           the same metric through the C ABI from host tables to host text: pgs_set_tree +
                pgs_set_genes (H2D) + pgs_sweep + pgs_write_outputs (every genomeNNN.fasta and
                alignment.maf byte delivered in host memory)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "simulated_leaf_nucleotides_per_sec"
UNIT = "leaf-nt/s"
CFG = {"num_genomes": 10000, "core_size": 5000, "gene_length": 1000, "mut_rate": 0.01, "seed": 1}


_REAL_STDOUT = None


def emit(text):
    """the bench line, to the process's original stdout (main() points fd 1 at stderr: library chatter)"""
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def _jsonable(x):
    """numpy scalars / arrays that slipped into the result line"""
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, np.ndarray):
        return x.tolist()
    return str(x)


CONFIGS = {
    # BASELINE.json configs[3]: the headline
    4: {"num_genomes": 10000, "core_size": 5000, "gene_length": 1000, "mut_rate": 0.01, "seed": 1},
    # configs[2]: theta / rho raised for a large accessory genome (SURVEY.md §8d suggests theta=2000, rho=1)
    3: {"num_genomes": 1000, "core_size": 3000, "gene_length": 1500, "mut_rate": 0.01, "seed": 1, "theta": 2000, "rho": 1},
    # configs[4]: memory pressure and a high mutation rate
    5: {"num_genomes": 50000, "core_size": 2000, "gene_length": 2000, "mut_rate": 0.5, "seed": 3},
}


def workload_name(n_gpus, weak=False, config=4):
    c = CONFIGS[config] if config != 4 else CFG
    extra = "".join(f" {k}={c[k]}" for k in ("theta", "rho") if k in c)
    shard = ""
    if n_gpus > 1:
        shard = (f" x{n_gpus} (weak: one such gene set per GPU)" if weak else f" (its genes sharded {n_gpus}-way, one shard per GPU)")
    return (f"config{config}: num-genomes={c['num_genomes']} core-size={c['core_size']}{shard} gene-length={c['gene_length']}"
            f" mut-rate={c['mut_rate']}{extra} seed={c['seed']}")


def bench_config(n_gpus, config=4):
    """`config` of the JSON line -- the same dictionary from both arms (--impl b200 / reference); what one
    rank holds is in the GPU arm's `shard` key, the reference arm's sample in its `cpu_baseline`"""
    c = CFG if config == 4 else CONFIGS[config]
    store_gb = c["num_genomes"] * c["core_size"] * ((c["gene_length"] + 63) // 64) * 16 / 1e9
    return {"workload": workload_name(n_gpus, config=config),
            "l2": f"inputs_larger_than_l2: every step rewrites the whole row store ({store_gb:.1f} GB of genome rows over {n_gpus} GPU(s), L2 = 126 MB)"}


# ------------------------------------------------------------------------------------------------
# synthetic inputs: the reference's own coalescent (create_coalescent, img.cxx:71-116) as drawn by
# the retained host code for seed 1 (pangenomesim_b200/host/model.cpp, bit-exact against the
# reference: tests/test_host_cli.py)

def coalescent(n, seed, mut_rate):
    from pangenomesim_b200.capi import host_coalescent
    return host_coalescent(n, seed, mut_rate)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baselines: the reference's own code (never this repo's kernels)

def ref_seqpath_run(n, length, genes, mut_rate, seed, procs=1):
    """Run `procs` copies of oracle/_ref/ref_seqpath (unmodified reference objects).  Returns
    (leaf_nt_total, seconds_wall, per-process json)."""
    exe = ROOT / "oracle" / "_ref" / "ref_seqpath"
    cmd = [str(exe), str(n), str(length), str(genes), str(mut_rate), str(seed)]
    t0 = time.perf_counter()
    ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True) for _ in range(procs)]
    outs = [json.loads(p.communicate()[0]) for p in ps]
    wall = time.perf_counter() - t0
    return sum(o["leaf_nt"] for o in outs), wall, outs


def port_run(n, length, genes, mut_rate, seed):
    """Fallback when the compiled reference did not travel: the plain-C restatement in oracle/."""
    from oracle import oracle_py
    e = oracle_py.RefEngine(seed)
    parent, left, right, tm = e.coalescent(n)
    leaf_index = np.where(np.arange(2 * n - 1) < n, np.arange(2 * n - 1), -1)
    t0 = time.perf_counter()
    for _ in range(genes):
        e.sim_core_gene(2 * n - 2, left, right, tm, leaf_index, mut_rate, length, n)
    return float(n) * length * genes, time.perf_counter() - t0


def cpu_baseline(sample_genes):
    n, L = CFG["num_genomes"], CFG["gene_length"]
    have_ref = (ROOT / "oracle" / "_ref" / "ref_seqpath").exists()
    if have_ref:
        nt, wall, outs = ref_seqpath_run(n, L, sample_genes, CFG["mut_rate"], CFG["seed"])
        sec = outs[0]["seconds"]  # simulate-only time inside the process (excludes tree setup)
        kind = "reference"
    else:
        nt, sec = port_run(n, L, sample_genes, CFG["mut_rate"], CFG["seed"])
        kind = "port"
    out = {"value": nt / sec, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{sample_genes} of {CFG['core_size']} core genes over the full {n}-genome coalescent "
                     f"(gene::gene + gene::mutate per branch, leaves discarded), {sec:.1f} s"}
    # T1 of SURVEY.md §8d: the unmodified reference CLI end to end (all 104 output files) at config 2,
    # the largest BASELINE config it can run
    exe = ROOT / "oracle" / "_ref" / "pangenomesim"
    if exe.exists():
        import shutil
        import tempfile
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        tmp = tempfile.mkdtemp(prefix="pgs_ref_", dir=base)
        try:
            t0 = time.perf_counter()
            subprocess.run([str(exe), "-o", tmp, "-p", "num-genomes=100,core-size=1000,gene-length=1000,seed=1"],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
            wall = time.perf_counter() - t0
            out["reference_cli_config2"] = {"value": 1e8 / wall, "unit": UNIT, "seconds": wall,
                                            "what": "unmodified reference binary, end to end incl. 200 MB of output"}
        except (subprocess.SubprocessError, OSError):
            pass
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, L = CFG["num_genomes"], CFG["gene_length"]
    have_ref = (ROOT / "oracle" / "_ref" / "ref_seqpath").exists()
    cores = os.cpu_count() or 1
    genes_per_proc = args.ref_genes
    times = []
    nt = 0.0
    for step in range(args.warmup + args.steps):
        if have_ref:
            nt, wall, _ = ref_seqpath_run(n, L, genes_per_proc, CFG["mut_rate"], CFG["seed"], procs=cores)
        else:
            nt, wall = port_run(n, L, genes_per_proc, CFG["mut_rate"], CFG["seed"])
        if step >= args.warmup:
            times.append(wall)
    total = sum(times)
    value = nt * len(times) / total
    used = cores if have_ref else 1
    sample = (f"{used} independent processes x {genes_per_proc} core genes each per step "
              f"(the reference is single-threaded: global engine, global.h:6), full {n}-genome coalescent, "
              f"wall clock incl. create_coalescent")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args.gpus, 4),  # the same dictionary as the GPU arm's (the sample is described in cpu_baseline)
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "reference" if have_ref else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's arm

def mismatch_allreduce_leg(torch, dist, Engine, rank, world, local):
    """K2 on every rank's gene shard, partial n x n counts summed with NCCL over NVLink; checked
    against the sum of the partials gathered on rank 0 and against the JC expectation."""
    n, L, genes = 2000, 1000, 250
    parent, rate, leaf = coalescent(n, 3, 0.05)
    eng = Engine(7, L, device=local)
    eng.set_tree(parent, rate, leaf)
    eng.set_core_genes(genes, 2 * n - 2, first_id=rank * genes)
    eng.sweep()
    part = torch.zeros((n, n), dtype=torch.int32, device="cuda")
    sites = eng.mismatch_device(part.data_ptr())
    torch.cuda.synchronize()
    k2_ms = eng.stats()["mismatch_ms"]
    before = part.to(torch.int64).sum()
    dist.all_reduce(before)
    warm = torch.zeros_like(part)
    dist.all_reduce(warm, op=dist.ReduceOp.SUM)  # NCCL sets up its channels for this size on first use
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    ev1.record()
    torch.cuda.synchronize()
    ok = bool(part.to(torch.int64).sum().item() == before.item())
    # JC check of one distant pair on the summed matrix
    S = sites * world
    d, v, seen = 0.0, 0, {}
    acc = 0.0
    while v != -1:
        seen[v] = acc
        acc += rate[v]
        v = int(parent[v])
    v, accb = 1, 0.0
    while v not in seen:
        accb += rate[v]
        v = int(parent[v])
    d = seen[v] + accb
    p = 0.75 * (1 - np.exp(-4 * d / 3))
    obs = part[0, 1].item() / S
    ok = bool(ok and abs(obs - p) < 5 * np.sqrt(p * (1 - p) / S) + 1e-9)
    eng.close()
    return {"ok": ok, "n_genomes": n, "genes_per_rank": genes, "k2_ms": k2_ms, "allreduce_ms": ev0.elapsed_time(ev1),
            "bytes": int(part.numel() * 4), "pair01_observed": obs, "pair01_jc_expected": float(p)}


def k2_leg(torch, dist, eng, stats, rank, world, peaks, popc_runs=1):
    """K2 (pairwise mismatch, north_star part 3) on this rank's genes at the config's FULL size, and
    -- N > 1 -- the path's only collective at its real size: the NCCL all-reduce of the packed
    upper-triangular counts (pgs_mismatch_allreduce).  Measured with the library's default kernel
    (the +-1 contraction on the tensor cores, FP4 operands) and, beside it, with the FP8 variant and
    with the popcount-XOR kernel north_star names -- all three produce the same integers."""
    from pangenomesim_b200.capi import comm_unique_id
    n, L = stats["n_genomes"], eng.gene_length
    out = {}
    os.environ.pop("PGS_K2_IMPL", None)
    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, 0)
        eng.comm_init(bytes(ident.cpu().numpy().tobytes()), rank, world)
        eng.mismatch_allreduce(want_matrix=False)  # warm-up: NCCL sets up its channels for this size
        _, sites, ar_ms = eng.mismatch_allreduce(want_matrix=False)
        t = torch.tensor([ar_ms, eng.stats()["mismatch_ms"], float(sites)], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ar_ms, k2_ms, sites_all = float(tm[0]), float(tm[1]), float(t[2])
        words = 4096 * ((n + 63) // 64) * ((n + 63) // 64 + 1) // 2
        out["allreduce"] = {"ms": ar_ms, "bytes": words * 4, "algbw_gb_per_s": words * 4 / (ar_ms * 1e-3) / 1e9,
                            "busbw_gb_per_s": words * 4 / (ar_ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
                            "what": "ncclAllReduce(sum, uint32) of every rank's packed upper-triangular mismatch counts, in place, "
                                    "max over ranks of the CUDA-event time"}
    else:
        eng.mismatch_tiles_device()
        eng.mismatch_tiles_device()
        k2_ms, sites_all = eng.stats()["mismatch_ms"], float(stats["n_dense_genes"]) * L
    pairs = n * (n - 1) / 2.0
    site_pairs_rank = pairs * (sites_all / world)
    # Tensor-core bound: three +-1 entries per site (the tetrahedron encoding, mismatch_tc.cu) = 3 multiply-adds = 6 flop
    # per site pair.  Peak: B200's dense FP4 rate; no FP4 GEMM was measured on this pool, so the figure is the nominal
    # 9 PFLOP/s of B200_PROFILING.md (the measured cuBLAS bf16 burst of MEASURED_PEAKS.json x 4 is printed beside it).
    fp4_peak = 9.0e15
    bf16 = float(peaks.get("bf16_tflops", 0.0)) * 1e12
    flops = 6.0 * site_pairs_rank / (k2_ms * 1e-3)
    out.update({"k2_ms": k2_ms, "kernel": "k2_expand_fp4 + k2_tc_pair<FP4> (tcgen05.mma.cta_group::2.kind::mxf4, unit block scales)",
                "site_pairs_total": pairs * sites_all, "site_pairs_per_s": pairs * sites_all / (k2_ms * 1e-3),
                "roofline": {"bound": "tensor", "achieved": flops / 1e12, "peak": fp4_peak / 1e12, "unit": "TFLOP/s per GPU (FP4, dense)",
                             "frac": flops / fp4_peak, "peak_is": "nominal (B200_PROFILING.md)",
                             "frac_of_4x_measured_bf16_burst": (flops / (4 * bf16)) if bf16 else None,
                             "definition": "6 flop per site pair (3 multiply-adds: a base is the +-1 vector ((-1)^lo, (-1)^hi, (-1)^(lo^hi)); "
                                           "mismatches = (3 S - dot) / 4) x n (n - 1) / 2 pairs x this rank's sites / CUDA-event time of the "
                                           "whole pgs_mismatch_tiles_device call (operand expansion included)"},
                "what": f"{n} genomes x this rank's dense genes, max over ranks; counts kept as packed upper-triangular tiles on the device"})
    # the other kernels on the same rows (rank-local, no collective)
    variants = {}
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    for impl, runs in (("tensor", 2), ("popc", popc_runs)):
        if runs <= 0:
            continue
        os.environ["PGS_K2_IMPL"] = impl
        for _ in range(runs):
            eng.mismatch_tiles_device()
        ms = eng.stats()["mismatch_ms"]
        v = {"k2_ms": ms, "site_pairs_per_s_this_rank": site_pairs_rank / (ms * 1e-3)}
        if impl == "popc":
            # per THREE 32-site masks of a genome pair: 8 ALU-pipe instructions (2 cycles each per warp) and 2 POPC on the
            # XU pipe (8 cycles each): the pipes are level, 96 site pairs per lane per 16 cycles (profiles/r02_k2_full.md)
            pipe_peak = 148 * 4 * 32 * 96 / 16.0 * sm_mhz * 1e6
            v.update({"kernel": "k2_split_planes + k2_mismatch_tiles (XOR / LOP3 / carry-save adder / POPC)",
                      "roofline": {"bound": "alu+xu", "achieved": v["site_pairs_per_s_this_rank"], "peak": pipe_peak, "unit": "site-pairs/s per GPU",
                                   "frac": v["site_pairs_per_s_this_rank"] / pipe_peak}})
        else:
            v.update({"kernel": "k2_expand_pm1 + k2_tc_pair<E4M3> (kind::f8f6f4)",
                      "roofline": {"bound": "tensor", "achieved": 6.0 * v["site_pairs_per_s_this_rank"] / 1e12, "peak": 4500.0,
                                   "unit": "TFLOP/s per GPU (FP8, dense, nominal)", "frac": 6.0 * v["site_pairs_per_s_this_rank"] / 4.5e15}})
        variants[impl] = v
    os.environ.pop("PGS_K2_IMPL", None)
    out["variants"] = variants
    return out


def cli_multi_device_leg(world):
    """The product CLI on real devices (rank 0 only, the other ranks' engines are closed and they wait
    on the CPU): pangenomesim-b200 --devices N on N distinct GPUs against --devices 1, every output
    file compared byte for byte; --check-distances sums the engines' K2 counts with NCCL."""
    import filecmp
    import shutil
    import tempfile
    exe = ROOT / "pangenomesim_b200" / "bin" / "pangenomesim-b200"
    if not exe.exists():
        return {"skipped": "pangenomesim_b200/bin/pangenomesim-b200 not built"}
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="pgs_cli_", dir=base)
    params = "num-genomes=1000,core-size=256,gene-length=1000,theta=30,rho=0.5,mut-rate=0.02,seed=7"
    out = {"params": params, "devices": world}
    try:
        for name, extra in (("one", ["--devices", "1"]), ("many", ["--devices", str(world), "--check-distances"])):
            t0 = time.perf_counter()
            r = subprocess.run([str(exe), "-o", f"{tmp}/{name}", "-p", params, "--rng-fast", "-v", *extra],
                               capture_output=True, text=True, timeout=600)
            out[f"seconds_{name}"] = time.perf_counter() - t0
            if r.returncode != 0:
                out["ok"] = False
                out["error"] = r.stderr[-400:]
                return out
            if name == "many":
                out["check_distances"] = next((ln for ln in r.stderr.splitlines() if ln.startswith("check-distances")), None)
                out["device_lines"] = [ln for ln in r.stderr.splitlines() if ln.startswith("device ")][:8]
        names = sorted(os.listdir(f"{tmp}/one"))
        extra_files = sorted(set(os.listdir(f"{tmp}/many")) - set(names))
        differ = [nm for nm in names if not filecmp.cmp(f"{tmp}/one/{nm}", f"{tmp}/many/{nm}", shallow=False)]
        out.update(ok=(not differ and extra_files == ["distance.observed.mat"]), files_compared=len(names), differing=differ[:5],
                   bytes=sum(os.path.getsize(f"{tmp}/one/{nm}") for nm in names))
    except (subprocess.SubprocessError, OSError) as ex:
        out.update(ok=False, error=f"{type(ex).__name__}: {ex}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def files_leg(eng, parent, rate, leaf, genes, root, n, L):
    """host tables -> sweep -> genomeNNN.fasta + alignment.maf written as FILES (pwrite / mapped
    writes by the library's writer threads) into a scratch directory that is removed afterwards.
    Bounded: `genes` of config 4's genes, so that ~2 x genes x 10 MB fit the scratch file system."""
    import shutil
    import tempfile
    from pangenomesim_b200.capi import OUT_GENOMES, OUT_HOST_EXPAND, OUT_MAF, PGS_ALL_GENOMES
    need = 2.2 * 2 * genes * n * (L + 40)  # bytes written, with head-room
    base = None
    for cand in (os.environ.get("PGS_BENCH_FILES_DIR"), "/dev/shm", tempfile.gettempdir()):
        if cand and os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
            base = cand
            break
    if base is None:
        return {"skipped": f"no scratch directory with {need / 1e9:.0f} GB free"}
    d = tempfile.mkdtemp(prefix="pgs_bench_", dir=base)
    try:
        gene_id = np.arange(genes, dtype=np.int64)
        origin = np.full(genes, root, np.int32)
        car_off = np.arange(genes + 1, dtype=np.int64)
        car = np.full(genes, PGS_ALL_GENOMES, np.int32)
        out = {}
        # (the first pass over a fresh box's page cache pays the first touch of 20 GB of guest memory -- 2.6 s against
        # 1.8-2.0 s for the same call afterwards: one untimed warm-up pass, as every other leg has)
        for mode, extra in (("warm-up", OUT_HOST_EXPAND), ("host_expand", OUT_HOST_EXPAND), ("device_formatter", 0)):
            sub = os.path.join(d, mode)
            os.mkdir(sub)
            t0 = time.perf_counter()
            eng.set_tree(parent, rate, leaf)
            eng.set_genes(gene_id, origin, car_off, car)
            eng.sweep()
            rep = eng.write_outputs(sub, OUT_GENOMES | OUT_MAF | extra, np.arange(genes, dtype=np.int64), [])
            wall = time.perf_counter() - t0
            out[mode] = {"value": float(n) * genes * L / wall, "unit": UNIT, "seconds": wall, "bytes": int(rep["bytes"]),
                         "files": int(rep["files"]), "gb_per_s": rep["bytes"] / wall / 1e9}
            shutil.rmtree(sub, ignore_errors=True)
        out.pop("warm-up", None)
        best = max(out, key=lambda k: out[k]["value"])
        res = dict(out[best])
        res.update(mode=best, modes=out, dir=base,
                   what=f"{n} genomes x the first {genes} genes of config 4: host tables -> sweep -> "
                        f"{res['files']} files on {base} (page cache), then removed, after one untimed warm-up pass; host_expand = packed rows over PCIe, text written by the "
                        f"host's cores (whole genomeNNN.fasta files, the MAF through a mapping); device_formatter = K3's text over PCIe")
        return res
    except Exception as ex:  # a full or read-only scratch directory must not cost the bench line
        return {"skipped": f"{type(ex).__name__}: {ex}"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def make_workload(config, rank, world, weak=False, shard_of=0):
    """Tree and gene table of one rank.  Config 4 / 5: the reference's coalescent and dense core genes,
    sharded by contiguous gene ranges (strong) or one full gene set per rank (weak).  Config 3: the
    retained host model (coalescent + IMG gain/loss), its gene table cut into contiguous ranges."""
    from pangenomesim_b200.capi import PGS_ALL_GENOMES, HostModel
    from pangenomesim_b200.sharding import shard_range
    c = CFG if config == 4 else CONFIGS[config]
    n, L, G = c["num_genomes"], c["gene_length"], c["core_size"]
    w = {"n": n, "L": L, "seed": c["seed"], "config": config}
    if config == 3:
        m = HostModel("fast", num_genomes=n, core_size=G, gene_length=L, theta=c["theta"], rho=c["rho"], seed=c["seed"],
                      mut_rate=c["mut_rate"])
        g0, g1 = shard_range(m.n_genes, rank, world)
        off = m.carrier_off[g0:g1 + 1]
        w.update(parent=m.parent, rate=m.rate, leaf=m.leaf_genome, all_order=m.all_order,
                 genes=(m.gene_id[g0:g1], m.origin[g0:g1], off - off[0], m.carriers[off[0]:off[-1]], m.all_order),
                 n_genes_total=m.n_genes, leaf_nt_total=float(m.genome_gene_count.sum()) * L,
                 ref=np.arange(g1 - g0, dtype=np.int64))
        m.close()
        return w
    parent, rate, leaf = coalescent(n, c["seed"], c["mut_rate"])
    root = 2 * n - 2
    if weak:
        g0, g1 = shard_range(G * world, rank, world)
    else:
        g0, g1 = shard_range(G, rank, world)
        if shard_of > 1 and world == 1:
            g0, g1 = shard_range(G, 0, shard_of)
    k = g1 - g0
    w.update(parent=parent, rate=rate, leaf=leaf, all_order=None,
             genes=(np.arange(g0, g1, dtype=np.int64), np.full(k, root, np.int32), np.arange(k + 1, dtype=np.int64),
                    np.full(k, PGS_ALL_GENOMES, np.int32), None),
             n_genes_total=G * (world if weak else 1), leaf_nt_total=float(n) * L * G * (world if weak else 1),
             ref=np.arange(k, dtype=np.int64))
    return w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[3, 4, 5], help="BASELINE.json config (4 = headline)")
    ap.add_argument("--e2e-steps", type=int, default=2, help="end-to-end steps (each moves ~100 GB over PCIe)")
    ap.add_argument("--cpu-genes", type=int, default=60, help="gene sample of the cpu_baseline leg (~0.2 s per gene)")
    ap.add_argument("--ref-genes", type=int, default=8, help="genes per process per step of --impl reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--level-sync", action="store_true", help="sweep level by level and keep all internal rows (PGS_FLAG_KEEP_INTERNAL)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the bounded write-to-files leg (N=1 only)")
    ap.add_argument("--no-mismatch", action="store_true", help="skip the K2 legs")
    ap.add_argument("--no-cli", action="store_true", help="skip the multi-device CLI leg (N > 1)")
    ap.add_argument("--files-genes", type=int, default=1000, help="genes of the write-to-files leg (20 MB of text each)")
    ap.add_argument("--small", action="store_true", help="debug: config 2 sized workload")
    ap.add_argument("--shard-of", type=int, default=0, help="development: on ONE GPU, sweep only rank 0's shard of a SHARD_OF-way split "
                                                                "of config 4 (what each GPU of a strong-scaling run does)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on
    # the first communicator, whoever creates it): file descriptor 1 is pointed at stderr for the whole
    # run and the line goes to a private copy of the original stdout (emit()).
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.small:
        CFG.update(num_genomes=100, core_size=1000)
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from pangenomesim_b200 import Engine
    from pangenomesim_b200.capi import OUT_DISCARD, OUT_GENOMES, OUT_HOST_EXPAND, OUT_MAF, OUT_PACKED

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # the headline: the config as written, its genes sharded over the ranks (strong scaling)
    w = make_workload(args.config, rank, world, weak=False, shard_of=args.shard_of)
    n, L = w["n"], w["L"]
    parent, rate, leaf = w["parent"], w["rate"], w["leaf"]

    stream = torch.cuda.Stream()
    flags = 1 if os.environ.get("PGS_NO_GRAPH") else 0  # PGS_FLAG_NO_GRAPH: plain launches (profiling)
    if args.level_sync:
        flags |= 2  # PGS_FLAG_KEEP_INTERNAL: level-synchronous sweep, every internal row kept in HBM
    eng = Engine(w["seed"], L, device=local, flags=flags)
    eng.set_stream(stream.cuda_stream)
    eng.set_tree(parent, rate, leaf)
    eng.set_genes(*w["genes"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_sweeps():
        for _ in range(args.warmup):
            eng.sweep()
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            eng.sweep()
        ev1.record(stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    # ---- kernel-resident steps
    # the first sweep of a plan is launched kernel by kernel with an event between the dense and the
    # sparse genes: K1's split (later sweeps replay one CUDA graph)
    eng.sweep()
    split = eng.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    total_ms = timed_sweeps()
    k1_ms, k0_ms = [], []
    for _ in range(3):
        eng.sweep()
        st = eng.stats()
        k1_ms.append(st["sweep_ms"])
        k0_ms.append(st["rootgen_ms"])
    if rank == 0:
        time.sleep(0.15)
        clocks = sampler.stop()
    st = eng.stats()
    leaf_nt_all = w["leaf_nt_total"]
    value = leaf_nt_all * args.steps / (total_ms * 1e-3)

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # Algorithmic HBM bytes of one sweep on this rank.  Level-synchronous definition:
    # every row written once + every parent row read once, L/4 bytes each (0.375 B per site-edge).
    # The fused walks keep internal nodes on chip, so -- as §8d prescribes for that case -- the bound
    # switches to "every leaf row written once": 0.25 B per leaf nucleotide.
    level_bytes = (st["rows_written"] + st["rows_read"]) * L / 4.0
    leaf_bytes = st["leaf_nucleotides"] / 4.0
    walk = st["sweep_mode"] == 1
    alg_bytes = leaf_bytes if walk else level_bytes
    # K1's duration over the TIMED region: the step time (CUDA events around the K back-to-back
    # steps) less K0's share of a step (library events around K0 and K1 of single sweeps)
    k1_single = float(np.mean(k1_ms))
    k1 = (total_ms / args.steps) * k1_single / (k1_single + float(np.mean(k0_ms)))
    achieved = alg_bytes / (k1 * 1e-3) / 1e9
    traffic = None
    tr = ROOT / "profiles" / "k1_traffic.json"
    if tr.exists() and args.config == 4 and world == 1 and not args.shard_of:
        try:
            traffic = json.loads(tr.read_text()).get("dram_bytes_per_sweep_walk" if walk else "dram_bytes_per_sweep")
        except ValueError:
            traffic = None
    row_bytes = st["blocks_per_gene"] * 16
    roofline = {"bound": "hbm", "kernel": "k1_top_masks + k1_walk_dense (+ k1_walk_sparse)" if walk else "k1_sweep_dense / k1_sweep_sparse",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_sweep": alg_bytes, "k1_ms_per_sweep": k1, "k1_ms_single_sweep": k1_single,
                "launches_per_sweep": st["kernel_launches"],
                "definition": ("leaf-write-only bound, 0.25 B per leaf nucleotide (internal nodes stay on chip: fused walks), this rank's shard"
                               if walk else "level-synchronous bound, 0.375 B per site-edge"),
                "level_synchronous_definition": {"algorithmic_bytes_per_sweep": level_bytes,
                                                 "achieved": level_bytes / (k1 * 1e-3) / 1e9,
                                                 "frac": level_bytes / (k1 * 1e-3) / 1e9 / peak},
                "rows_stored_per_sweep": st["rows_stored"], "rows_loaded_per_sweep": st["rows_loaded"],
                "stored_bytes_per_sweep": st["rows_stored"] * row_bytes}
    if split["sparse_ms"] > 0 or split["dense_ms"] > 0:
        # K1 split (one plain sweep): the sparse genes against THEIR level-synchronous byte count
        # (every sparse row written + every sparse parent row read, L/4 B each)
        dense_rows = st["n_dense_genes"] * (st["n_nodes"] - 1)
        sp_bytes = (st["rows_written"] + st["rows_read"] - dense_rows - st["n_dense_genes"] * (st["n_genomes"] - 1)) * row_bytes
        roofline["split"] = {"dense_ms": split["dense_ms"], "sparse_ms": split["sparse_ms"],
                             "sparse_level_synchronous_bytes": sp_bytes,
                             "sparse_frac_of_peak": (sp_bytes / (split["sparse_ms"] * 1e-3) / 1e9 / peak) if split["sparse_ms"] > 0 else None,
                             "what": "first sweep of the plan, launched kernel by kernel; sparse bytes = rows written + parent rows read (padded rows)"}

    # ---- end to end through the C ABI with host tables in, host text out.  Two ways to the same
    # bytes: K3 formats on the GPU and the text crosses PCIe (4.1 x the packed rows), or
    # (PGS_OUT_HOST_EXPAND) the packed rows cross PCIe and the host's cores write the text.
    e2e = e2e_device = e2e_host = None
    if not args.no_e2e:
        h2d = sum(a.nbytes for a in (parent, rate, leaf) + tuple(x for x in w["genes"] if x is not None))
        threads = max(2, (os.cpu_count() or 2) // world)
        os.environ.setdefault("PGS_EXPAND_THREADS", str(threads))
        row_bytes_packed = st["blocks_per_gene"] * 16

        def e2e_run(what, describe):
            plan_s = []

            def e2e_step():
                t_plan = time.perf_counter()
                eng.set_tree(parent, rate, leaf)
                eng.set_genes(*w["genes"])
                plan_s.append(time.perf_counter() - t_plan)
                eng.sweep()
                return eng.write_outputs(None, what, w["ref"], [])
            e2e_step()  # warm-up (allocations, pinned buffers)
            barrier()
            t0 = time.perf_counter()
            rep = None
            for _ in range(args.e2e_steps):
                rep = e2e_step()
            barrier()
            wall = max_over_ranks(time.perf_counter() - t0)
            host = bool(what & OUT_HOST_EXPAND)
            # bytes that cross PCIe per step: the text itself, or the packed rows of every record / MAF line
            d2h = int(rep["d2h_bytes"])
            return {"value": leaf_nt_all * args.e2e_steps / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": d2h, "text_bytes_per_step": int(rep["bytes"]), "steps": args.e2e_steps,
                    "ms_per_step": 1e3 * wall / args.e2e_steps, "k3_device_ms_per_step": rep["device_ms"],
                    "files_per_step": int(rep["files"]), "plan_ms_per_step": 1e3 * float(np.mean(plan_s[1:])),
                    "text_gb_per_s": rep["bytes"] * args.e2e_steps / wall / 1e9, "d2h_gb_per_s": d2h * args.e2e_steps / wall / 1e9,
                    "what": "pgs_set_tree + pgs_set_genes + pgs_sweep + pgs_write_outputs(genomeNNN.fasta + alignment.maf) " + describe
                            + "; no filesystem; per-rank bytes"}
        e2e_device = e2e_run(OUT_GENOMES | OUT_MAF | OUT_DISCARD,
                             "formatted by K3 on the GPU, the text copied into pinned host memory and dropped")
        e2e_host = e2e_run(OUT_GENOMES | OUT_MAF | OUT_DISCARD | OUT_HOST_EXPAND,
                           f"PGS_OUT_HOST_EXPAND: packed rows copied into pinned host memory (genome order for the FASTA files, MAF line order), "
                           f"every genomeNNN.fasta and every run of 256 MAF lines written in full (headers + bases, SIMD) through L2-sized "
                           f"buffers by {threads} host threads, then dropped")
        e2e = dict(e2e_host if e2e_host["value"] >= e2e_device["value"] else e2e_device)
        e2e["mode"] = "host_expand" if e2e is not None and e2e_host["value"] >= e2e_device["value"] else "device_formatter"

    # ---- the same end to end with the optional packed side output instead of the text (genomes.pgs2:
    # the genomes' 2-bit rows copied straight from the store, an eighth of the bytes)
    e2e_packed = None
    if not args.no_e2e:
        def packed_step():
            eng.set_tree(parent, rate, leaf)
            eng.set_genes(*w["genes"])
            eng.sweep()
            return eng.write_outputs(None, OUT_PACKED | OUT_DISCARD, [], [])
        packed_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rep = packed_step()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e_packed = {"value": leaf_nt_all * args.e2e_steps / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / args.e2e_steps,
                      "d2h_bytes_per_step": int(rep["bytes"]), "d2h_gb_per_s": rep["bytes"] * args.e2e_steps / wall / 1e9,
                      "what": "pgs_set_tree + pgs_set_genes + pgs_sweep + pgs_write_outputs(PGS_OUT_PACKED: genomes.pgs2, "
                              "2-bit rows + index) into pinned host memory, bytes discarded; NOT a format of the reference"}

    # ---- end to end INTO FILES on a bounded slice of the workload (all 10 000 genomes, the first
    # --files-genes genes: 10 000 genomeNNN.fasta + one alignment.maf), rank 0 of a 1-GPU run only
    files = None
    if world == 1 and args.config == 4 and not args.no_e2e and not args.no_files:
        files = files_leg(eng, parent, rate, leaf, min(args.files_genes, CFG["core_size"]), 2 * n - 2, n, L)

    # ---- K2 at full size on this rank's shard (+ the all-reduce of the partial counts for N > 1)
    k2 = None
    if not args.no_mismatch and args.config != 3:
        eng.set_tree(parent, rate, leaf)
        eng.set_genes(*w["genes"])
        eng.sweep()
        k2 = k2_leg(torch, dist, eng, eng.stats(), rank, world, peaks)

    # ---- weak-scaling leg (N > 1): every rank sweeps a full gene set of the config
    weak = None
    if world > 1 and args.config != 3:
        ww = make_workload(args.config, rank, world, weak=True)
        eng.set_tree(parent, rate, leaf)
        eng.set_genes(*ww["genes"])
        ms = timed_sweeps() / args.steps
        weak = {"value": ww["leaf_nt_total"] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "scaling": "weak",
                "what": workload_name(world, weak=True, config=args.config) + ", no collective"}

    # ---- the path's only collective: all-reduce of the per-rank partial mismatch counts (K2),
    # untimed verification leg on a reduced problem (2000 genomes, 250 genes per rank)
    mm_leg = None
    if world > 1 and not args.no_mismatch:
        mm_leg = mismatch_allreduce_leg(torch, dist, Engine, rank, world, local)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.config == 4:
        cpu = cpu_baseline(args.cpu_genes if not args.small else 200)

    # ---- the product CLI with one engine per GPU (N > 1): rank 0 drives it, everybody else has let go
    # of its GPU and waits on the CPU (a gloo barrier: no kernel spins beside the CLI's own work)
    cli = None
    eng.close()
    if world > 1 and not args.no_e2e and not args.no_cli:
        torch.cuda.synchronize()
        cpu_group = dist.new_group(backend="gloo")
        dist.barrier(group=cpu_group)
        if rank == 0:
            cli = cli_multi_device_leg(world)
        dist.barrier(group=cpu_group)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": bench_config(world, args.config),
            "shard": {"l2": f"inputs_larger_than_l2 ({st['device_bytes'] / 1e9:.1f} GB of rows per GPU, rewritten every step)"
                            if st["device_bytes"] > 400e6 else
                            f"{st['device_bytes'] / 1e6:.0f} MB of rows per GPU, rewritten every step (not larger than L2)",
                      "rng": "Philox4x32-10 keyed (seed, edge, gene, 64-site block); 16-bit decision fields, one call per sibling pair and 4 blocks",
                      "genes_total": w["n_genes_total"], "genes_this_rank": st["n_genes"], "dense_genes_this_rank": st["n_dense_genes"],
                      "rows_per_gpu": st["rows_total"], "levels": st["levels"], "hbm_bytes_per_gpu": st["device_bytes"],
                      "sweep": "fused walks (k1_top_masks + k1_walk_dense, k1_walk_sparse)" if st["sweep_mode"] == 1 else "level-synchronous"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(st["kernel_launches"]) * args.steps * world, "clocks": clocks,
            "e2e_device_formatter": e2e_device, "e2e_host_expand": e2e_host,
            "mismatch": k2, "mismatch_allreduce_check": mm_leg, "weak_scaling": weak, "cli_multi_device": cli, "e2e_files": files, "e2e_packed": e2e_packed,
        }
        emit(json.dumps(line, default=_jsonable))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
