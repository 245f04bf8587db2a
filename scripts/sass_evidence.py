#!/usr/bin/env python3
"""Kernel artefacts for profiles/: ptxas -v resource lines and SASS excerpts of the hot kernels,
taken from the objects `make lib` has just built (no GPU needed).

    python scripts/sass_evidence.py profiles/r02
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "build" / "obj"
KERNELS = [  # (object, substring of the demangled name, what to show)
    ("kernels.o", "k1_walk_dense<3, 2, true, true>", "dense walk, quads, 3 CTAs/SM (config 4 default)"),
    ("kernels.o", "k1_top_masks", "origin rows + masks of the top branches"),
    ("kernels.o", "k1_walk_sparse", "sparse (accessory) walk"),
    ("mismatch_tc.o", "k2_tc_pair<", "pairwise mismatch on the tensor cores, CTA pair (default: FP4)"),
    ("mismatch_tc.o", "k2_expand_fp4", "operand expansion for the FP4 tensor kernel"),
    ("mismatch.o", "k2_mismatch_tiles", "pairwise mismatch tiles (popcount kernel)"),
    ("mismatch.o", "k2_split_planes", "bit-plane split"),
    ("format.o", "k3_maf", "MAF formatter"),
    ("format.o", "k3_genome_fasta", "genome FASTA formatter"),
    ("format.o", "k3_gather_maf_rows", "row gather for host-expanded MAF"),
]
INTERESTING = re.compile(r"\b(LDG|STG|LDS|STS|LDGSTS|LDGDEPBAR|POPC|ATOM|ATOMG|RED|REDG|BAR|UTMA\w*|UBLKCP|UTC\w*|HMMA|IMMA|LDL|STL|CCTL|PREFETCH|LDTM|STTM|ACQBULK|PREEXIT|SYNCS|UCGABAR\w*)\b")


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    out = Path(sys.argv[1])
    out.parent.mkdir(parents=True, exist_ok=True)
    # ---- ptxas -v
    lines = []
    for log in sorted(OBJ.glob("*.ptxas.log")):
        text = log.read_text().splitlines()
        for i, ln in enumerate(text):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if m:
                spill = text[i + 2].strip() if i + 2 < len(text) else ""
                used = text[i + 3].strip() if i + 3 < len(text) else ""
                lines.append(f"{log.name[:-10]:9s} {demangle(m.group(1))}\n          {used.replace('ptxas info    : ', '')}\n          {spill}")
    Path(str(out) + "_ptxas.txt").write_text(
        "# nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v (make lib)\n" + "\n".join(lines) + "\n")
    # ---- SASS
    md = ["# SASS excerpts (cuobjdump -sass of the objects `make lib` built for sm_100a)\n",
          "Instruction mix over the whole kernel, then its memory, barrier and POPC instructions in program order (first 30).\n"]
    for obj, pat, what in KERNELS:
        funcs = subprocess.run(["cuobjdump", "-sass", str(OBJ / obj)], capture_output=True, text=True).stdout
        blocks = re.split(r"\n\s*Function : ", funcs)
        for b in blocks[1:]:
            name = b.split("\n", 1)[0].strip()
            if pat not in demangle(name):
                continue
            ins = [ln for ln in b.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
            ops = Counter()
            for ln in ins:
                m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if m:
                    op = m.group(1)
                    ops[op if op.startswith(("LDG", "STG", "LDS", "STS")) else op.split(".")[0]] += 1
            md.append(f"\n## `{demangle(name)[:110]}` -- {what}\n")
            md.append(f"{len(ins)} instructions. Mix: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(18)) + "\n")
            wide = [k for k in ops if k.startswith(("LDG", "STG")) and ("128" in k or "256" in k)]
            if wide:
                md.append("Wide global accesses: " + ", ".join(f"`{k}` x{ops[k]}" for k in sorted(wide)) + "\n")
            keep = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in ins if INTERESTING.search(ln)]
            md.append("```\n" + "\n".join(keep[:30]) + "\n```\n")
    Path(str(out) + "_sass.md").write_text("".join(md))
    print("wrote", out, "_ptxas.txt / _sass.md")


if __name__ == "__main__":
    main()
