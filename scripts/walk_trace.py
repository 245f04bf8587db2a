#!/usr/bin/env python3
"""Development: summarise a PGS_WALK_TRACE file (per-CTA clocks of one k1_walk_dense launch).

    PGS_NO_GRAPH=1 PGS_WALK_TRACE=gpurun_out/trace.txt python bench.py --shard-of 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-mismatch
    python scripts/walk_trace.py gpurun_out/trace.txt
"""
import sys
import numpy as np

a = np.loadtxt(sys.argv[1], dtype=np.int64)
unit, slab, t0, t1, t2, sm = a.T
base = t0.min()
t0, t1, t2 = (t0 - base) / 1e3, (t1 - base) / 1e3, (t2 - base) / 1e3
print(f"CTAs {len(a)}  kernel span {t2.max():.1f} us  first start {t0.min():.1f}  last start {t0.max():.1f}")
pro = t1 - t0
dur = t2 - t0
print(f"prologue us: mean {pro.mean():.2f} median {np.median(pro):.2f} p90 {np.percentile(pro, 90):.2f} max {pro.max():.2f}")
print(f"CTA duration us: mean {dur.mean():.1f} max {dur.max():.1f} min {dur.min():.1f}; sum/444 = {dur.sum() / 444:.1f}")
# resident CTAs over time
ev = sorted([(t, 1) for t in t0] + [(t, -1) for t in t2])
grid = np.linspace(0, t2.max(), 23)
k = 0
cur = 0
out = []
for g in grid:
    while k < len(ev) and ev[k][0] <= g:
        cur += ev[k][1]
        k += 1
    out.append(cur)
print("resident CTAs at", " ".join(f"{g:.0f}:{c}" for g, c in zip(grid, out)))
# per-SM finish times
fin = np.array([t2[sm == s].max() for s in np.unique(sm)])
print(f"SM finish us: min {fin.min():.1f} mean {fin.mean():.1f} max {fin.max():.1f}")
# per unit: duration of its CTAs vs start
nu = unit.max() + 1
for u in list(range(0, nu, max(1, nu // 12))):
    m = unit == u
    print(f"unit {u}: start {t0[m].mean():.1f} dur {dur[m].mean():.1f} (min {dur[m].min():.1f} max {dur[m].max():.1f}) prologue {pro[m].mean():.2f}")
