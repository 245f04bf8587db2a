"""Synthetic inputs for the tests: Kingman coalescent trees in the reference's pool layout
(leaves 0..n-1, internal nodes in increasing time, root last; reference img.cxx:65-116) and an
IMG-like accessory gene table (gain at a node, carried by a subset of the leaves below)."""
from __future__ import annotations

import numpy as np


def coalescent(n: int, seed: int, mut_rate: float = 0.01):
    """Returns parent[int32 2n-1], rate[float64] (= branch time * mut_rate), leaf_genome[int32]."""
    rng = np.random.default_rng(seed)
    N = 2 * n - 1
    parent = np.full(N, -1, dtype=np.int32)
    abs_time = np.zeros(N)
    active = list(range(n))
    t = 0.0
    for i in range(n - 1):
        k = n - i
        p = n + i
        a = active.pop(rng.integers(len(active)))
        b = active.pop(rng.integers(len(active)))
        parent[a] = p
        parent[b] = p
        active.append(p)
        t += rng.exponential(1.0 / (k * (k - 1) / 2))
        abs_time[p] = t
    rate = np.zeros(N)
    for v in range(N - 1):
        rate[v] = (abs_time[parent[v]] - abs_time[v]) * mut_rate
    leaf_genome = np.full(N, -1, dtype=np.int32)
    leaf_genome[:n] = np.arange(n)
    return parent, rate, leaf_genome


def children_of(parent):
    ch = [[] for _ in parent]
    for v, p in enumerate(parent):
        if p >= 0:
            ch[p].append(v)
    return ch


def leaves_below(parent, leaf_genome):
    """list per node of genome indices below it"""
    ch = children_of(parent)
    out = [None] * len(parent)
    order = np.argsort([depth(parent, v) for v in range(len(parent))])[::-1]
    for v in order:
        if not ch[v]:
            out[v] = [int(leaf_genome[v])] if leaf_genome[v] >= 0 else []
        else:
            out[v] = [g for c in ch[v] for g in out[c]]
    return out


def depth(parent, v):
    d = 0
    while parent[v] >= 0:
        v = parent[v]
        d += 1
    return d


def gene_table(parent, leaf_genome, n_core: int, n_acc: int, seed: int, shuffle: bool = True):
    """Core genes (ids 0..n_core-1: origin root, all genomes) plus accessory genes (ids from
    n_core: origin a random node, carried by a random non-empty subset of the genomes below it).
    Returns gene_id, origin, carrier_off, carrier_genome in a shuffled table order."""
    rng = np.random.default_rng(seed)
    N = len(parent)
    root = int(np.where(parent < 0)[0][0])
    below = leaves_below(parent, leaf_genome)
    entries = []
    for g in range(n_core):
        entries.append((g, root, [-1]))
    for a in range(n_acc):
        v = int(rng.integers(N))
        cand = below[v]
        k = int(rng.integers(1, len(cand) + 1))
        car = list(rng.choice(cand, size=k, replace=False))
        entries.append((n_core + a, v, [int(c) for c in car]))
    if shuffle:
        rng.shuffle(entries)
    gene_id = np.array([e[0] for e in entries], dtype=np.int64)
    origin = np.array([e[1] for e in entries], dtype=np.int32)
    off = np.zeros(len(entries) + 1, dtype=np.int64)
    car = []
    for i, e in enumerate(entries):
        car.extend(e[2])
        off[i + 1] = len(car)
    return gene_id, origin, off, np.array(car, dtype=np.int32)
