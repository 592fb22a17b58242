"""Seeded synthetic alignment streams shaped like a 30x short-read WGS BAM.

Used by tests/ and bench.py to make inputs; it is a workload generator, not part of the
depth engine.  `reads()` produces what a BAM decoder would hand to the feeder (pos, flag,
mapq, cigar shape); `segments()` applies the reference's filter (flag & 0x704 == 0, MAPQ >= Q,
the defaults of the `samtools depth -Q` child at depth/depth.go:45) and expands CIGARs into
the M/=/X blocks [start,end) that the GPU path consumes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

SEED0 = 0x601EF7

CHR20_LEN = 64_444_167
GRCH38 = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
    ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
    ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
    ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415), ("chrM", 16569),
]

FLAG_UNMAP, FLAG_SECONDARY, FLAG_QCFAIL, FLAG_DUP, FLAG_SUPP = 0x4, 0x100, 0x200, 0x400, 0x800
FILTER_MASK = 0x704  # UNMAP | SECONDARY | QCFAIL | DUP (samtools depth default)


@dataclass
class Reads:
    pos: np.ndarray      # int32 0-based leftmost reference position
    flag: np.ndarray     # uint16
    mapq: np.ndarray     # uint8
    kind: np.ndarray     # uint8: 0 = 150M, 1 = kM dD (150-k)M, 2 = kM iI (150-k-i)M, 3 = sS (150-s)M
    k: np.ndarray        # int32 first-block length (kinds 1,2) or soft-clip length s (kind 3)
    x: np.ndarray        # int32 d (kind 1) or i (kind 2)
    read_len: int


def reads(length: int, coverage: float = 30.0, read_len: int = 150, contig_index: int = 0,
          gap: bool = True, pileup: bool = True) -> Reads:
    rng = np.random.Generator(np.random.PCG64(SEED0 + contig_index))
    n = int(coverage * length // read_len)
    hi = max(1, length - read_len)
    pos = rng.integers(0, hi, size=n, dtype=np.int64)
    if pileup and length > 50_000:
        p0 = int(0.70 * length)
        plen = min(10_000, length // 100)
        extra = int((200 - coverage) * plen // read_len)
        pos = np.concatenate([pos, rng.integers(max(0, p0 - read_len + 1), p0 + plen, size=extra, dtype=np.int64)])
    if gap and length > 50_000:
        g0, g1 = int(0.40 * length), int(0.40 * length) + max(1, int(0.0466 * length))  # 3 Mb of chr20
        keep = (pos + read_len + 16 <= g0) | (pos >= g1)
        pos = pos[keep]
    pos.sort(kind="stable")
    n = pos.size
    u = rng.random(n)
    kind = np.zeros(n, np.uint8)
    kind[u >= 0.97] = 1
    kind[u >= 0.98] = 2
    kind[u >= 0.99] = 3
    k = rng.integers(1, read_len - 12, size=n, dtype=np.int32)
    x = rng.integers(1, 11, size=n, dtype=np.int32)
    s = rng.integers(1, 51, size=n, dtype=np.int32)
    k = np.where(kind == 3, s, k).astype(np.int32)
    um = rng.random(n)
    mapq = np.full(n, 60, np.uint8)
    mid = (um >= 0.05) & (um < 0.10)
    mapq[um < 0.05] = 0
    mapq[mid] = rng.integers(1, 60, size=int(mid.sum()), dtype=np.uint8)
    uf = rng.random(n)
    flag = np.zeros(n, np.uint16)
    edges = np.cumsum([0.06, 0.002, 0.005, 0.003, 0.001])
    flag[uf < edges[0]] = FLAG_DUP
    flag[(uf >= edges[0]) & (uf < edges[1])] = FLAG_QCFAIL
    flag[(uf >= edges[1]) & (uf < edges[2])] = FLAG_SECONDARY
    flag[(uf >= edges[2]) & (uf < edges[3])] = FLAG_SUPP
    flag[(uf >= edges[3]) & (uf < edges[4])] = FLAG_UNMAP
    return Reads(pos.astype(np.int32), flag, mapq, kind, k, x, read_len)


def segments(r: Reads, min_mapq: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Filtered M-block intervals, in BAM record order (a deletion read yields two consecutive blocks)."""
    ok = ((r.flag & FILTER_MASK) == 0) & (r.mapq >= min_mapq)
    pos, kind, k, x = r.pos[ok].astype(np.int64), r.kind[ok], r.k[ok].astype(np.int64), r.x[ok].astype(np.int64)
    L = r.read_len
    n = pos.size
    # first block
    s1 = pos.copy()
    e1 = pos + L
    is_del, is_ins, is_clip = kind == 1, kind == 2, kind == 3
    e1[is_del] = pos[is_del] + k[is_del]
    e1[is_ins] = pos[is_ins] + (L - x[is_ins])           # kM iI (L-k-i)M: the two M blocks abut on the reference
    e1[is_clip] = pos[is_clip] + (L - k[is_clip])        # sS (L-s)M
    # second block of deletion reads
    s2 = pos[is_del] + k[is_del] + x[is_del]
    e2 = pos[is_del] + L + x[is_del]
    # interleave: output index of read i's first block = i + (#deletion reads before i)
    off = np.cumsum(is_del) - is_del
    idx1 = np.arange(n) + off
    total = n + int(is_del.sum())
    start = np.empty(total, np.int32)
    end = np.empty(total, np.int32)
    start[idx1] = s1
    end[idx1] = e1
    idx2 = idx1[is_del] + 1
    start[idx2] = s2
    end[idx2] = e2
    return start, end


def chr20_like(length: int = CHR20_LEN, contig_index: int = 19, min_mapq: int = 1):
    return segments(reads(length, contig_index=contig_index), min_mapq)
