"""The reference's own fixture depth/test/t.bam (as tests/golden/t_bam_segments.npz, see make_tbam_fixture.py)
through the oracle (CPU) and through the CUDA path (GPU), in the configurations of depth/functional-test.sh."""
import os
import sys

import numpy as np
import pytest

from goleft_b200 import capi
from oracle import loader as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
FAI_WINDOWS = [100, 1000000000, 55, 60, 71, 13, 2001]          # functional-test.sh:45-70
BED_WINDOWS = [10, 1000000, 50, 55, 60, 71, 13, 2002]          # functional-test.sh:73-97
WINDOWS_BED = [("chr22", 14250, 15500), ("chr22", 1575, 15800), ("chrM", 100, 1000), ("chrM", 2000, 5000), ("chrM", 1, 3),
               ("chrM", 9, 13), ("chrM", 16, 17), ("chrM", 24, 29), ("chrM", 39, 43)]   # depth/test/windows.bed


def load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "t_bam_segments.npz"))
    refs = list(zip([str(x) for x in z["ref_names"]], [int(x) for x in z["ref_lens"]]))
    segs = {i: (z["start_%d" % i], z["end_%d" % i]) for i in range(len(refs))}
    return refs, segs, int(z["n_records"]), int(z["n_pass"])


def rows(b):
    return [ln.split("\t") for ln in b.decode().splitlines()]


def oracle_fai(refs, segs, W):
    hd, ca = b"", b""
    for tid, (name, L) in enumerate(refs):
        s, e = segs[tid]
        depth = orc.pileup_brute(s, e, 0, L)
        for cs, ce in orc.gen_chunks(L, W):
            h, c = orc.walk_chunk(name, cs, ce, W, 4, 0, depth[cs:ce])
            hd += h; ca += c
    return hd, ca


def test_fixture_facts():
    """SURVEY.md §4 (derived there with an independent throw-away reader): record / pass counts, covered bases,
    total and maximum depth per contig at -Q 1."""
    refs, segs, nrec, npass = load()
    assert refs == [("chrM", 16571), ("chr22", 20001)] and nrec == 80330 and npass == 75808 + 246
    d = orc.pileup_brute(*segs[0], 0, 16571)
    assert (int((d > 0).sum()), int(d.sum()), int(d.max())) == (5076, 5743876, 2012)
    d = orc.pileup_brute(*segs[1], 0, 20001)
    assert (int((d > 0).sum()), int(d.sum()), int(d.max())) == (9811, 23813, 39)


@pytest.mark.parametrize("W", FAI_WINDOWS)
def test_oracle_fai_mode_invariants(W):
    refs, segs, _, _ = load()
    hd, ca = oracle_fai(refs, segs, W)
    for text in (hd, ca):                                   # check_with_fai_bt + check_uniq (functional-test.sh:10-39)
        r = rows(text)
        assert len({tuple(x) for x in r}) == len(r) and len({tuple(x[:3]) for x in r}) == len(r)
        for name, L in refs:
            rr = [x for x in r if x[0] == name]
            assert int(rr[0][1]) == 0 and int(rr[-1][2]) == L
            assert all(a[2] == b[1] for a, b in zip(rr, rr[1:]))
    if W == 100:
        assert len(rows(hd)) == 367 and len(rows(ca)) == 148
        assert [x[3] for x in rows(hd)[:8]] == ["608.2", "1101", "625.5", "457.4", "789.3", "897.3", "1583", "1801"]


def test_oracle_w250_row_counts():
    refs, segs, _, _ = load()
    hd, ca = oracle_fai(refs, segs, 250)                    # the depth default window (depth.go:164)
    assert len(rows(hd)) == 148 and len(rows(ca)) == 148


# ------------------------------------------------------------------ the feeder on a BAM written here
from bamutil import make_bam as _bam


def test_feeder_on_synthetic_bam(tmp_path):
    rng = np.random.default_rng(0)
    refs = [("chrA", 100000), ("HLA-A*01:01:01:01", 5000)]
    recs, exp = [], {0: ([], []), 1: ([], [])}
    pos = 0
    for i in range(30000):
        tid = 0 if i < 25000 else 1
        if i == 25000:
            pos = 0
        pos += int(rng.integers(0, 4)) if tid == 0 else int(rng.integers(0, 2)) % 2
        flag = int(rng.choice([0, 16, 0x400, 0x100, 0x200, 0x4, 0x800, 99]))
        mapq = int(rng.choice([0, 1, 30, 60]))
        kind = int(rng.integers(0, 6))
        cigar = [[(100, "M")], [(40, "M"), (5, "D"), (60, "M")], [(30, "M"), (3, "I"), (67, "M")], [(10, "S"), (90, "M")],
                 [(50, "="), (1, "X"), (20, "N"), (49, "M")], [(5, "H"), (20, "M"), (2, "P"), (30, "M"), (1000, "N"), (50, "M")]][kind]
        recs.append((tid if not (flag & 4) else -1, pos, mapq, flag, cigar))
        if not (flag & 4) and (flag & 0x704) == 0 and mapq >= 1:
            ref, cur = pos, None
            for ln, op in cigar:
                if op in "M=X":
                    if cur and cur[1] == ref:
                        cur[1] = ref + ln
                    else:
                        if cur:
                            exp[tid][0].append(cur[0]); exp[tid][1].append(cur[1])
                        cur = [ref, ref + ln]
                    ref += ln
                elif op in "DN":
                    ref += ln
            exp[tid][0].append(cur[0]); exp[tid][1].append(cur[1])
    p = tmp_path / "x.bam"
    p.write_bytes(_bam(refs, recs))
    for threads in (1, 4):
        r = capi.bam_segments(str(p), 1, threads)
        assert r["refs"] == refs and r["n_records"] == len(recs)
        for tid in (0, 1):
            assert r["segments"][tid][0].tolist() == exp[tid][0] and r["segments"][tid][1].tolist() == exp[tid][1]
    only = capi.bam_segments(str(p), 1, 2, only_tid=1)
    assert 0 not in only["segments"] and only["segments"][1][0].tolist() == exp[1][0]
    with pytest.raises(capi.GlError):
        capi.bam_segments(str(tmp_path / "missing.bam"))


# ------------------------------------------------------------------ GPU: the fixture through the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("W", FAI_WINDOWS)
def test_gpu_fai_mode_text(ctx, W):
    refs, segs, _, _ = load()
    exp_hd, exp_ca = oracle_fai(refs, segs, W)
    hd, ca = b"", b""
    for tid, (name, L) in enumerate(refs):
        s, e = segs[tid]
        chunks = orc.gen_chunks(L, W)
        step = chunks[0][1] - chunks[0][0] if len(chunks) > 1 else 0
        ws, r0, rc = ctx.depth_region(0, L, s, e, W, 4, 0, run_break=step)
        for cs, ce in chunks:
            lo, hi = np.searchsorted(r0, cs), np.searchsorted(r0, ce)
            h, c = capi.format_chunk(name, cs, ce, W, ws[cs // W:(ce - 1) // W + 1], r0[lo:hi], rc[lo:hi])
            hd += h; ca += c
    assert hd == exp_hd and ca == exp_ca


@pytest.mark.gpu
@pytest.mark.parametrize("W", BED_WINDOWS)
def test_gpu_bed_mode_text(ctx, W):
    refs, segs, _, _ = load()
    tid_of = {n: i for i, (n, _) in enumerate(refs)}
    for name, rs, re in WINDOWS_BED:
        s, e = segs[tid_of[name]]
        depth = orc.pileup_brute(s, e, rs, re)
        exp = orc.walk_chunk(name, rs, re, W, 4, 0, depth)
        ws, r0, rc = ctx.depth_region(rs, re, s, e, W, 4, 0)
        assert capi.format_chunk(name, rs, re, W, ws, r0, rc) == exp


def test_indexed_feeder_equals_stream_decoder(tmp_path):
    """bam_feed (index-guided, parallel units, sorted emission) yields the same M/=/X blocks as the streaming decoder, as
    packed8 and as sorted int32, on a BGZF-realistic synthetic BAM; region queries return every block that can overlap"""
    import os, sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
    import glsynth
    from goleft_b200 import capi
    from oracle import loader as orc
    contigs = [("chrA", 3_000_000, 1), ("chrB", 700_000, 2), ("chrM", 16_569, 24)]
    bam = str(tmp_path / "t.bam")
    glsynth.write_bam(bam, contigs, coverage=20.0)
    old = capi.bam_segments(bam, 1, 4)
    b = capi.Bam(bam)
    assert b.has_index and [r[0] for r in b.refs] == [c[0] for c in contigs]
    for tid, (nm, L, idx) in enumerate(contigs):
        s, e = glsynth.segments(L, idx, coverage=20.0)
        gs, ge = old["segments"][tid]
        assert np.array_equal(gs, s) and np.array_equal(ge, e)                  # the stream decoder's filter == the generator's
        o = np.lexsort((e, s))
        for threads in (1, 3, 0):
            d = b.decode(tid, threads=threads)
            assert d["format"] == 8 and (np.diff(d["anchors"]) >= 0).all()
            us, ue = capi.unpack_segments8(d["anchors"], d["dstart"], d["len"])
            o2 = np.lexsort((ue, us))
            assert np.array_equal(us[o2], s[o]) and np.array_equal(ue[o2], e[o])
            d = b.decode(tid, threads=threads, want=32)
            assert (np.diff(d["start"]) >= 0).all()
            o3 = np.lexsort((d["end"], d["start"]))
            assert np.array_equal(d["start"][o3], s[o]) and np.array_equal(d["end"][o3], e[o])
    s, e = glsynth.segments(3_000_000, 1, coverage=20.0)
    for rs, re in ((0, 1), (1_000_000, 1_000_001), (1_234_567, 1_300_000), (2_990_000, 3_000_000), (16_383, 16_385)):
        d = b.decode(0, rs, re, want=32)
        got = orc.pileup_diff(d["start"], d["end"], rs, re) if d["n"] else np.zeros(re - rs, np.int32)
        assert np.array_equal(got, orc.pileup_diff(s, e, rs, re))
        assert d["n_records"] < 400_000 * (re - rs + 40_000) / 3_000_000 + 20_000     # only the region's blocks were parsed
    b.close()
