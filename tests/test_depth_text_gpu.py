"""GPU parity for the device BED formatter (depth_text.cu) and the one-call contig entry:
the bytes must equal the oracle's line-by-line walker (oracle_depth.c restating depth/depth.go:238-364)
run chunk by chunk like `goleft depth` in .fai mode, and the %.4g tokens must equal printf's."""
import numpy as np
import pytest

from goleft_b200 import capi, synth
from oracle import loader as orc

pytestmark = pytest.mark.gpu


def oracle_contig_text(chrom, L, s, e, W, mincov, maxmean):
    """what `goleft depth` writes for one reference in .fai mode: the callback's rows for every chunk, in order"""
    hd, ca = [], []
    for cs, ce in orc.gen_chunks(L, W):
        d = orc.pileup_diff(s, e, cs, ce)
        h, c = orc.walk_chunk(chrom, cs, ce, W, mincov, maxmean, d)
        hd.append(h)
        ca.append(c)
    return b"".join(hd), b"".join(ca)


def step_for(W):
    return max(1, 10_000_000 // W) * W            # depth/depth.go:132


@pytest.mark.parametrize("W", [500, 250, 100, 13, 2001, 1_000_000, 3])
def test_contig_text_equals_walker(ctx, W):
    L = 1_234_567
    s, e = synth.segments(synth.reads(L, contig_index=5))
    exp = oracle_contig_text("chr5", L, s, e, W, 4, 0)
    got = ctx.depth_bed_contig("chr5", L, s, e, W, 4, 0, step_for(W))
    assert got[0] == exp[0]
    assert got[1] == exp[1]
    a, d, ln = capi.pack_segments8(s, e, threads=0)
    got8 = ctx.depth_bed_contig_packed8("chr5", L, a, d, ln, W, 4, 0, step_for(W))
    assert got8 == exp


def test_contig_text_multi_chunk_low_coverage(ctx):
    # 2.5 chunks of 10 Mb at 3x: thousands of class runs, forced run breaks at the chunk edges, maxmeandepth on
    L = 25_000_123
    s, e = synth.segments(synth.reads(L, coverage=3.0, contig_index=7))
    for W, mincov, maxmean in ((250, 4, 0), (500, 2, 6)):
        exp = oracle_contig_text("7", L, s, e, W, mincov, maxmean)
        got = ctx.depth_bed_contig("7", L, s, e, W, mincov, maxmean, step_for(W))
        assert got[0] == exp[0]
        assert got[1] == exp[1]


def test_contig_text_empty_and_tiny(ctx):
    z = np.zeros(0, np.int32)
    for L in (1, 499, 500, 501, 4096, 10_000_001):
        exp = oracle_contig_text("chrM", L, z, z, 500, 4, 0)
        assert ctx.depth_bed_contig("chrM", L, z, z, 500, 4, 0, 10_000_000) == exp
    s = np.array([10, 20, 700], np.int32)
    e = np.array([30, 25, 701], np.int32)
    exp = oracle_contig_text("x", 1000, s, e, 100, 1, 0)
    assert ctx.depth_bed_contig("x", 1000, s, e, 100, 1, 0, 10_000_000) == exp


def test_long_chrom_name_and_long_reads(ctx):
    rng = np.random.default_rng(3)
    L = 300_000
    s = np.sort(rng.integers(0, L - 30_000, 400)).astype(np.int32)
    e = (s + rng.integers(8_000, 30_000, 400)).astype(np.int32)       # long reads -> int32 path inside the call
    name = "HLA-DRB1*15:03:01:02_some_very_long_decoy_contig_name_0123456789"[:64]
    exp = oracle_contig_text(name, L, s, e, 250, 4, 0)
    assert ctx.depth_bed_contig(name, L, s, e, 250, 4, 0, 10_000_000) == exp
    assert ctx.depth_last_path() == 4                                    # segments beyond the fused path's 16384-base look-back: bucketed events
    with pytest.raises(capi.GlError):
        ctx.depth_bed_contig(name + "x", L, s, e, 250, 4, 0, 10_000_000)


def test_g4_tokens_equal_printf(ctx):
    """explicit rows: every kind of %.4g outcome, incl. exact ties (sum/len = x.xxx5), decade roll-over, %e forms"""
    rng = np.random.default_rng(11)
    sums, lens = [], []
    # ties and near-ties: means k/16, k/32, k/2^j have finite binary expansions that end in ...5
    for ln in (2, 4, 8, 16, 32, 64, 125, 250, 500, 625, 1000, 2000, 3, 7, 13):
        for k in list(range(0, 300)) + [9999, 99995, 999949, 999950, 999951, 12345678, 2**31 - 1, 99999, 100000, 100001]:
            sums.append(k)
            lens.append(ln)
    big = rng.integers(0, 2**40, 20000)
    sums += list(big)
    lens += list(rng.integers(1, 5000, big.size))
    small = rng.integers(0, 5, 5000)
    sums += list(small)
    lens += list(rng.integers(1, 2**30, small.size))
    # 9999.5 boundary in several decades
    for ln in (2, 20, 200, 2000, 20000):
        for k in (19998, 19999, 20000, 20001, 199989, 199990, 199991):
            sums.append(k)
            lens.append(ln)
    sums = np.array(sums, np.int64)
    lens = np.array(lens, np.int64)
    row_s = np.arange(sums.size, dtype=np.int64) % 1000
    row_e = row_s + lens
    ok = row_e < 2**31 - 1
    sums, lens, row_s, row_e = sums[ok], lens[ok], row_s[ok], row_e[ok]
    got = ctx.depth_format_rows("c", row_s, row_e, sums).decode().splitlines()
    assert len(got) == sums.size
    for i in range(sums.size):
        mean = 0.0 if sums[i] == 0 else float(sums[i]) / float(lens[i])
        assert got[i] == "c\t%d\t%d\t%s" % (row_s[i], row_e[i], "%.4g" % mean), (i, sums[i], lens[i])


@pytest.fixture
def packed_transport(monkeypatch):
    """the one-call int32 entry with its fixed-block packed16 transport forced on (auto needs >= 2^20 segments)"""
    monkeypatch.setenv("GL_BED_PACK", "16")
    yield
    monkeypatch.delenv("GL_BED_PACK", raising=False)


def test_packed16_transport_same_bytes(ctx, packed_transport):
    # 5.3 Mb at 30x: a centromere gap (one block straddles it -> escape list), a 200x pile-up, deletion reads whose second
    # block is out of order, a segment count that is not a multiple of 256, several upload chunks
    L = 5_300_017
    s, e = synth.segments(synth.reads(L, contig_index=9))
    assert s.size % 256 != 0 and s.size > 4096
    a, o, ln, es, ee = capi.pack_segments16_fixed(s, e)
    assert es.size > 0                                            # the gap block escaped
    for W, mincov, maxmean in ((500, 4, 0), (250, 4, 60)):
        exp = oracle_contig_text("chr9", L, s, e, W, mincov, maxmean)
        before = ctx.launch_count()
        got = ctx.depth_bed_contig("chr9", L, s, e, W, mincov, maxmean, step_for(W))
        assert got[0] == exp[0] and got[1] == exp[1]
        assert ctx.depth_last_path() == 1                         # unpacked into the store, then the fused int32 path
        assert ctx.launch_count() - before >= 6                   # + depth_unpack16_kernel


def test_packed16_transport_empty_unsorted_and_long_segments(ctx, packed_transport):
    rng = np.random.default_rng(5)
    L = 2_000_000
    # unsorted + empty/reversed segments + negative starts and ends beyond the region
    s = rng.integers(-200, L + 100, 300_000).astype(np.int32)
    e = (s + rng.integers(-5, 160, s.size)).astype(np.int32)
    exp = oracle_contig_text("u", L, s, e, 500, 4, 0)
    assert ctx.depth_bed_contig("u", L, s, e, 500, 4, 0, 10_000_000) == exp
    # long reads: every block escapes (segments > 65535 or starts too far apart) -> the call falls back to plain int32
    s = np.sort(rng.integers(0, L - 100_000, 5000)).astype(np.int32)
    e = (s + rng.integers(60_000, 100_000, s.size)).astype(np.int32)
    exp = oracle_contig_text("lr", L, s, e, 500, 4, 0)
    assert ctx.depth_bed_contig("lr", L, s, e, 500, 4, 0, 10_000_000) == exp
    # sparse: 20 kb reads 40 kb apart -> most blocks escape but the list fits
    s = (np.arange(6000, dtype=np.int64) * 300).astype(np.int32)
    e = (s + 150).astype(np.int32)
    s[::512] += 70_000                                            # one outlier per second block
    e[::512] += 70_000
    exp = oracle_contig_text("sp", L, s, e, 250, 1, 0)
    assert ctx.depth_bed_contig("sp", L, s, e, 250, 1, 0, 10_000_000) == exp
