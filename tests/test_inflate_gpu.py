"""GPU BGZF inflate (inflate.cu) against zlib: every DEFLATE block type, long / overlapping matches, code lengths beyond the
primary table, empty members, records straddling members, malformed input."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bgzf(payload: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, block=0xff00, memlevel=8) -> bytes:
    parts = []
    for i in range(0, max(len(payload), 1), block):
        chunk = payload[i:i + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
        comp = c.compress(chunk) + c.flush()
        bsize = 18 + len(comp) + 8
        assert bsize <= 65536
        parts.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + comp)
        parts.append(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    parts.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))      # EOF marker: an empty member
    return b"".join(parts)


def payloads():
    rng = np.random.default_rng(42)
    text = (b"chr20\t1234567\t1234817\t29.87\n" * 4000)
    rand = rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes()
    runs = b"".join(bytes([int(v)]) * int(n) for v, n in zip(rng.integers(0, 256, 3000), rng.integers(1, 600, 3000)))
    skew = rng.choice(np.arange(256, dtype=np.uint8), 300_000, p=np.array([2.0 ** -(i // 8 + 1) for i in range(256)]) / sum(2.0 ** -(i // 8 + 1) for i in range(256))).tobytes()
    period = (bytes(range(7)) * 20000) + (b"ab" * 30000) + (b"x" * 70000)
    return {"text": text, "random": rand, "runs": runs, "skewed": skew, "periodic": period, "tiny": b"A", "empty": b""}


@pytest.mark.parametrize("name", ["text", "random", "runs", "skewed", "periodic", "tiny", "empty"])
@pytest.mark.parametrize("mode", ["default6", "fast1", "best9", "stored", "fixed", "huffman_only", "rle", "small_blocks"])
def test_inflate_equals_zlib(ctx, name, mode):
    data = payloads()[name]
    kw = {"default6": dict(level=6), "fast1": dict(level=1), "best9": dict(level=9), "stored": dict(level=0, block=0xf000),
          "fixed": dict(level=6, strategy=zlib.Z_FIXED), "huffman_only": dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY),
          "rle": dict(level=6, strategy=zlib.Z_RLE), "small_blocks": dict(level=6, block=1000, memlevel=1)}[mode]
    if mode == "stored" and name == "random":
        kw["block"] = 0xf000
    z = bgzf(data, **kw)
    out, status = ctx.bgzf_inflate(z)
    assert (status == 0).all(), status[status != 0][:10]
    assert out == data


def test_inflate_synthetic_bam(ctx, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
    import glsynth
    bam = str(tmp_path / "s.bam")
    glsynth.write_bam(bam, [("chrA", 1_500_000, 1), ("chrM", 16_569, 24)], coverage=12.0)
    z = open(bam, "rb").read()
    exp = b"".join(zlib.decompress(z[p + 18:p + bs - 8], -15) for p, bs in _members(z))
    out, status = ctx.bgzf_inflate(z)
    assert (status == 0).all() and out == exp and out[:4] == b"BAM\x01"


def _members(z):
    p = 0
    while p + 18 <= len(z):
        bs = struct.unpack_from("<H", z, p + 16)[0] + 1
        yield p, bs
        p += bs


def test_malformed_members_are_flagged(ctx):
    data = payloads()["text"]
    z = bytearray(bgzf(data))
    good, status = ctx.bgzf_inflate(bytes(z))
    assert (status == 0).all()
    # corrupt the deflate payload of the second member: either it decodes to the wrong bytes count / an invalid code (flagged)
    mem = list(_members(bytes(z)))
    p, bs = mem[1]
    z[p + 30:p + 60] = bytes(30)
    out, status = ctx.bgzf_inflate(bytes(z))
    assert status[0] == 0 and status[1] != 0 and (status[2:] == 0).all()
    # a member whose magic is wrong
    z2 = bytearray(bgzf(data))
    z2[0] = 0
    _, status = ctx.bgzf_inflate(bytes(z2))
    assert status[0] == 1


def test_gpu_feeder_equals_generator_and_host_feeder(ctx, tmp_path):
    """gl_bam_decode_device: inflate + parse + filter + CIGAR walk on the device give exactly the generator's segments in BAM
    order, for every reference; the depth text from them equals the host-feeder route"""
    from goleft_b200 import capi
    sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
    import glsynth
    contigs = [("chrA", 2_500_000, 1), ("chrB", 300_000, 2), ("chrM", 16_569, 24)]
    bam = str(tmp_path / "g.bam")
    glsynth.write_bam(bam, contigs, coverage=15.0)
    b = capi.Bam(bam)
    for tid, (nm, L, idx) in enumerate(contigs):
        s, e = glsynth.segments(L, idx, coverage=15.0)
        d = capi.bam_decode_device(ctx, b, tid)
        assert d["n"] == s.size and np.array_equal(d["start"], s) and np.array_equal(d["end"], e), nm
        host = b.decode(tid, want=32)
        assert d["n_records"] == host["n_records"] and d["n_pass"] == host["n_pass"]
    # MAPQ threshold and a BAM made by the Python writer (other record layouts: mates, unplaced reads at the end)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bamutil import make_bam_indexed
    rng = np.random.default_rng(2)
    refs = [("chrM", 16571), ("chr22", 300001)]
    recs = []
    for t, (nm, L) in enumerate(refs):
        for p0 in np.sort(rng.integers(0, L - 200, 40000)):
            cig = [[(100, "M")], [(45, "M"), (7, "D"), (55, "M")], [(30, "M"), (2, "I"), (68, "M")], [(12, "S"), (88, "M")], [(20, "M"), (300, "N"), (80, "M")],
                   [(50, "="), (1, "X"), (49, "=")]][int(rng.integers(0, 6))]
            recs.append((t, int(p0), int(rng.choice([0, 5, 30, 60])), int(rng.choice([0, 16, 0x400, 0x100, 0x200, 0x800, 4])), cig))
    recs.append((-1, -1, 0, 4, [(100, "M")]))
    bam_b, bai_b = make_bam_indexed(refs, recs)
    p = tmp_path / "p.bam"
    p.write_bytes(bam_b); (tmp_path / "p.bam.bai").write_bytes(bai_b)
    b2 = capi.Bam(str(p))
    for q in (1, 20):
        old = capi.bam_segments(str(p), q, 2)
        for t in range(2):
            d = capi.bam_decode_device(ctx, b2, t, min_mapq=q)
            es, ee = old["segments"].get(t, (np.zeros(0, np.int32), np.zeros(0, np.int32)))
            assert np.array_equal(d["start"], es) and np.array_equal(d["end"], ee), (q, t)
    b.close(); b2.close()
