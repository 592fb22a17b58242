"""Test helpers: write small BAM / BAI files from Python (BAM spec)."""
import struct
import zlib


BLOCK = 60000


def _bgzf(payload: bytes, offsets=None) -> bytes:
    """offsets (a list) receives the compressed offset of every block"""
    parts = []
    at = 0
    for i in range(0, len(payload), BLOCK):
        if offsets is not None:
            offsets.append(at)
        chunk = payload[i:i + 60000]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 18 + len(comp) + 8
        parts.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + comp)
        parts.append(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        at += bsize
    if offsets is not None:
        offsets.append(at)
    parts.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(parts)


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def make_bam_indexed(refs, records, sample="sampleA", mates=None):
    """(bam bytes, bai bytes): a real index — bins with merged chunks, back-filled 16 KB linear index, the 37450 stats bin
    (SAM spec 5.2) — for records sorted by (tid, pos); unplaced records (tid -1) last."""
    rec_off = []
    bam_payload = make_bam(refs, records, sample, mates, _raw=True, _rec_off=rec_off)
    blocks = []
    bam = _bgzf(bam_payload, blocks)

    def voff(o):
        return (blocks[o // BLOCK] << 16) | (o % BLOCK)
    ops_ref = "MDN=X"
    bai = b"BAI\x01" + struct.pack("<i", len(refs))
    n_no_coor = 0
    per_ref = [dict(bins={}, lin={}, beg=None, end=0, mapped=0, unmapped=0, last_bin=None) for _ in refs]
    for (tid, pos, mapq, flag, cigar), (o0, o1) in zip(records, rec_off):
        if tid < 0:
            n_no_coor += 1
            continue
        span = sum(ln for ln, op in cigar if op in ops_ref)
        end = pos + 1 if flag & 4 else pos + max(span, 1)
        R = per_ref[tid]
        b = _reg2bin(pos, end)
        v0, v1 = voff(o0), voff(o1)
        if R["beg"] is None:
            R["beg"] = v0
        R["end"] = v1
        R["unmapped" if flag & 4 else "mapped"] += 1
        ch = R["bins"].setdefault(b, [])
        if R["last_bin"] == b and ch:
            ch[-1][1] = v1
        else:
            ch.append([v0, v1])
        R["last_bin"] = b
        for k in range(pos >> 14, ((end - 1) >> 14) + 1):
            R["lin"].setdefault(k, v0)
    for R in per_ref:
        n_bins = len(R["bins"]) + (1 if R["beg"] is not None else 0)
        bai += struct.pack("<i", n_bins)
        for b in sorted(R["bins"]):
            bai += struct.pack("<Ii", b, len(R["bins"][b]))
            for c0, c1 in R["bins"][b]:
                bai += struct.pack("<QQ", c0, c1)
        if R["beg"] is not None:
            bai += struct.pack("<Ii", 37450, 2) + struct.pack("<4Q", R["beg"], R["end"], R["mapped"], R["unmapped"])
        n_lin = max(R["lin"]) + 1 if R["lin"] else 0
        lin = [R["lin"].get(k, 0) for k in range(n_lin)]
        for k in range(n_lin - 2, -1, -1):
            if lin[k] == 0:
                lin[k] = lin[k + 1]
        bai += struct.pack("<i", n_lin) + struct.pack("<%dQ" % n_lin, *lin)
    return bam, bai + struct.pack("<Q", n_no_coor)


def make_bam(refs, records, sample="sampleA", mates=None, _raw=False, _rec_off=None) -> bytes:
    text = b"@HD\tVN:1.6\tSO:coordinate\n@RG\tID:a\tSM:" + sample.encode() + b"\n"
    parts = [b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))]
    for n, l in refs:
        parts.append(struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l))
    ops = "MIDNSHP=X"
    for ri, (tid, pos, mapq, flag, cigar) in enumerate(records):
        name = b"r\0"
        mate_pos, tlen = mates[ri] if mates is not None else (-1, 0)
        cig = b"".join(struct.pack("<I", (ln << 4) | ops.index(op)) for ln, op in cigar)
        lseq = sum(ln for ln, op in cigar if op in "MIS=X")
        body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, 0, len(cigar), flag, lseq, tid if mates is not None else -1, mate_pos, tlen)
        body += name + cig + b"\0" * ((lseq + 1) // 2) + b"\xff" * lseq
        if _rec_off is not None:
            at = sum(len(x) for x in parts) if not _rec_off else _rec_off[-1][1]
            _rec_off.append((at, at + 4 + len(body)))
        parts.append(struct.pack("<i", len(body)) + body)
    return b"".join(parts) if _raw else _bgzf(b"".join(parts))




def make_bai(linear, stats=None) -> bytes:
    """linear: per reference list of u64 virtual offsets; stats: per reference (mapped, unmapped) or None"""
    b = b"BAI\x01" + struct.pack("<i", len(linear))
    for r, iv in enumerate(linear):
        if stats is not None and stats[r] is not None:
            b += struct.pack("<i", 1) + struct.pack("<Ii", 37450, 2) + struct.pack("<4Q", 0, 0, stats[r][0], stats[r][1])
        else:
            b += struct.pack("<i", 0)
        b += struct.pack("<i", len(iv)) + struct.pack("<%dQ" % len(iv), *iv)
    return b + struct.pack("<Q", 0)
