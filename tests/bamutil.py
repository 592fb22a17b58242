"""Test helpers: write small BAM / BAI files from Python (BAM spec)."""
import struct
import zlib


def _bgzf(payload: bytes) -> bytes:
    parts = []
    for i in range(0, len(payload), 60000):
        chunk = payload[i:i + 60000]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 18 + len(comp) + 8
        parts.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + comp)
        parts.append(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    parts.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(parts)


def make_bam(refs, records, sample="sampleA", mates=None) -> bytes:
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
        parts.append(struct.pack("<i", len(body)) + body)
    return _bgzf(b"".join(parts))




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
