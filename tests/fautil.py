"""Tiny FASTA + .fai writer for the --stats tests."""
import numpy as np


def make_fasta(refs, seed=0, line=60, crlf=False):
    """refs = [(name, length)] -> (fasta bytes, fai text, {name: (offset, length, line_bases, line_width)})"""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnCGcgRy", np.uint8)
    out = bytearray()
    fai, recs = "", {}
    nl = b"\r\n" if crlf else b"\n"
    for name, L in refs:
        out += b">" + name.encode() + b" synthetic" + nl
        off = len(out)
        seq = alphabet[rng.integers(0, alphabet.size, L)]
        # stretches of masked / N / CpG-rich sequence so that every column varies
        for _ in range(max(1, L // 3000)):
            a = int(rng.integers(0, max(1, L - 200)))
            kind = int(rng.integers(0, 3))
            n = int(rng.integers(20, 200))
            if kind == 0:
                seq[a:a + n] = np.frombuffer(b"N", np.uint8)[0]
            elif kind == 1:
                seq[a:a + n] |= 0x20
            else:
                seq[a:a + n:2] = ord("C"); seq[a + 1:a + n:2] = ord("G")
        b = seq.tobytes()
        for i in range(0, L, line):
            out += b[i:i + line] + nl
        fai += "%s\t%d\t%d\t%d\t%d\n" % (name, L, off, line, line + len(nl))
        recs[name] = (off, L, line, line + len(nl))
    return bytes(out), fai, recs


def position(rec, p):
    off, L, lb, lw = rec
    return p // lb * lw + p % lb          # relative to the record's first base
