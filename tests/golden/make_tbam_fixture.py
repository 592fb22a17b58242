"""Regenerates tests/golden/t_bam_segments.npz from the reference's fixture depth/test/t.bam (run in the build
container; /root/reference does not exist on the GPU box).

Two independent decoders must agree before the fixture is written:
  1. the product's C++ feeder (gl_bam_decode_segments, multi-threaded BGZF + BAM), and
  2. a pure-Python restatement (gzip module + struct) of the BAM spec and the `samtools depth -Q 1` filter.
The fixture holds, per reference, the filtered M/=/X blocks [start,end) in record order.
"""
import gzip
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def py_segments(path, min_mapq=1):
    b = gzip.open(path, "rb").read()          # BGZF is a series of gzip members
    assert b[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", b, 4)
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", b, off); off += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", b, off)
        name = b[off + 4: off + 4 + l_name - 1].decode()
        (l_ref,) = struct.unpack_from("<i", b, off + 4 + l_name)
        refs.append((name, l_ref)); off += 8 + l_name
    S = [[] for _ in refs]; E = [[] for _ in refs]
    nrec = npass = 0
    while off < len(b):
        (bs,) = struct.unpack_from("<i", b, off)
        tid, pos, l_name, mapq, _bin, n_cig, flag = struct.unpack_from("<iiBBHHH", b, off + 4)
        nrec += 1
        if tid >= 0 and (flag & 0x704) == 0 and mapq >= min_mapq:
            npass += 1
            cig = struct.unpack_from("<%dI" % n_cig, b, off + 4 + 32 + l_name)
            ref = pos; cur = None
            for v in cig:
                op, ln = v & 15, v >> 4
                if op in (0, 7, 8):
                    if ln == 0:
                        continue
                    if cur is not None and cur[1] == ref:
                        cur[1] = ref + ln
                    else:
                        if cur is not None:
                            S[tid].append(cur[0]); E[tid].append(cur[1])
                        cur = [ref, ref + ln]
                    ref += ln
                elif op in (2, 3):
                    ref += ln
            if cur is not None:
                S[tid].append(cur[0]); E[tid].append(cur[1])
        off += 4 + bs
    return refs, S, E, nrec, npass


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/depth/test/t.bam"
    from goleft_b200 import capi
    refs, S, E, nrec, npass = py_segments(src)
    r = capi.bam_segments(src, 1, 4)
    assert r["refs"] == refs and r["n_records"] == nrec and r["n_pass"] == npass
    out = {"ref_names": np.array([n for n, _ in refs]), "ref_lens": np.array([l for _, l in refs], np.int64),
           "n_records": np.int64(nrec), "n_pass": np.int64(npass)}
    for tid in range(len(refs)):
        s, e = r["segments"].get(tid, (np.zeros(0, np.int32), np.zeros(0, np.int32)))
        assert np.array_equal(s, np.array(S[tid], np.int32)) and np.array_equal(e, np.array(E[tid], np.int32))
        out["start_%d" % tid] = s; out["end_%d" % tid] = e
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "t_bam_segments.npz"), **out)
    print("ok", refs, nrec, npass, [len(x) for x in S])
