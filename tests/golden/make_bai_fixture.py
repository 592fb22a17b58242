"""Regenerates tests/golden/sample_issue_27_bai_linear_index.json from the reference's in-tree fixture
indexcov/test-data/sample_issue_27_0001.bam.bai (run in the build container; /root/reference does not exist
on the GPU box).  Only the linear index (n_intv x u64 per reference) and the stats-bin totals are kept."""
import json
import struct
import sys


def parse_bai(path):
    b = open(path, "rb").read()
    assert b[:4] == b"BAI\x01"
    off = 4
    (n_ref,) = struct.unpack_from("<i", b, off); off += 4
    refs, mapped, unmapped = [], 0, 0
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", b, off); off += 4
        for _ in range(n_bin):
            bn, nch = struct.unpack_from("<Ii", b, off); off += 8
            if bn == 37450 and nch == 2:      # StatsDummyBin 0x924a, indexcov/types.go:19
                _, _, m, u = struct.unpack_from("<4Q", b, off)
                mapped += m; unmapped += u
            off += 16 * nch
        (n_intv,) = struct.unpack_from("<i", b, off); off += 4
        refs.append(list(struct.unpack_from("<%dQ" % n_intv, b, off))); off += 8 * n_intv
    return refs, mapped, unmapped


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/indexcov/test-data/sample_issue_27_0001.bam.bai"
    refs, m, u = parse_bai(src)
    json.dump({"source": "indexcov/test-data/sample_issue_27_0001.bam.bai (reference fixture, linear index only)",
               "mapped": m, "unmapped": u, "ioffsets": refs},
              open("tests/golden/sample_issue_27_bai_linear_index.json", "w"))
