"""Regenerates tests/golden/viral_crai_slices.npz from the reference fixture indexcov/test-data/viral.crai
(run in the build container).  Keeps (seqID, alnStart, alnSpan, sliceLen) of every mapped slice."""
import gzip
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/indexcov/test-data/viral.crai"
rows = []
for ln in gzip.open(src, "rt"):
    t = ln.split("\t")
    if int(t[0]) >= 0:
        rows.append((int(t[0]), int(t[1]), int(t[2]), int(t[5])))
a = np.array(rows, np.int64)
np.savez_compressed("tests/golden/viral_crai_slices.npz", seq=a[:, 0].astype(np.int32), start=a[:, 1], span=a[:, 2],
                    nbytes=a[:, 3].astype(np.int32))
print(len(rows))
