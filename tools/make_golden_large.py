"""Golden vector for BASELINE config 4 at its full size: the UNMODIFIED reference (oracle/_ref/
libpipref_big.so: same sources, SOL_SIZE / MAXCOL raised with -D, oracle/Makefile `refbig`) solves the
4096 x 4097 consecutive-ones tableau of bench.py (seed 2026); the cells go to
tests/golden/large_consecutive_ones_4096.json together with the pivot count of the oracle port (the
reference has no pivot counter), which must produce the same cells.  ~1 minute of CPU time."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from workloads import synth  # noqa: E402

n, seed = 4096, 2026
po.build(ref=True, port=True)
tab = synth.consecutive_ones(n, n, seed=seed)
st_r, cells_r = po.Ref(big=True).traiter(n, 0, n, 0, -1, 1, tab, [], cap=1 << 20)
stats = po.PortStats()
st_p, cells_p = po.Port().traiter(n, 0, n, 0, -1, 1, tab, [], sol_size=1 << 20, maxcol=1 << 16, stats=stats)
assert (st_r, cells_r) == (st_p, cells_p), "the port disagrees with the reference"
out = dict(n=n, seed=seed, nq=1, status=st_r, pivots=int(stats.pivots), cuts=int(stats.cuts_const),
           source="oracle/_ref/libpipref_big.so (unmodified reference, SOL_SIZE=1048576 MAXCOL=65536)",
           cells=[list(c) for c in cells_r])
with open(os.path.join(ROOT, "tests", "golden", "large_consecutive_ones_4096.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("status", st_r, "cells", len(cells_r), "pivots", stats.pivots)
