"""per-round timing of the size-class ladder for one workload (PIPLIB_B200_TIMING=1)"""
import os, sys
os.environ["PIPLIB_B200_TIMING"] = "1"
sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth
name, n = sys.argv[1], int(sys.argv[2])
dom, ctx = synth.generate(name, n)
db = api.DeviceBatch(dom, ctx, synth.bignum(name), **synth.options(name))
for i in range(2):
    ms = db.run(False)
    s = api.last_stats()
    print("%s n=%d: %.1f ms, %d pivots, %.1f Mpivots/s, max_rows %d" % (name, n, ms, s.pivots, s.pivots / ms / 1e3, s.max_rows), flush=True)
db.close()
