"""tiny driver for ncu captures: one device-resident batch, a few kernel-only runs."""
import sys
sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "loopnest16x24p3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dom, ctx = synth.generate(name, n)
db = api.DeviceBatch(dom, ctx, -1)
for _ in range(reps):
    ms = db.run(False)
s = api.last_stats()
print("n=%d dev_ms %.2f pivots %d launches %d" % (n, ms, s.pivots, s.launches))
db.close()
