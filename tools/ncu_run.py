"""tiny driver for ncu captures: one device-resident batch, a few kernel-only runs."""
import sys
sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "loopnest16x24p3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dom, ctx = synth.generate(name, n)
db = api.DeviceBatch(dom, ctx, synth.bignum(name), **synth.options(name))
best = 1e30
for _ in range(reps):
    ms = db.run(False)
    best = min(best, ms)
s = api.last_stats()
print("n=%d dev_ms last %.2f best %.2f pivots %d launches %d rounds %d -> %.2f M problems/s" % (n, ms, best, s.pivots, s.launches, s.rounds, n / best / 1e3))
db.close()
