"""run selected problems of a workload (indices on the command line) and print the phase split of the
-DPIP_PROFILE build:  PIPLIB_B200_LIB=piplib_b200/lib/libpiplib_dp_prof.so python tools/one.py vivien32 133"""
import os, sys
os.environ.setdefault("PIPLIB_B200_TIMING", "1")
sys.path.insert(0, ".")
import numpy as np
from piplib_b200 import api  # noqa: E402
from workloads import synth
name = sys.argv[1]
idx = [int(x) for x in sys.argv[2:]]
dom, ctx = synth.generate(name, max(idx) + 1)
dom, ctx = np.ascontiguousarray(dom[idx]), np.ascontiguousarray(ctx[idx])
db = api.DeviceBatch(dom, ctx, synth.bignum(name), **synth.options(name))
for i in range(2):
    ms = db.run(False)
    s = api.last_stats()
    print("%s %s: %.1f ms, %d pivots, %.1f us/pivot, max_rows %d" % (name, idx, ms, s.pivots, 1e3 * ms / max(1, s.pivots), s.max_rows), flush=True)
tot = sum(s.phase_cycles[:len(api.PHASES)])
if tot:
    print("phase split (leader warp cycles): " + ", ".join("%s %.1f%%" % (p, 100.0 * c / tot) for p, c in zip(api.PHASES, s.phase_cycles)))
    print("warp-cycles per pivot: %.0f" % (tot / max(1, s.pivots)))
db.close()
