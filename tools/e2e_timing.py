import os, sys, time
sys.path.insert(0, ".")
os.environ["PIPLIB_B200_TIMING"] = "1"
from piplib_b200 import api  # noqa: E402
from workloads import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dom, ctx = synth.generate("loopnest16x24p3", n)
r = None
for it in range(4):
    t = time.perf_counter()
    r = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True, out=r)
    print("e2e %.3f s -> %.0f problems/s; ser words %d" % (time.perf_counter() - t, n / (time.perf_counter() - t), int(r["ser_off"][n])), flush=True)
