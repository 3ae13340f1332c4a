"""end-to-end timing of pip_solve_dense_dp with the per-chunk breakdown (PIPLIB_B200_TIMING):
   python tools/e2e_timing.py [n] [pinned|pageable] [workload]"""
import os, sys, time
sys.path.insert(0, ".")
os.environ["PIPLIB_B200_TIMING"] = "1"
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pinned = len(sys.argv) > 2 and sys.argv[2] in ("pinned", "alloc")
alloc = len(sys.argv) > 2 and sys.argv[2] == "alloc"
wl = sys.argv[3] if len(sys.argv) > 3 else "loopnest16x24p3"
dom, ctx = synth.generate(wl, n)
bg, opts = synth.bignum(wl), synth.options(wl)
r = api.alloc_result(n, pinned="alloc" if alloc else pinned)
if alloc:
    dom, ctx = api.pinned_copy(dom), api.pinned_copy(ctx)
elif pinned:
    api.pin(dom), api.pin(ctx)
for it in range(4):
    t = time.perf_counter()
    r = api.solve_dense(dom, ctx, bg, want_hashes=True, want_ser=True, out=r, **opts)
    dt = time.perf_counter() - t
    print("e2e[%s] %.3f s -> %.0f problems/s; ser words %d" % ("pinned" if pinned else "pageable", dt, n / dt, int(r["ser_off"][n])), flush=True)
