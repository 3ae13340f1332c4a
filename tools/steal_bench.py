"""subtree donation on / off on small batches of heavy parametric families (kernel time of one dense call)"""
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402
for wl, n in (("boulet", 512), ("boulet", 2368), ("boulet", 9472), ("fimmel", 2368), ("fimmel", 9472), ("expansion", 2368), ("expansion", 9472), ("loopnest16x24p3", 2368), ("loopnest16x24p3", 9472), ("loopnest16x24p3", 37888), ("sor1d", 9472)):
    dom, ctx = synth.generate(wl, n, seed=2026)
    bg, opts = synth.bignum(wl), synth.options(wl)
    out = {}
    for mode in ("0", "1"):
        os.environ["PIPLIB_B200_STEAL"] = mode
        best = 1e30
        for it in range(4):
            t = time.perf_counter()
            r = api.solve_dense(dom, ctx, bg, want_hashes=True, want_ser=True, **opts)
            dt = time.perf_counter() - t
            if it:
                best = min(best, dt)
        out[mode] = (best, r["hashes"].copy(), r["status"].copy(), float(api.last_stats().device_ms))
    same = bool(np.array_equal(out["0"][1], out["1"][1]) and np.array_equal(out["0"][2], out["1"][2]))
    print(json.dumps({"workload": wl, "problems": n, "seconds_without": out["0"][0], "seconds_with_donation": out["1"][0],
                      "device_ms_without": out["0"][3], "device_ms_with": out["1"][3], "same_answers": same}), flush=True)
