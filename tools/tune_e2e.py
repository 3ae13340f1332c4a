"""GPU tuning aid: e2e throughput of pip_solve_dense_dp for several chunk / lane settings."""
import os
import sys
import time

sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dom, ctx = synth.generate("loopnest16x24p3", n)
grid = [(131072, 4, 8, 3), (131072, 4, 16, 3), (131072, 6, 8, 3), (131072, 5, 8, 3), (196608, 4, 8, 4),
        (262144, 4, 8, 4), (262144, 3, 12, 4), (98304, 6, 8, 3), (65536, 8, 6, 2)]
for chunk, lanes, threads, ramp in grid:
    os.environ["PIPLIB_B200_CHUNK"] = str(chunk)
    os.environ["PIPLIB_B200_LANES"] = str(lanes)
    os.environ["PIPLIB_B200_THREADS"] = str(threads)
    os.environ["PIPLIB_B200_RAMP"] = str(ramp)
    best = 1e9
    res = None
    for it in range(4):
        t = time.perf_counter()
        res = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True, out=res)
        dt = time.perf_counter() - t
        if it:
            best = min(best, dt)
    s = api.last_stats()
    print("chunk %6d lanes %d threads %d ramp %d: %.3f s -> %.0f problems/s (kernel %.3f h2d %.3f d2h %.3f host %.3f; h2d %.0f MB d2h %.0f MB)"
          % (chunk, lanes, threads, ramp, best, n / best, s.seconds_kernel, s.seconds_h2d, s.seconds_d2h, s.seconds_host,
             s.h2d_bytes / 1e6, s.d2h_bytes / 1e6), flush=True)
