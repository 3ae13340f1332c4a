"""GPU diagnostic: time the kernel-only and host-buffer paths on a synthetic batch and, with
PIPLIB_B200_LIB pointing at the -DPIP_PROFILE build, print the per-phase warp-cycle split."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "loopnest16x24p3"
sizes = [int(x) for x in sys.argv[2:]] or [20000]
for n in sizes:
    dom, ctx = synth.generate(name, n)
    t = time.time()
    db = api.DeviceBatch(dom, ctx, -1)
    print("create %.3f s" % (time.time() - t), flush=True)
    for it in range(3):
        t = time.time()
        ms = db.run(False)
        s = api.last_stats()
        print("%s n=%d run wall %.3f dev_ms %.1f rounds %d launches %d pivots %d cells %d -> %.0f problems/s %.2f Mpivots/s"
              % (name, n, time.time() - t, ms, s.rounds, s.launches, s.pivots, s.cells, n / ms * 1e3,
                 s.pivots / ms / 1e3), flush=True)
    tot = sum(s.phase_cycles[:len(api.PHASES)])
    if tot:
        print("phase split (warp cycles): " + ", ".join("%s %.1f%%" % (p, 100.0 * c / tot)
              for p, c in zip(api.PHASES, s.phase_cycles)))
        print("warp-cycles per pivot: %.0f" % (tot / max(1, s.pivots)))
    st, _ = db.results(False)
    u, c = np.unique(st, return_counts=True)
    print(dict(zip(u.tolist(), c.tolist())))
    db.close()
    t = time.time()
    r = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True)
    s = api.last_stats()
    print("dense wall %.3f: kernel %.3f d2h %.3f h2d %.3f host %.3f (h2d %.1f MB, d2h %.1f MB)"
          % (time.time() - t, s.seconds_kernel, s.seconds_d2h, s.seconds_h2d, s.seconds_host,
             s.h2d_bytes / 1e6, s.d2h_bytes / 1e6), flush=True)
