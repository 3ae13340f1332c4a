"""One process, one pip_solve_dense_dp call, every visible GPU (pip_set_devices_dp): the chunks of the batch go
to whichever device has a free lane.  Prints one JSON line per device count.
   python tools/multi_gpu_onecall.py [n] [workload]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
wl = sys.argv[2] if len(sys.argv) > 2 else "loopnest16x24p3"
dom, ctx = synth.generate(wl, n)
bg, opts = synth.bignum(wl), synth.options(wl)
ng = torch.cuda.device_count()
api.pin(dom), api.pin(ctx)
out = api.alloc_result(n, pinned=True)
ref = None
for k in [1, 2, 4, 8]:
    if k > ng:
        break
    api.set_devices(list(range(k)))
    best = 1e30
    for it in range(3):
        t = time.perf_counter()
        out = api.solve_dense(dom, ctx, bg, want_hashes=True, want_ser=True, out=out, **opts)
        dt = time.perf_counter() - t
        if it:
            best = min(best, dt)
    h = out["hashes"].copy()
    same = True if ref is None else bool(np.array_equal(h, ref))
    if ref is None:
        ref = h
    print(json.dumps({"devices": k, "problems": n, "workload": wl, "seconds": best, "problems_per_sec": n / best,
                      "same_answers_as_one_device": same, "api": "one pip_solve_dense_dp call, pip_set_devices_dp",
                      "buffers": "pinned"}), flush=True)
api.set_devices([])
