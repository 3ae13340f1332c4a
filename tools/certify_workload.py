#!/usr/bin/env python
"""Certify a synthetic workload range against run-away problems, using the reference itself
(oracle/_ref, this container only): every problem of [first, first+n) must finish within
--limit-ms.  Writes tests/golden/workload_<name>.json (status histogram, worst time).

usage: python tools/certify_workload.py loopnest16x24p3 1000000 [first] [--procs 8]
"""
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from workloads import synth  # noqa: E402


def work(args):
    name, first, n, limit = args
    ref = po.Ref()
    L = ref.lib
    L.pipref_set_timeout_ms(limit)
    dom, ctx = synth.generate(name, n, first=first)
    ser = np.zeros(1 << 20, dtype=np.int64)
    nn = C.c_long(0)
    opts = po.opts_array(**synth.options(name))
    hist, worst, slow = {}, 0.0, []
    dr, dc = dom.shape[1], dom.shape[2]
    cr, cc = ctx.shape[1], ctx.shape[2]
    for i in range(n):
        t = time.perf_counter()
        st = L.pipref_solve_ser(dr, dc, dom[i].ctypes.data_as(C.c_void_p), 1, cr, cc,
                                ctx[i].ctypes.data_as(C.c_void_p), synth.bignum(name), opts,
                                ser.ctypes.data_as(C.c_void_p), C.c_long(1 << 20), C.byref(nn), None, C.c_long(0))
        dt = time.perf_counter() - t
        hist[st] = hist.get(st, 0) + 1
        worst = max(worst, dt)
        if st >= 2000:
            slow.append((first + i, st))
    return hist, worst, slow


def main():
    name = sys.argv[1]
    n = int(sys.argv[2])
    first = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 0
    procs = 8
    if "--procs" in sys.argv:
        procs = int(sys.argv[sys.argv.index("--procs") + 1])
    limit = 2000
    per = synth.CHUNK * 4
    jobs = [(name, a, min(per, first + n - a), limit) for a in range(first, first + n, per)]
    t0 = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        outs = pool.map(work, jobs)
    hist, worst, slow = {}, 0.0, []
    for h, w, s in outs:
        for k, v in h.items():
            hist[k] = hist.get(k, 0) + v
        worst = max(worst, w)
        slow += s
    res = dict(workload=name, first=first, n=n, statuses={str(k): v for k, v in sorted(hist.items())},
               worst_seconds=worst, not_finished=slow, limit_ms=limit, wall_seconds=time.time() - t0)
    print(json.dumps(res))
    path = os.path.join(ROOT, "tests", "golden", "workload_%s_%d_%d.json" % (name, first, n))
    json.dump(res, open(path, "w"))


if __name__ == "__main__":
    main()
