#!/usr/bin/env python
"""BASELINE configs 3 and 5: the cut-heavy and the dependence-analysis families.

  python tools/bench_configs.py [--n N] [--cpu-sample M] [name ...]

For every family: generate n problems, check a sample against the oracle (status + quast hash),
time the kernels with inputs resident in HBM (CUDA events inside the library), time the host-buffer
call pip_solve_dense_dp, and time the reference on the host cores on a bounded sample.
Prints one JSON line per family."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from piplib_b200 import api, build  # noqa: E402
from workloads import synth  # noqa: E402

DEFAULT_N = {"sor1d": 1000000, "cg1": 1000000, "fimmel": 200000, "esced": 500000, "expansion": 100000,
             "boulet": 20000, "test10i": 500000, "test12i": 500000, "vivien32": 4000}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--check", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    build.build()
    from oracle import pyoracle as po
    from oracle.cpu_arm import cpu_arm
    po.build(ref=False, port=True)
    port = po.Port()
    cores = os.cpu_count() or 1
    for name in a.names or list(DEFAULT_N):
        n = a.n or DEFAULT_N[name]
        bg, opts = synth.bignum(name), synth.options(name)
        dom, ctx = synth.generate(name, n)
        m = min(a.check, n)
        t0 = time.perf_counter()
        _, st_o, h_o, stats = port.bench_dense(0, m, dom[:m], ctx[:m], bg, **opts)
        cpu1 = (time.perf_counter() - t0) / m
        r = api.solve_dense(dom[:m], ctx[:m], bg, **opts)
        st_g = np.where(r["status"] == 1, 0, r["status"])
        ok = np.array_equal(st_g, st_o) and np.array_equal(r["hashes"][st_o == 0], h_o[st_o == 0])
        parity_pivots = int(api.last_stats().pivots) == int(stats.pivots)
        db = api.DeviceBatch(dom, ctx, bg, **opts)
        db.run(False)
        ms = min(db.run(False) for _ in range(a.steps))
        s = api.last_stats()
        status, _ = db.results(False)
        db.close()
        res = None
        res = api.solve_dense(dom, ctx, bg, want_hashes=True, want_ser=True, out=res, **opts)
        t1 = time.perf_counter()
        res = api.solve_dense(dom, ctx, bg, want_hashes=True, want_ser=True, out=res, **opts)
        e2e = time.perf_counter() - t1
        sample = int(max(cores, min(n, a.cpu_seconds / max(cpu1, 1e-7) * cores)))
        c = cpu_arm(dom, ctx, sample, cores, bg=bg, opts=opts)
        u, cnt = np.unique(status, return_counts=True)
        print(json.dumps({
            "workload": name, "n": n, "shape": list(dom.shape[1:]), "parity_sample": m, "parity": bool(ok),
            "pivot_count_equal": bool(parity_pivots),
            "problems_per_sec": n / (ms / 1e3), "pivots_per_sec": int(s.pivots) / (ms / 1e3),
            "elem_updates_per_sec": int(s.elem_updates) / (ms / 1e3),
            "pivots_per_problem": int(s.pivots) / n, "cuts_per_problem": int(s.cuts) / n,
            "max_rows": int(s.max_rows), "max_cols": int(s.max_cols), "rounds": int(s.rounds),
            "ms_kernels": ms, "e2e_problems_per_sec": n / e2e,
            "status_counts": {str(int(x)): int(y) for x, y in zip(u, cnt)},
            "cpu_baseline": {"value": sample / c["seconds"], "unit": "problems/s", "cores": c["cores"],
                             "kind": c["kind"], "sample": sample}}), flush=True)


if __name__ == "__main__":
    main()
