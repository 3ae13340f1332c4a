"""TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by piplib_b200/).

The reference's CPU path over a dense batch on all host cores: one process per core (the reference is
not re-entrant), static split of the index range, each worker pinned to a core.  Used by
bench.py (cpu_baseline leg, reference arm, parity gate before timing) and by tests/ (full-batch parity).
`kind` "reference" = oracle/_ref/libpipref.so (the unmodified reference), "port" = oracle/pip_oracle.c.
"""
import multiprocessing as mp
import os
import time

import numpy as np

_CPU_DATA = None      # (dom, ctx) inherited by the forked workers (never pickled)


def usable_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


def _cpu_worker(args):
    kind, first, count, core, bg, opts = args
    dom, ctx = _CPU_DATA
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    from oracle import pyoracle as po
    if kind == "reference":
        sec, st, h = po.Ref().bench_dense(first, count, dom, ctx, bg, **opts)
        piv = 0
    else:
        sec, st, h, stats = po.Port().bench_dense(first, count, dom, ctx, bg, **opts)
        piv = int(stats.pivots)
    return sec, st, h, piv


def cpu_arm(dom, ctx, sample, cores=None, bg=-1, opts=None, kind=None):
    """problems [0, sample) of (dom, ctx) on `cores` processes (int = that many of the usable cores,
    list = those core ids).  Returns dict(kind, seconds = slowest worker's own loop time, wall, status,
    hashes, cores, n, pivots (port only))."""
    from oracle import pyoracle as po
    if kind is None:
        kind = "reference" if os.path.exists(po.REF_SO) else "port"
    if kind == "port":
        po.build(ref=False, port=True)
    ids = usable_cores()
    if isinstance(cores, int):
        ids = ids[:max(1, cores)]
    elif cores is not None:
        ids = list(cores)
    sample = int(min(sample, dom.shape[0]))
    per = (sample + len(ids) - 1) // len(ids)
    jobs = []
    for k, c in enumerate(ids):
        a, b = k * per, min(sample, (k + 1) * per)
        if a < b:
            jobs.append((kind, a, b - a, c, bg, opts or {}))
    global _CPU_DATA
    _CPU_DATA = (dom[:sample], None if ctx is None else ctx[:sample])
    ctxm = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctxm.Pool(len(jobs)) as pool:
        outs = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    _CPU_DATA = None
    return dict(kind=kind, seconds=max(o[0] for o in outs), wall=wall,
                status=np.concatenate([o[1] for o in outs]), hashes=np.concatenate([o[2] for o in outs]),
                cores=len(jobs), n=sample, pivots=sum(o[3] for o in outs))


def same_answers(status_gpu, hashes_gpu, ref):
    """GPU statuses / quast hashes against a cpu_arm() result over the same first ref['n'] problems.
    The library reports an empty context as status 1 (VOID); both CPU harnesses say 0 + the stream [-1]."""
    n = ref["n"]
    st_g = np.where(status_gpu[:n] == 1, 0, status_gpu[:n])
    ok = ref["status"] == 0
    return bool(np.array_equal(st_g, ref["status"]) and np.array_equal(hashes_gpu[:n][ok], ref["hashes"][ok]))
