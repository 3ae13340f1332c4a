#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference (run in the build container only).

Reads the reference's own fixtures under /root/reference/{test,example} (inputs + golden .ll
text) and runs the UNMODIFIED reference (oracle/_ref/libpipref.so, built by oracle/Makefile from
/root/reference/source) for live answers where the reference ships none: stale goldens
(boulet, bouleti, dirk), test/challenges/*, option variants the fixtures never exercise, and
seeded random problems (including the fatal verdicts).  The GPU box has no /root/reference, so
the tests only ever read the JSON written here.

usage: python tools/make_golden.py
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
DISABLED = {"boulet", "bouleti", "dirk"}      # test/Makefile.am:16-20


def cli_suite(ref):
    cases = []
    files = sorted(glob.glob(REF + "/test/*.dat")) + sorted(glob.glob(REF + "/test/challenges/*"))
    for f in files:
        name = os.path.basename(f)
        chal = "/challenges/" in f
        if name.endswith(".dat"):
            name = name[:-4]
        p = po.parse_dat(open(f, encoding="latin-1").read())
        st, cells = ref.traiter(p["nvar"], p["nparm"], p["ni"], p["nc"], p["bigparm"], p["nq"],
                                p["tab"], p["ctx"])
        ll = None
        llf = f[:-4] + ".ll"
        if not chal and name not in DISABLED and os.path.exists(llf):
            ll = open(llf, encoding="latin-1").read()
        p.update(name=("challenges/" + name) if chal else name, golden_ll=ll, ref_status=st,
                 ref_cells=[list(c) for c in cells])
        cases.append(p)
        # the rational twin / integer twin of every problem, live from the reference
        q = dict(p)
        q["nq"] = 1 - p["nq"]
        st, cells = ref.traiter(q["nvar"], q["nparm"], q["ni"], q["nc"], q["bigparm"], q["nq"],
                                q["tab"], q["ctx"])
        q.update(name=p["name"] + "@nq%d" % q["nq"], golden_ll=None, ref_status=st,
                 ref_cells=[list(c) for c in cells])
        cases.append(q)
    return cases


VARIANTS = [
    {},
    {"Nq": 0},
    {"Maximize": 1},
    {"Urs_unknowns": 1},
    {"Urs_parms": 1},
    {"Urs_parms": 1, "Urs_unknowns": 1},
    {"Simplify": 1},
    {"Maximize": 1, "Nq": 0},
]


def lib_suite(ref):
    cases = []
    for f in sorted(glob.glob(REF + "/example/*.pip")):
        name = os.path.basename(f)[:-4]
        p = po.parse_pip(open(f).read())
        for v in VARIANTS:
            opts = dict(p["opts"])
            tag = name
            if v:
                if any(p["opts"].get(k) == val for k, val in v.items()) and len(v) == 1:
                    continue
                opts.update(v)
                tag = name + "@" + ",".join("%s=%d" % kv for kv in sorted(v.items()))
            st, ser = ref.solve(p["dom"], p["ctx"], p["bignum"], ctx_cols=p["ctx_shape"][1], **opts)
            c = dict(p)
            c.update(name=tag, opts=opts, ref_status=st, ref_ser=ser,
                     golden_ll=open(f[:-4] + ".ll").read() if not v else None)
            cases.append(c)
    return cases


def random_lib_cases(ref, n, seed):
    """small random PolyLib problems through pip_solve (all verdicts kept)."""
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        nn = int(rng.integers(1, 5))
        npar = int(rng.integers(0, 4))
        nl = int(rng.integers(1, 7))
        nm = int(rng.integers(0, 3)) if npar else 0
        lim = int(rng.choice([1, 2, 3, 7]))
        dom = rng.integers(-lim, lim + 1, size=(nl, nn + npar + 2))
        dom[:, 0] = (rng.random(nl) > 0.15).astype(np.int64)      # a few equalities
        dom[:, -1] = rng.integers(-4, 13, size=nl)
        ctx = rng.integers(-lim, lim + 1, size=(nm, npar + 2))
        if nm:
            ctx[:, 0] = 1
            ctx[:, -1] = rng.integers(-2, 9, size=nm)
        opts = {"Nq": int(rng.random() > 0.25)}
        r = rng.random()
        if r < 0.1:
            opts["Maximize"] = 1
        elif r < 0.2:
            opts["Urs_unknowns"] = 1
        if rng.random() < 0.1 and npar:
            opts["Urs_parms"] = 1
        have_ctx = bool(npar) or rng.random() < 0.5
        st, ser = ref.solve(dom, ctx if have_ctx else None, -1,
                            ctx_cols=npar + 2, **opts)
        if len(ser) > 4000 or st >= 3000:
            continue
        cases.append(dict(name="rand%d" % i, dom_shape=list(dom.shape), dom=dom.tolist(),
                          ctx_shape=[nm, npar + 2] if have_ctx else None,
                          ctx=ctx.tolist() if have_ctx else None, bignum=-1, opts=opts,
                          ref_status=st, ref_ser=ser))
    return cases


def random_cli_cases(ref, n, seed):
    """random tableau-level problems (the .dat path), incl. big parameters and large entries."""
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        nvar = int(rng.integers(1, 6))
        nparm = int(rng.integers(0, 4))
        ni = int(rng.integers(1, 8))
        nc = int(rng.integers(0, 3)) if nparm else 0
        big = bool(rng.random() < 0.15) and nparm > 0
        bigparm = nvar + 1 + int(rng.integers(0, nparm)) if big else -1
        lim = int(rng.choice([1, 2, 5, 40, 100000]))
        tab = rng.integers(-lim, lim + 1, size=(ni, nvar + nparm + 1))
        ctx = rng.integers(-3, 4, size=(nc, nparm + 1))
        nq = int(rng.random() > 0.3)
        st, cells = ref.traiter(nvar, nparm, ni, nc, bigparm, nq, tab, ctx)
        if len(cells) > 1500 or st >= 3000:
            continue
        cases.append(dict(name="randt%d" % i, comment="", nvar=nvar, nparm=nparm, ni=ni, nc=nc,
                          bigparm=bigparm, nq=nq, tab=tab.tolist(), ctx=ctx.tolist(),
                          golden_ll=None, ref_status=st, ref_cells=[list(c) for c in cells]))
    return cases


def options_suite(ref, n, seed):
    """the options the reference's own tests never exercise (SURVEY.md 8c "unpinned"): Deepest_cut on
    every integer example and on random integer problems; Compute_dual on random rational problems
    without parameters (with parameters the reference reads uninitialised memory, source/traiter.c:585
    vs 616-617), a share of them with equalities (pip_quast_equalities_dual_xx)."""
    cases = []
    for f in sorted(glob.glob(REF + "/example/*.pip")):
        name = os.path.basename(f)[:-4]
        p = po.parse_pip(open(f).read())
        if not p["opts"].get("Nq", 1):
            continue
        opts = dict(p["opts"])
        opts["Deepest_cut"] = 1
        st, ser = ref.solve(p["dom"], p["ctx"], p["bignum"], ctx_cols=p["ctx_shape"][1], **opts)
        c = dict(p)
        c.update(name=name + "@Deepest_cut=1", opts=opts, ref_status=st, ref_ser=ser, golden_ll=None)
        cases.append(c)
    rng = np.random.default_rng(seed)
    for i in range(n):
        deepest = i % 2 == 0
        nn = int(rng.integers(1, 6))
        npar = int(rng.integers(0, 3)) if deepest else 0
        nl = int(rng.integers(1, 8))
        nm = int(rng.integers(0, 3)) if npar else 0
        lim = int(rng.choice([1, 2, 3, 7, 30]))
        dom = rng.integers(-lim, lim + 1, size=(nl, nn + npar + 2))
        dom[:, 0] = (rng.random(nl) > 0.2).astype(np.int64)
        dom[:, -1] = rng.integers(-6, 20, size=nl)
        ctx = rng.integers(-lim, lim + 1, size=(nm, npar + 2))
        if nm:
            ctx[:, 0] = 1
            ctx[:, -1] = rng.integers(-2, 9, size=nm)
        if deepest:
            opts = {"Nq": 1, "Deepest_cut": 1}
            have_ctx = bool(npar) or rng.random() < 0.5
        else:
            opts = {"Nq": 0, "Compute_dual": 1}
            have_ctx = False
        st, ser = ref.solve(dom, ctx if have_ctx else None, -1, ctx_cols=npar + 2, **opts)
        if len(ser) > 4000 or st >= 3000:
            continue
        cases.append(dict(name="opt%d" % i, dom_shape=list(dom.shape), dom=dom.tolist(),
                          ctx_shape=[nm, npar + 2] if have_ctx else None,
                          ctx=ctx.tolist() if have_ctx else None, bignum=-1, opts=opts,
                          ref_status=st, ref_ser=ser))
    return cases


def main():
    po.build()
    ref = po.Ref()
    ref.lib.pipref_set_timeout_ms(3000)       # run-away random problems are dropped (status 3000)
    os.makedirs(OUT, exist_ok=True)
    if "--options-only" in sys.argv:
        suites = {"options_suite.json": options_suite(ref, 400, 777)}
    else:
      suites = {
        "options_suite.json": options_suite(ref, 400, 777),
        "cli_suite.json": cli_suite(ref),
        "lib_suite.json": lib_suite(ref),
        "random_lib.json": random_lib_cases(ref, 400, 20261018),
        "random_cli.json": random_cli_cases(ref, 400, 4242),
      }
    for fn, cases in suites.items():
        path = os.path.join(OUT, fn)
        with open(path, "w") as f:
            json.dump(cases, f, separators=(",", ":"))
        stat = {}
        for c in cases:
            stat[c["ref_status"]] = stat.get(c["ref_status"], 0) + 1
        print("%-16s %4d cases  %8d bytes  statuses %s" % (fn, len(cases), os.path.getsize(path), stat))


if __name__ == "__main__":
    main()
