"""Synthetic problem families for the BASELINE.json configurations.

Purely random coefficient matrices are useless as a benchmark (a quarter of them die with "solution
too complex"), so the families are *structured*: loop-nest shaped
polyhedra (bounds, parametric bounds, a few couplings) perturbed by a counter-based RNG.
Problem i of a family depends only on (seed, i): chunks of CHUNK problems are generated from
Philox streams keyed by (seed, chunk index), so any rank can generate any index range.

All generators return PolyLib-format dense arrays:
  dom [n, rows, 1 + nvar + nparm + 1]   row = [eq(0)/ineq(1) | unknowns | parameters | constant]
  ctx [n, crows, 1 + nparm + 1]
"""
import numpy as np

CHUNK = 1 << 14


def _rng(seed, chunk):
    return np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(chunk)]))


def _loopnest_chunk(rng, m, nvar, nrows, nparm, p2=0.05, p3=0.05):
    ncol = 1 + nvar + nparm + 1
    dom = np.zeros((m, nrows, ncol), dtype=np.int64)
    dom[:, :, 0] = 1
    ar = np.arange(m)
    X0, P0, K = 1, 1 + nvar, ncol - 1
    # rows 0..nvar-1: one bound per unknown, lower (x_i >= c) or parametric upper (x_i <= p_j + c)
    for i in range(min(nvar, nrows)):
        upper = rng.random(m) < 0.5
        pj = rng.integers(0, nparm, size=m)
        c_lo = rng.integers(0, 5, size=m)
        c_up = rng.integers(-4, 13, size=m)
        dom[ar, i, X0 + i] = np.where(upper, -1, 1)
        dom[ar[upper], i, P0 + pj[upper]] = 1
        dom[:, i, K] = np.where(upper, c_up, -c_lo)
    # remaining rows: couplings a*x_i - b*x_k (+/- p_j) + c >= 0 with small coefficients
    for r in range(nvar, nrows):
        i = rng.integers(0, nvar, size=m)
        k = (i + 1 + rng.integers(0, nvar - 1, size=m)) % nvar
        a = 1 + (rng.random(m) < p2)
        b = 1 + (rng.random(m) < p2)
        s = np.where(rng.random(m) < 0.5, 1, -1)
        dom[ar, r, X0 + i] = s * a
        dom[ar, r, X0 + k] = -s * b
        withp = rng.random(m) < 0.5
        pj = rng.integers(0, nparm, size=m)
        ps = np.where(rng.random(m) < 0.7, 1, -1)
        dom[ar[withp], r, P0 + pj[withp]] = ps[withp]
        third = rng.random(m) < p3
        t = (k + 1 + rng.integers(0, nvar - 2, size=m)) % nvar
        t = np.where(t == i, (t + 1) % nvar, t)
        t = np.where(t == k, (t + 1) % nvar, t)
        dom[ar[third], r, X0 + t[third]] += rng.integers(-2, 3, size=m)[third]
        dom[:, r, K] = rng.integers(-4, 13, size=m)
    # context: p_j >= c_j
    ctx = np.zeros((m, nparm, 1 + nparm + 1), dtype=np.int64)
    ctx[:, :, 0] = 1
    for j in range(nparm):
        ctx[:, j, 1 + j] = 1
        ctx[:, j, -1] = -rng.integers(0, 9, size=m)
    return dom, ctx


def loopnest(n, seed=2026, nvar=16, nrows=24, nparm=3, first=0, p2=0.05, p3=0.05):
    """BASELINE config 2: ~16 unknowns x 24 constraints, a few parameters.

    p2 = probability that a coupling coefficient is 2 instead of 1, p3 = probability of a third
    unknown in a coupling.  Both drive the Gomory-cut rate; at the defaults a problem needs ~60
    pivots, ~0.9 cuts and ~2.4 splits on average and one in ~6000 ends "solution too complex".
    Larger values grow a heavy tail of run-away cut chains on which the reference itself spends
    minutes per problem (measured: p2=0.5,p3=0.3 -> 1 in 20000), useless for a throughput figure."""
    doms, ctxs = [], []
    lo = first
    hi = first + n
    c = lo // CHUNK
    while c * CHUNK < hi:
        d, x = _loopnest_chunk(_rng(seed, c), CHUNK, nvar, nrows, nparm, p2, p3)
        a = max(lo, c * CHUNK) - c * CHUNK
        b = min(hi, (c + 1) * CHUNK) - c * CHUNK
        doms.append(d[a:b])
        ctxs.append(x[a:b])
        c += 1
    return np.ascontiguousarray(np.concatenate(doms)), np.ascontiguousarray(np.concatenate(ctxs))


def consecutive_ones(nvar, nrows, seed=2026, cmax=50, dense=False):
    """BASELINE config 4: one large tableau whose rows have the consecutive-ones property
    (sum_{j=a..b} x_j >= c): totally unimodular, so entries stay small through hundreds of pivots
    (dense random data overflows int64 within a few).  Returns the .dat-order tableau
    [nrows, nvar+1] = [coefficients | constant]."""
    rng = _rng(seed, 1 << 20)
    a = rng.integers(0, nvar, size=nrows)
    b = rng.integers(0, nvar, size=nrows)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    if dense:
        # long intervals: every row starts in the first eighth and ends in the last eighth of the columns, so
        # nearly every row holds the pivot column and a pivot's update touches nearly the whole tableau
        # (the default draw leaves 86 % of the row updates the identity, which the kernel skips)
        lo = rng.integers(0, max(1, nvar // 8), size=nrows)
        hi = nvar - 1 - rng.integers(0, max(1, nvar // 8), size=nrows)
    c = rng.integers(1, cmax + 1, size=nrows)
    j = np.arange(nvar)[None, :]
    tab = np.zeros((nrows, nvar + 1), dtype=np.int64)
    tab[:, :nvar] = ((j >= lo[:, None]) & (j <= hi[:, None])).astype(np.int64)
    tab[:, nvar] = -c
    return tab


def _chunked(gen, n, first):
    """problems [first, first+n) of a family whose chunk c is gen(c) -> (dom, ctx)"""
    doms, ctxs = [], []
    lo, hi = first, first + n
    c = lo // CHUNK
    while c * CHUNK < hi:
        d, x = gen(c)
        a = max(lo, c * CHUNK) - c * CHUNK
        b = min(hi, (c + 1) * CHUNK) - c * CHUNK
        doms.append(d[a:b])
        ctxs.append(x[a:b])
        c += 1
    return np.ascontiguousarray(np.concatenate(doms)), np.ascontiguousarray(np.concatenate(ctxs))


def perturbed(n, seed=2026, first=0, template="sor1d", amp=2, salt=0, unperturbed=()):
    """BASELINE config 5 (and 3, family B): a shipped problem shape with every constraint constant
    moved by a uniform draw from [-amp, amp] (no fatal verdicts in 20 000 draws per
    template, 4-100 pivots per problem).  `unperturbed` lists draws replaced by the template itself:
    the ones on which the reference does not terminate (a context sub-solve whose Gomory cuts never
    converge: it grows until the machine is out of memory), found by tools/certify_workload.py and
    recorded in tests/golden/workload_<name>_*.json."""
    from .templates import TEMPLATES
    t = TEMPLATES[template]
    dom0 = np.asarray(t["dom"], dtype=np.int64)
    ctx0 = np.asarray(t["ctx"], dtype=np.int64).reshape(len(t["ctx"]), t["ctx_cols"])
    key = sum(ord(ch) << (8 * i) for i, ch in enumerate(template[:6]))

    def gen(c):
        rng = _rng(seed ^ key ^ (salt << 48), c)
        dom = np.repeat(dom0[None], CHUNK, axis=0)
        dom[:, :, -1] += rng.integers(-amp, amp + 1, size=(CHUNK, dom0.shape[0]))
        for g in unperturbed:
            if c * CHUNK <= g < (c + 1) * CHUNK:
                dom[g - c * CHUNK] = dom0
        ctx = np.repeat(ctx0[None], CHUNK, axis=0)
        return dom, ctx
    return _chunked(gen, n, first)


def test_ni(n, seed=2026, first=0, N=10, spread=None):
    """BASELINE config 3, family A: the test<N>i.dat shape -- row i is sum_{j<=i} (j+1) x_j >= c_i
    with c_i around (i+1)! + 1, all-integer (deep Gomory cut chains, wide denominators)."""
    fact = np.cumprod(np.arange(1, N + 1, dtype=np.int64))

    def gen(c):
        rng = _rng(seed ^ (0x7e57 + N), c)
        dom = np.zeros((CHUNK, N, 1 + N + 1), dtype=np.int64)
        dom[:, :, 0] = 1
        for i in range(N):
            dom[:, i, 1:2 + i] = np.arange(1, i + 2)
            s = spread if spread is not None else max(1, int(fact[i]) // 8)
            dom[:, i, -1] = -(fact[i] + 1 + rng.integers(-s, s + 1, size=CHUNK))
        ctx = np.zeros((CHUNK, 0, 2), dtype=np.int64)
        return dom, ctx
    return _chunked(gen, n, first)


WORKLOADS = {
    "loopnest16x24p3": dict(fn=loopnest, kw=dict(nvar=16, nrows=24, nparm=3)),
    "loopnest8x12p2": dict(fn=loopnest, kw=dict(nvar=8, nrows=12, nparm=2)),
    # config 5: dependence-analysis shapes
    "sor1d": dict(fn=perturbed, kw=dict(template="sor1d")),
    "boulet": dict(fn=perturbed, kw=dict(template="boulet", unperturbed=(15234,))),
    "cg1": dict(fn=perturbed, kw=dict(template="cg1")),
    "fimmel": dict(fn=perturbed, kw=dict(template="fimmel")),
    "esced": dict(fn=perturbed, kw=dict(template="esced")),
    "expansion": dict(fn=perturbed, kw=dict(template="expansion")),
    # config 3: cut-heavy integer problems
    "test10i": dict(fn=test_ni, kw=dict(N=10)),
    "test12i": dict(fn=test_ni, kw=dict(N=12)),
    "vivien32": dict(fn=perturbed, kw=dict(template="vivien32", amp=1)),
}
BIGNUM = {"expansion": 21}                   # big-parameter column of the PolyLib matrix, else -1
OPTS = {"boulet": dict(Nq=0)}                # test/boulet.dat is the rational variant (bouleti the integer one)


def bignum(name):
    return BIGNUM.get(name, -1)


def options(name):
    return dict(OPTS.get(name, {}))


def generate(name, n, seed=2026, first=0):
    w = WORKLOADS[name]
    return w["fn"](n, seed=seed, first=first, **w["kw"])
