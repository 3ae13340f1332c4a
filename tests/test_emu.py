"""The device source of the solver (piplib_b200/csrc/pip_solver.h) executed on the CPU by the
fiber emulator in tests/emu (32 cooperative fibers = one warp), against the reference's golden
vectors.  This keeps the kernel logic under test in the no-GPU CI run; it is a debugging aid,
not a product path (the shared library has no CPU solver).  Lane scheduling order is permuted
to expose missing warp synchronisation."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pyoracle as po
import emu

CLI = [c for c in load_golden("cli_suite.json")
       if "pipFile" not in c["name"] and not c["name"].startswith(("boulet", "challenges"))]
RCLI = load_golden("random_cli.json")
HEAVY = {"randt64", "randt108", "randt334"}


@pytest.fixture(scope="module", autouse=True)
def _build():
    emu.build()


@pytest.mark.parametrize("order_mode", [0, 1, 2])
def test_cli_fixtures(order_mode):
    out = emu.solve_tableau_cases(CLI, slack_level=3, work_words=1 << 18, order_mode=order_mode)
    bad = []
    for c, (st, cells, _) in zip(CLI, out):
        if st != c["ref_status"] or cells != c["ref_cells"]:
            bad.append(c["name"])
        elif c["golden_ll"] is not None:
            text = po.cli_output_text(c["comment"], st, [tuple(x) for x in cells])
            if po.strip_ws_lines(text) != po.strip_ws_lines(c["golden_ll"]):
                bad.append(c["name"] + ":text")
    assert not bad, bad


def test_random_tableaus():
    cases = [c for c in RCLI if c["name"] not in HEAVY]
    out = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2)
    bad = [c["name"] for c, (st, cells, _) in zip(cases, out)
           if st != c["ref_status"] or cells != c["ref_cells"]]
    assert not bad, bad


def test_global_memory_code_path():
    """the instantiation used by the global-memory classes (blocked row walks, options compiled in)"""
    cases = CLI + [c for c in RCLI if c["name"] not in HEAVY]
    out = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=2)
    bad = [c["name"] for c, (st, cells, _) in zip(cases, out)
           if st != c["ref_status"] or cells != c["ref_cells"]]
    assert not bad, bad


def test_capacity_escalation_is_reported():
    """a too-small arena must yield the internal CAPACITY status, never a wrong answer."""
    c = [x for x in load_golden("cli_suite.json") if x["name"] == "bouleti"][0]
    (st, cells, _), = emu.solve_tableau_cases([c], slack_level=2, work_words=2048)
    assert st == 4001 and cells == []


@pytest.mark.parametrize("staged", [0, 1])
def test_large_tableau_path_on_nonparametric_fixtures(staged):
    """the grid-per-problem code path (pip_large.h, here one emulated CTA) on every non-parametric
    fixture incl. vivien32 (260 cuts) and the random tableaus: same cells as the reference"""
    cases = [c for c in load_golden("cli_suite.json")
             if c["nparm"] == 0 and c["nc"] == 0 and "sysmo" not in c["name"]]
    cases += [c for c in RCLI if c["nparm"] == 0 and c["nc"] == 0]
    bad = []
    for c in cases:
        st, cells, info = emu.solve_large(c, cut_rows=400, order_mode=2, staged=staged)
        if st != c["ref_status"] or cells != c["ref_cells"]:
            bad.append(c["name"])
    assert len(cases) > 100 and not bad, bad


@pytest.mark.parametrize("staged", [0, 1])
def test_large_tableau_consecutive_ones_vs_oracle(port, staged):
    from workloads import synth
    nvar = 96
    tab = synth.consecutive_ones(nvar, nvar, seed=3)
    case = dict(nvar=nvar, nparm=0, ni=nvar, nc=0, bigparm=-1, nq=1, tab=tab.tolist(), ctx=[])
    st_o, cells_o = port.traiter(nvar, 0, nvar, 0, -1, 1, tab, [])
    st, cells, info = emu.solve_large(case, cut_rows=64, staged=staged)
    assert st == st_o and [tuple(x) for x in cells] == cells_o and info[0] > 0


@pytest.mark.parametrize("order_mode", [0, 2])
def test_int32_instantiation_is_exact_or_widens(order_mode):
    """PipSolver<int> (int32 storage, exact 64-bit intermediates): every fixture either gives the
    reference's cells or reports WIDEN (4003), never a different answer"""
    cases = CLI + [c for c in RCLI if c["name"] not in HEAVY]
    out = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=order_mode, narrow=1)
    bad, widened = [], 0
    for c, (st, cells, _) in zip(cases, out):
        if st == 4003:
            widened += 1
        elif st != c["ref_status"] or cells != c["ref_cells"]:
            bad.append(c["name"])
    assert not bad, bad
    assert 0 < widened < len(cases) // 3


def test_deepest_cut_and_dual_in_emulation(port):
    """PipOptions the fixtures never exercise, device source vs the oracle (itself pinned against the
    live reference by tests/golden/options_suite.json): Deepest_cut on every integer tableau,
    Compute_dual on the rational ones without parameters (nq bit 2 / bit 1, tests only)"""
    base = CLI + [c for c in RCLI if c["name"] not in HEAVY]
    deep = [dict(c, nq=5) for c in base if c["nq"] == 1]
    dual = [dict(c, nq=2) for c in base if c["nparm"] == 0 and c["nc"] == 0]
    assert len(deep) > 100 and len(dual) > 60
    for cases in (deep, dual):
        out = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=2)
        bad = []
        for c, (st, cells, _) in zip(cases, out):
            st_o, cells_o = port.traiter(c["nvar"], c["nparm"], c["ni"], c["nc"], c["bigparm"], c["nq"],
                                         c["tab"], c["ctx"])
            if st_o >= 3000 or st == 4001:        # oracle time-out / CAPACITY: re-run one class up by the host
                continue
            if st != st_o or [tuple(x) for x in cells] != cells_o:
                bad.append((c["name"], st, st_o))
        assert not bad, bad


def test_warp_decoder_agrees_with_thread_decoder():
    """pip_decode_warp.h (one warp per problem, lane-parallel vectors, tiled output) against
    pip_decode.h on every golden cell stream, under every combination of the decode parameters
    (big-parameter column, Urs_parms, SHIFT / NEGATE / REMOVE) and both output widths."""
    import random
    rng = random.Random(7)
    streams = [c["ref_cells"] for c in CLI + RCLI if c["ref_status"] == 0 and c["ref_cells"]]
    streams += [[[3, 0, 0]], [[1, 0, 0]], []]
    # a vector longer than the tile (scalar path) and one that exactly fills it
    for m in (200, 127):
        streams.append([[3, 1, 0], [4, m, 0]] + [[7, 6 * j - 5, 1 + (j % 4)] for j in range(m)])
    checked = 0
    for cells in streams:
        # parameter combinations the library can produce: REMOVE / SHIFT need a big-parameter column
        # inside every vector, the Urs_parms copies sit right after it (source/piplib.c:850-867)
        shortest = min([c[1] for c in cells if c[0] == 4] + [64])
        for trial in range(6):
            urs = rng.choice([0, 0, 1, 2])
            bg = rng.choice([-1, 0, 1, shortest - 2 - 2 * urs])
            flags = rng.choice([0, 1, 2, 3, 5, 7, 4])
            if bg < 0 or bg + 1 + 2 * urs > shortest - 1:
                bg, urs, flags = -1, 0, flags & 2
            if trial == 0:
                bg, urs, flags = -1, 0, 0
            narrow = trial & 1
            a = emu.decode(0, cells, bg, urs, flags, narrow=narrow)
            b = emu.decode(1, cells, bg, urs, flags, narrow=narrow, order_mode=trial % 3)
            assert a[:4] == b[:4], (cells[:12], bg, urs, flags)
            assert (a[4] == b[4]).all(), (cells[:12], bg, urs, flags)
            checked += 1
    assert checked > 300


def _dense_to_cases(dom, ctx, nq=1):
    """PolyLib rows [flag | unknowns | parameters | constant] -> the .dat view [unknowns | constant |
    parameters], an equality as a pair of rows (source/tab.c:292-393 without options)"""
    import numpy as np
    n, dr, dc = dom.shape
    np_ = ctx.shape[2] - 2
    nv = dc - 2 - np_
    cases = []
    for i in range(n):
        rows = []
        for r in dom[i]:
            t = list(r[1:1 + nv]) + [r[-1]] + list(r[1 + nv:1 + nv + np_])
            rows.append(t)
            if r[0] == 0:
                rows.append([-x for x in t])
        crow = []
        for r in ctx[i]:
            t = list(r[1:1 + np_]) + [r[-1]]
            crow.append(t)
            if r[0] == 0:
                crow.append([-x for x in t])
        cases.append(dict(nvar=nv, nparm=np_, ni=len(rows), nc=len(crow), bigparm=-1, nq=nq,
                          tab=[[int(x) for x in t] for t in rows], ctx=[[int(x) for x in t] for t in crow]))
    return cases


@pytest.mark.parametrize("narrow", [0, 1])
def test_register_resident_feasibility_solves_equal_the_general_path(port, narrow):
    """pip_subsolve_regs (the compa_test / context feasibility solves with the sub-tableau in registers)
    (built with -DPIP_USE_SUBREG; an experiment that lost on the GPU to instruction supply, profiles/README.md,
    kept compiled out) against the default device source (every sub-solve through the arena): same
    status, same cells, and the same counters problem by problem -- pivots, cuts, sub-solves, element
    updates, largest tableau -- on the parametric fixtures, the random tableaus and the bench workloads;
    pivot totals against the oracle port"""
    import os
    from workloads import synth
    so2 = emu.SO                      # the default build: every sub-solve through the arena
    so1 = emu.build(defines=("PIP_USE_SUBREG",), so=os.path.join(os.path.dirname(emu.SO), "libpipemu_subreg.so"))
    cases = [c for c in CLI + RCLI if c["nparm"] > 0 and c["name"] not in HEAVY]
    for wl, n in (("loopnest16x24p3", 160), ("loopnest8x12p2", 300), ("sor1d", 200), ("cg1", 100), ("fimmel", 40)):
        dom, ctx = synth.generate(wl, n, seed=5)
        cases += _dense_to_cases(dom, ctx)
    assert len(cases) > 800
    a = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=narrow, so=so1)
    b = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=0, narrow=narrow, so=so2)
    bad, piv, subs = [], 0, 0
    keys = ("status", "ncells", "pivots", "cuts", "subsolves", "splits", "max_rows", "max_cols",
            "elem_updates_lo", "elem_updates_hi")
    for k, (c, (st, cells, r), (st2, cells2, r2)) in enumerate(zip(cases, a, b)):
        if st == 4003 or st2 == 4003:
            # WIDEN: the problem is re-run by the int64 class, its counters are dropped.  (The two builds need not
            # agree on it: the row update asks for |z| < 2^30, the register sub-solver for the int32 range.)
            continue
        if st != st2 or cells != cells2 or any(int(r[x]) != int(r2[x]) for x in keys):
            bad.append((k, st, st2, [(x, int(r[x]), int(r2[x])) for x in keys if int(r[x]) != int(r2[x])]))
        piv += int(r["pivots"])
        subs += int(r["subsolves"])
    assert not bad, bad[:5]
    assert subs > 8000 and piv > 20000, (subs, piv)
    if narrow == 0:
        total = 0
        for c in cases:
            stt = po.PortStats()
            port.traiter(c["nvar"], c["nparm"], c["ni"], c["nc"], c["bigparm"], c["nq"], c["tab"], c["ctx"], stats=stt)
            total += int(stt.pivots)
        assert total == piv


@pytest.mark.parametrize("narrow", [0, 1, 2])
def test_word_mode_equals_cells_then_decode(narrow):
    """word mode (the solver writes the serialised quast itself) against the cell stream of the same
    problem run through the reference decoder pip_ser_cells: the same words (the hash is taken from the
    words by the copy kernel), the same counters -- fixtures, random tableaus and the bench workloads, int64 / int32 / global-memory builds"""
    from workloads import synth
    cases = [c for c in CLI + RCLI if c["name"] not in HEAVY]
    for wl, n in (("loopnest16x24p3", 120), ("loopnest8x12p2", 200), ("sor1d", 150), ("fimmel", 30)):
        dom, ctx = synth.generate(wl, n, seed=17)
        cases += _dense_to_cases(dom, ctx)
    a = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=narrow, emit_words=True)
    b = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=0, narrow=narrow)
    keys = ("status", "ncells", "pivots", "cuts", "subsolves", "splits", "ser_words")
    bad, checked = [], 0
    for k, ((st, words, r, h), (st2, cells, r2)) in enumerate(zip(a, b)):
        if st != st2 or (st != 4003 and any(int(r[x]) != int(r2[x]) for x in keys if x != "ser_words")):
            bad.append((k, st, st2))
            continue
        if st not in (0, 1):
            continue
        ok, ln, h2, wide, w2 = emu.decode(0, cells, -1, 0, 0) if st == 0 else (1, 1, None, 0, np.asarray([-1]))
        if st == 0 and (words != [int(x) for x in w2] or len(words) != ln):
            bad.append((k, "words", len(words), ln))
        if st == 1 and words != [-1]:
            bad.append((k, "void", words))
        checked += 1
    assert not bad, bad[:6]
    assert checked > 600


@pytest.mark.parametrize("narrow", [0, 1])
def test_subtree_donation_splices_to_the_same_stream(narrow):
    """subtree donation in test mode (PipSteal mode 2, pip_types.h): the ELSE branch of every outermost open
    split is published as an offer, the donor finishes its own part, the offered subtrees are then solved as
    separate segments (which donate in turn) and the segment walk of the copy kernel (pip_segments.h)
    splices them in pre-order.  The resulting stream, status and counters must be those of the undonated
    solve -- parametric fixtures, random tableaus, bench workloads, and a tight SOL_SIZE that makes the
    cell limit fall into donated segments."""
    from workloads import synth
    cases = [c for c in CLI + RCLI if c["nparm"] > 0 and c["name"] not in HEAVY]
    for wl, n in (("loopnest16x24p3", 120), ("loopnest8x12p2", 200), ("sor1d", 100), ("fimmel", 30)):
        dom, ctx = synth.generate(wl, n, seed=29)
        cases += _dense_to_cases(dom, ctx)
    for sol_size in (0, 160):
        a = emu.solve_tableau_cases_steal(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=narrow,
                                          sol_size=sol_size)
        b = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=0, narrow=narrow,
                                    emit_words=True, sol_size=sol_size)
        bad, multi, fatal26 = [], 0, 0
        for k, ((st, words, r, nseg), (st2, words2, r2, _)) in enumerate(zip(a, b)):
            if st == 4003 or st2 == 4003:
                if st != st2:
                    bad.append((k, st, st2))
                continue
            if st != st2 or (st in (0, 1) and words != words2):
                bad.append((k, st, st2, nseg, len(words), len(words2)))
            elif st == 0 and any(int(r[x]) != int(r2[x]) for x in ("pivots", "cuts", "subsolves", "splits", "ncells")):
                bad.append((k, "counters", nseg))
            multi += nseg > 1
            fatal26 += st == 1026
        assert not bad, bad[:6]
        assert multi > 150, multi
        if sol_size:
            assert fatal26 > 20, fatal26


@pytest.mark.parametrize("narrow", [0, 1])
def test_heavy_problem_handover_gives_the_same_stream(narrow):
    """the engine's heavy-problem hand-over (PipLaunch::budget, pip_types.h): in a first launch without donation
    a problem that reaches a split with more than `budget` pivots behind it stops and is listed; the donation
    launch then solves the list from scratch behind the first launch's windows.  Streams, statuses and counters
    must be those of the plain solve whatever the budget (nothing of the abandoned attempt may survive)."""
    from workloads import synth
    cases = [c for c in CLI + RCLI if c["nparm"] > 0 and c["name"] not in HEAVY]
    for wl, n in (("loopnest16x24p3", 150), ("loopnest8x12p2", 200), ("fimmel", 30)):
        dom, ctx = synth.generate(wl, n, seed=31)
        cases += _dense_to_cases(dom, ctx)
    b = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=0, narrow=narrow, emit_words=True)
    for budget, cap, lo, hi in ((1, 1 << 30, 200, 10 ** 9), (40, 1 << 30, 20, len(cases) - 20), (1, 25, 25, 25),
                                (100000, 1 << 30, 0, 0)):
        a = emu.solve_tableau_cases_steal(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=narrow,
                                          budget=budget, handed_max=cap)
        handed = emu.solve_tableau_cases_steal.handed
        assert lo <= handed <= hi, (budget, handed)
        bad = []
        for k, ((st, words, r, nseg), (st2, words2, r2, _)) in enumerate(zip(a, b)):
            if st == 4003 or st2 == 4003:
                if st != st2:
                    bad.append((k, st, st2))
                continue
            if st != st2 or (st in (0, 1) and words != words2):
                bad.append((k, st, st2, nseg, len(words), len(words2)))
            elif st == 0 and any(int(r[x]) != int(r2[x]) for x in ("pivots", "cuts", "subsolves", "splits", "ncells")):
                bad.append((k, "counters", nseg))
        assert not bad, (budget, bad[:6])


@pytest.mark.parametrize("narrow", [0, 1])
def test_uniform_batch_layout_and_arena_images(narrow):
    """a dense batch as the engine runs it -- arena layout carved once for the largest row counts, arena images
    built ahead of the solve by the solver's own loader (pip_load_problem), two block copies per problem in the
    solver -- against the general path (layout and load per problem): same words, statuses, counters.  The
    batches mix problems with and without equality rows, so images are smaller than the layout's row count."""
    from workloads import synth
    for wl, n in (("loopnest16x24p3", 150), ("loopnest8x12p2", 250), ("sor1d", 150), ("fimmel", 40)):
        dom, ctx = synth.generate(wl, n, seed=43)
        if wl == "loopnest8x12p2":
            dom = dom.copy()
            dom[::3, 1, 0] = 0                      # every third problem: one equality (an extra tableau row)
        cases = _dense_to_cases(dom, ctx)
        a = emu.solve_uniform_cases(cases, slack_level=3, work_words=1 << 18, order_mode=2, narrow=narrow)
        b = emu.solve_tableau_cases(cases, slack_level=3, work_words=1 << 18, order_mode=0, narrow=narrow, emit_words=True)
        bad = []
        for k, ((st, words, r), (st2, words2, r2, _)) in enumerate(zip(a, b)):
            if st != st2 or (st in (0, 1) and words != words2):
                bad.append((wl, k, st, st2))
            elif st == 0 and any(int(r[x]) != int(r2[x]) for x in ("pivots", "cuts", "subsolves", "splits", "ncells")):
                bad.append((wl, k, "counters"))
        assert not bad, bad[:6]
