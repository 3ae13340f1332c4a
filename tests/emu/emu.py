"""DEBUGGING AID (tests only): drive the host-emulated device source (tests/emu/libpipemu.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

from piplib_b200.ctypes_defs import CELL_DTYPE, RESULT_DTYPE, pack_tableau_problems

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpipemu.so")


def build(defines=(), so=None):
    """defines + so: a variant of the emulated device source (e.g. PIP_NO_SUBREG: every feasibility
    solve through the general path) next to the default build"""
    srcs = [os.path.join(HERE, f) for f in ("emu_runtime.cpp", "emu_driver.cpp")]
    subprocess.check_call(["g++", "-O2", "-g", "-fwrapv", "-DPIP_EMU", "-fPIC", "-shared", "-Wall",
                           "-Wno-unused-function", "-Wno-unknown-pragmas"] + ["-D" + d for d in defines] + srcs +
                          ["-o", so or SO])
    return so or SO


def solve_tableau_cases(cases, work_words=1 << 16, stack_words=1 << 20, slack_level=2, order_mode=0,
                        sol_size=0, maxcol=0, narrow=0, so=None, emit_words=False):
    """emit_words: word mode -- the solver writes the serialised quast itself (problems flagged SIMPLE_SER);
    the second element of every result is then the word list and a fourth one, the stream's hash, is added"""
    lib = C.CDLL(so or SO)
    probs, pool = pack_tableau_problems(cases)
    if emit_words:
        probs["flags"] |= 8                      # PIP_F_SIMPLE_SER
    n = len(probs)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    cap = 4096 * (n + 1)
    cells = np.zeros(cap, dtype=CELL_DTYPE)
    hashes = np.zeros(max(n, 1), dtype=np.uint64)
    lib.pipemu_solve_batch(probs.ctypes.data_as(C.c_void_p), n, pool.ctypes.data_as(C.c_void_p),
                           res.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p),
                           C.c_longlong(cap), work_words, C.c_longlong(stack_words), slack_level,
                           order_mode, sol_size, maxcol, narrow, 1 if emit_words else 0,
                           hashes.ctypes.data_as(C.c_void_p))
    out = []
    raw = cells.view(np.uint8)
    for i in range(n):
        r = res[i]
        if emit_words:
            nw = int(r["ser_words"]) if int(r["status"]) in (0, 1) else 0
            at = int(r["cell_off"]) * CELL_DTYPE.itemsize
            if narrow == 1:
                w = raw[at:at + 4 * nw].view(np.int32).astype(np.int64)
            else:
                w = raw[at:at + 8 * nw].view(np.int64)
            out.append((int(r["status"]), [int(x) for x in w], r, int(hashes[i])))
            continue
        c = cells[r["cell_off"]:r["cell_off"] + r["ncells"]]
        out.append((int(r["status"]), [[int(x["kind"]), int(x["p1"]), int(x["p2"])] for x in c], r))
    return out


def solve_uniform_cases(cases, work_words=1 << 16, stack_words=1 << 20, slack_level=2, order_mode=0, narrow=0):
    """a batch of ONE shape as the engine runs dense batches: arena layout carved once (PipLaunch::layout),
    arena images built ahead of the solve by the solver's own loader, word mode.  [(status, words, record)]"""
    lib = C.CDLL(SO)
    probs, pool = pack_tableau_problems(cases)
    probs["flags"] |= 8
    n = len(probs)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    cap = 4096 * (n + 1)
    cells = np.zeros(cap, dtype=CELL_DTYPE)
    lib.pipemu_solve_uniform(probs.ctypes.data_as(C.c_void_p), n, pool.ctypes.data_as(C.c_void_p),
                             res.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p), C.c_longlong(cap),
                             work_words, C.c_longlong(stack_words), slack_level, order_mode, narrow,
                             int(max(c["ni"] for c in cases)), int(max(c["nc"] for c in cases)))
    out = []
    raw = cells.view(np.uint8)
    for i in range(n):
        r = res[i]
        nw = int(r["ser_words"]) if int(r["status"]) in (0, 1) else 0
        at = int(r["cell_off"]) * CELL_DTYPE.itemsize
        w = raw[at:at + 4 * nw].view(np.int32).astype(np.int64) if narrow == 1 else raw[at:at + 8 * nw].view(np.int64)
        out.append((int(r["status"]), [int(x) for x in w], r))
    return out


def solve_tableau_cases_steal(cases, work_words=1 << 16, stack_words=1 << 20, slack_level=2, order_mode=0, narrow=0,
                              sol_size=0, budget=0, handed_max=1 << 30):
    """word mode with subtree donation in test mode (PipSteal mode 2): every outermost ELSE branch becomes a
    separate segment solved after the donor finished; returns [(status, words, record, segments)].
    budget > 0: the engine's heavy-problem hand-over -- a plain launch with a pivot budget, then the donation
    launch over the problems that stopped; solve_tableau_cases_steal.handed = how many did"""
    lib = C.CDLL(SO)
    probs, pool = pack_tableau_problems(cases)
    probs["flags"] |= 8
    n = len(probs)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    cap = 4096 * 64
    cells = np.zeros(cap, dtype=CELL_DTYPE)
    wcap = 1 << 22
    words = np.zeros(wcap, dtype=np.int64)
    woff = np.zeros(n + 1, dtype=np.int64)
    nsegs = np.zeros(n, dtype=np.int32)
    handed = C.c_int(0)
    lib.pipemu_solve_batch_steal(probs.ctypes.data_as(C.c_void_p), n, pool.ctypes.data_as(C.c_void_p),
                                 res.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p), C.c_longlong(cap),
                                 work_words, C.c_longlong(stack_words), slack_level, order_mode, sol_size, narrow,
                                 words.ctypes.data_as(C.c_void_p), C.c_longlong(wcap), woff.ctypes.data_as(C.c_void_p),
                                 nsegs.ctypes.data_as(C.c_void_p), C.c_uint(budget), C.byref(handed), C.c_uint(handed_max))
    solve_tableau_cases_steal.handed = handed.value
    return [(int(res[i]["status"]), [int(x) for x in words[woff[i]:woff[i + 1]]], res[i], int(nsegs[i])) for i in range(n)]


def solve_large(case, cut_rows=64, sol_size=0, maxcol=0, order_mode=0, staged=0):
    """one non-parametric problem through the grid-per-problem code path (one emulated CTA)"""
    assert case["nparm"] == 0
    lib = C.CDLL(SO)
    tab = np.ascontiguousarray(np.asarray(case["tab"], dtype=np.int64).reshape(-1))
    cap = max(sol_size, 4096) + 8
    cells = np.zeros(cap, dtype=CELL_DTYPE)
    st, nc = C.c_int(0), C.c_int(0)
    info = (C.c_longlong * 4)()
    lib.pipemu_large_staged(staged)
    lib.pipemu_solve_large(case["nvar"], case["ni"], case["nq"], tab.ctypes.data_as(C.c_void_p), cut_rows,
                           sol_size, maxcol, C.byref(st), cells.ctypes.data_as(C.c_void_p), C.byref(nc),
                           info, order_mode)
    c = cells[:nc.value]
    return st.value, [[int(x["kind"]), int(x["p1"]), int(x["p2"])] for x in c], list(info)


def decode(which, cells, bg, urs, flags, cap=1 << 16, narrow=0, order_mode=0, pad=64):
    """cells -> serialised quast words by the thread decoder (which=0, pip_decode.h) or by the
    warp-per-problem decoder (which=1, pip_decode_warp.h, 32 emulated lanes)"""
    lib = C.CDLL(SO)
    n = len(cells)
    arr = np.zeros(n + pad, dtype=CELL_DTYPE)
    for i, (k, a, b) in enumerate(cells):
        arr[i] = (k, 0, a, b)
    out = np.zeros(cap, dtype=np.int64)
    ln, h, w = C.c_longlong(0), C.c_ulonglong(0), C.c_uint(0)
    ok = lib.pipemu_decode(which, arr.ctypes.data_as(C.c_void_p), n, bg, urs, flags,
                           out.ctypes.data_as(C.c_void_p), C.c_longlong(cap), narrow, C.byref(ln), C.byref(h),
                           C.byref(w), order_mode)
    m = min(ln.value, cap)
    words = out[:m].copy() if not narrow else out.view(np.int32)[:m].astype(np.int64)
    return ok, ln.value, h.value, w.value, words
