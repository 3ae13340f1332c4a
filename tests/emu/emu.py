"""DEBUGGING AID (tests only): drive the host-emulated device source (tests/emu/libpipemu.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

from piplib_b200.ctypes_defs import CELL_DTYPE, RESULT_DTYPE, pack_tableau_problems

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpipemu.so")


def build():
    srcs = [os.path.join(HERE, f) for f in ("emu_runtime.cpp", "emu_driver.cpp")]
    subprocess.check_call(["g++", "-O2", "-g", "-fwrapv", "-DPIP_EMU", "-fPIC", "-shared", "-Wall",
                           "-Wno-unused-function"] + srcs + ["-o", SO])


def solve_tableau_cases(cases, work_words=1 << 16, stack_words=1 << 20, slack_level=2, order_mode=0,
                        sol_size=0, maxcol=0):
    lib = C.CDLL(SO)
    probs, pool = pack_tableau_problems(cases)
    n = len(probs)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    cap = 4096 * (n + 1)
    cells = np.zeros(cap, dtype=CELL_DTYPE)
    lib.pipemu_solve_batch(probs.ctypes.data_as(C.c_void_p), n, pool.ctypes.data_as(C.c_void_p),
                           res.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p),
                           C.c_longlong(cap), work_words, C.c_longlong(stack_words), slack_level,
                           order_mode, sol_size, maxcol)
    out = []
    for i in range(n):
        r = res[i]
        c = cells[r["cell_off"]:r["cell_off"] + r["ncells"]]
        out.append((int(r["status"]), [[int(x["kind"]), int(x["p1"]), int(x["p2"])] for x in c], r))
    return out
