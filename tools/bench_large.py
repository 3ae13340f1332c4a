"""BASELINE config 4: one large tableau (default 4096 x 4097 int64, 134 MB) solved by the whole
grid.  Prints one JSON line: pivots/s, per-pivot time, achieved algorithmic HBM bandwidth
(16*R*C + 8*C + 8*R bytes per pivot) against the measured peak, next to the
oracle's single-core time on the same problem (the reference cannot split one tableau)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
check = "--no-check" not in sys.argv
dense = "--dense" in sys.argv          # long intervals: nearly every row is touched by every pivot
tab = synth.consecutive_ones(n, n, seed=2026, dense=dense)
p = api.LargeProblem(n, n, 1, tab, cut_rows=1024, sol_size=1 << 20, maxcol=1 << 16)
print("warm %.3f ms" % p.run(), flush=True)
ms = []
for _ in range(reps):
    ms.append(p.run())
    print("run %.3f ms" % ms[-1], flush=True)
st, cells, info = p.fetch()
p.close()
best = float(np.median(ms))
piv = info["pivots"]
R, C = n - 1, n + 1
alg = (16.0 * R * C + 8.0 * C + 8.0 * R) * piv
peak = 6533.2
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
src = "fallback"
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
    src = "measured"
line = {"workload": "consecutive-ones %s%d x %d int64, Nq=1" % ("(long intervals) " if dense else "", n, n + 1), "status": st, "pivots": piv,
        "cuts": info["cuts"], "skipped_identity_rows": info["skipped_rows"],
        "identity_rows_skipped_frac": info["skipped_rows"] / max(1.0, float(R) * max(piv, 1)), "kernel_ms": best,
        "us_per_pivot": 1e3 * best / max(piv, 1),
        "phase_share": {"choice": info["cycles_choice"] / max(1, info["cycles_choice"] + info["cycles_update"]),
                        "update": info["cycles_update"] / max(1, info["cycles_choice"] + info["cycles_update"]),
                        "choice_sub_us_per_pivot": {k: v / 1965.0 / max(1, piv) for k, v in info["sub"].items() if k != "spare"}}, "pivots_per_sec": piv / (best / 1e3),
        "roofline": {"bound": "hbm", "achieved": alg / (best / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (best / 1e3) / 1e9 / peak, "peak_source": src,
                     "note": "dense 16*R*C figure; rows whose update is the identity are skipped"}}
if check:
    from oracle import pyoracle as po
    po.build(ref=False, port=True)
    t = time.perf_counter()
    stats = po.PortStats()
    st_o, cells_o = po.Port().traiter(n, 0, n, 0, -1, 1, tab, [], sol_size=1 << 20, maxcol=1 << 16, stats=stats)
    dt = time.perf_counter() - t
    line["cpu_baseline"] = {"seconds": dt, "pivots_per_sec": stats.pivots / dt, "cores": 1, "kind": "port"}
    line["parity"] = bool(st_o == st and [tuple(x) for x in cells] == cells_o and stats.pivots == piv)
w = info["sub"].get("spare", 0)
if w:
    line["walk_stats"] = {"rows_walked": w & 0xfffff, "min_ratio_zero": (w >> 20) & 0xfffff, "no_strike": (w >> 40) & 0xfffff}
print(json.dumps(line))
