"""small workload for compute-sanitizer (racecheck / memcheck): a few dozen light fixtures through
the class-S int64 and int32 kernels, a split-heavy one, and a small large-tableau solve."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from piplib_b200 import api  # noqa: E402
from workloads import synth  # noqa: E402

cases = json.load(open("tests/golden/cli_suite.json"))
light = [c for c in cases if c["name"] in ("max", "rairoi", "test7i", "loz", "invert", "pairi", "linear", "lineri",
                                            "equus", "petit", "maxb", "discr", "crescat", "test12i", "dirk")]
out = api.traiter_batch(light)
bad = [c["name"] for c, (st, cells) in zip(light, out) if st != c["ref_status"] or cells != c["ref_cells"]]
dom, ctx = synth.generate("loopnest8x12p2", 256, seed=1)
r = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True)
c = [x for x in cases if x["name"] == "test7i"][0]
p = api.LargeProblem(c["nvar"], c["ni"], c["nq"], c["tab"], cut_rows=64)
p.run()
st, cells, info = p.fetch()
p.close()
print("sanitize_run: fixtures bad=%s dense ok=%d large ok=%s" % (bad, int((r["status"] == 0).sum()), cells == c["ref_cells"]))
