#!/usr/bin/env python
"""bench.py -- the hot path of piplib-b200 on BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload NAME]
  python bench.py --impl reference ...      (the reference's own CPU path, all host cores)

A "step" is one pass of the solver over one batch of B synthetic problems per GPU (weak
scaling: every rank solves its own B problems; no data-path collective exists).
  value   problems/s, whole job, inputs already resident in HBM (kernels only, CUDA events)
  e2e     problems/s through the C-ABI call pip_solve_dense_dp with HOST buffers: PolyLib
          matrices in, serialised quasts + hashes out, host<->device copies inside the region
The workload is configs[1] of BASELINE.json: ~16 unknowns x 24 constraints, 3 parameters
(piplib_b200/synth.py: loopnest16x24p3), data = synthetic.  Inputs (4 GB per 10^6 problems) are
far larger than L2, so no explicit L2 flush is needed between steps.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from piplib_b200 import dist as pdist  # noqa: E402
from piplib_b200 import synth  # noqa: E402

METRIC = "problems_per_sec"
UNIT = "problems/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self):
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for f in self.samples:
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arms (the checker libraries; never the thing shipped)
# ------------------------------------------------------------------------------------------
_CPU_DATA = None      # (dom, ctx) inherited by the forked workers (never pickled)


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
        piv = stats.pivots
    return sec, st, h, piv


def cpu_arm(dom, ctx, sample, cores, bg=-1, opts=None):
    """reference CPU path on `cores` processes (one per core: the library is not re-entrant,
    SURVEY.md 8b), static split of problems [0, sample)."""
    from oracle import pyoracle as po
    kind = "reference" if os.path.exists(po.REF_SO) else "port"
    if kind == "port":
        po.build(ref=False, port=True)
    per = (sample + cores - 1) // cores
    jobs = []
    for c in range(cores):
        a, b = c * per, min(sample, (c + 1) * per)
        if a < b:
            jobs.append((kind, a, b - a, c, bg, opts or {}))
    global _CPU_DATA
    _CPU_DATA = (dom[:sample], None if ctx is None else ctx[:sample])
    ctxm = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctxm.Pool(len(jobs)) as pool:
        outs = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    tmax = max(o[0] for o in outs)
    status = np.concatenate([o[1] for o in outs])
    hashes = np.concatenate([o[2] for o in outs])
    return dict(kind=kind, seconds=tmax, wall=wall, status=status, hashes=hashes, cores=len(jobs),
                n=sample)


def port_pivots(dom, ctx, sample):
    """pivot count of a sample (the reference has no counter; the oracle restatement, which gives
    bit-identical answers, counts them)."""
    from oracle import pyoracle as po
    po.build(ref=False, port=True)
    sec, st, h, stats = po.Port().bench_dense(0, sample, dom, ctx, -1)
    return stats.pivots, stats.elem_updates


def large_tableau_line(pk, pk_src, n=4096, reps=2):
    """BASELINE config 4 (the HBM-bound kernel of the path): one n x (n+1) int64 tableau solved by the
    whole grid, cooperative launch, CUDA events inside the library.  Algorithmic bytes per pivot =
    16*R*C + 8*C + 8*R (dense figure)."""
    from piplib_b200 import api
    tab = synth.consecutive_ones(n, n, seed=2026)
    p = api.LargeProblem(n, n, 1, tab, cut_rows=1024, sol_size=1 << 20, maxcol=1 << 16)
    p.run()
    ms = min(p.run() for _ in range(reps))
    st, cells, info = p.fetch()
    p.close()
    piv = max(1, info["pivots"])
    R, C = n - 1, n + 1
    alg = (16.0 * R * C + 8.0 * C + 8.0 * R) * piv
    ach = alg / (ms / 1e3) / 1e9
    traffic = None            # DRAM bytes per launch from the committed ncu --set full capture (same n, same seed)
    tj = os.path.join(ROOT, "profiles", "r1_large_kernel_traffic.json")
    if os.path.exists(tj):
        t = json.load(open(tj))
        if t.get("n") == n and t.get("pivots") == info["pivots"]:
            traffic = float(t["dram_bytes_read"] + t["dram_bytes_write"])
    return {"workload": "consecutive-ones %d x %d int64, Nq=1, one problem over the whole grid" % (n, n + 1),
            "status": st, "pivots": info["pivots"], "kernel_ms": ms, "us_per_pivot": 1e3 * ms / piv,
            "pivots_per_sec": piv / (ms / 1e3),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"], "peak_source": pk_src, "traffic": traffic,
                         "algorithmic_bytes": alg,
                         "note": "dense algorithmic figure 16*R*C+8*C+8*R per pivot; rows whose update is the identity "
                                 "are skipped, so DRAM traffic is far below it (profiles/r1_large_kernel_traffic.json) "
                                 "and the fraction can pass 1: it compares the pivot time with one streaming pass "
                                 "over the dense tableau"}}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PIP_BENCH_BATCH", 1000000)))
    ap.add_argument("--workload", default="loopnest16x24p3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the config-4 large-tableau measurement")
    ap.add_argument("--check", type=int, default=4096, help="problems cross-checked against the oracle")
    a = ap.parse_args()

    rank, world, local = pdist.env_rank()
    cores = os.cpu_count() or 1
    W = max(a.warmup, 0)
    K = max(a.steps, 1)
    B = a.batch
    # shape of the workload from one generated problem (PolyLib rows: flag, unknowns, parameters, constant)
    d1, c1 = synth.generate(a.workload, 1, seed=a.seed)
    nparm = max(c1.shape[2] - 2, 0)
    config = {"workload": "%s: %d problems/GPU/step, %d unknowns x %d constraints, %d parameters, "
                          "int64, pip_solve path (Nq=%d)" % (a.workload, B, d1.shape[2] - 2 - nparm, d1.shape[1], nparm,
                                                             synth.options(a.workload).get("Nq", 1)),
              "batch_per_gpu": B, "l2": "inputs (%.1f GB/step) exceed L2; no flush needed" %
              (B * (d1.shape[1] * d1.shape[2] + c1.shape[1] * c1.shape[2]) * 8 / 1e9), "seed": a.seed}

    # ---------------- reference arm: the reference's own CPU implementation ----------------
    if a.impl == "reference":
        if rank != 0:
            return
        sample = a.cpu_sample or min(B, cores * 4096)
        dom, ctx = synth.generate(a.workload, sample, seed=a.seed, first=0)
        for _ in range(W):
            cpu_arm(dom, ctx, min(sample, cores * 256), cores)
        tot_s, r = 0.0, None
        for _ in range(K):
            r = cpu_arm(dom, ctx, sample, cores)
            tot_s += r["seconds"]
        value = sample * K / tot_s
        pivots, _ = port_pivots(dom, ctx, min(sample, 4096))
        ppp = pivots / min(sample, 4096)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * tot_s / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": config, "pivots_per_sec": value * ppp,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": "first %d problems of the workload per step, one process "
                                           "per core" % sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- our arm ---------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from piplib_b200 import api, build
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- piplib-b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator comes up (NCCL_DEBUG=VERSION):
        # stdout carries exactly one JSON line, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    api.set_device(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first, _ = pdist.problem_range(rank, world, B)
    dom, ctx = synth.generate(a.workload, B, seed=a.seed, first=first)

    # parity gate before any timing: a slice of this rank's batch against the oracle
    ncheck = min(a.check, B)
    if ncheck:
        from oracle import pyoracle as po
        po.build(ref=False, port=True)
        _, st_o, h_o, stats_o = po.Port().bench_dense(0, ncheck, dom[:ncheck], ctx[:ncheck], -1)
        r = api.solve_dense(dom[:ncheck], ctx[:ncheck], -1)
        st_g = np.where(r["status"] == 1, 0, r["status"])
        if not (np.array_equal(st_g, st_o) and np.array_equal(r["hashes"][st_o == 0], h_o[st_o == 0])):
            raise SystemExit("bench.py: GPU results differ from the oracle -- refusing to time")

    # kernel-only: inputs resident in HBM
    db = api.DeviceBatch(dom, ctx, -1)
    for _ in range(max(W, 3)):
        db.run(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    stats = None
    for _ in range(K):
        dev_ms += db.run(False)
        stats = api.last_stats()
        launches += stats.launches
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    status, _ = db.results(False)
    pivots_step = int(stats.pivots)
    elem_step = int(stats.elem_updates)
    cells_step = int(stats.cells)
    rounds = int(stats.rounds)
    db.close()

    # end to end through the C-ABI with host buffers
    e2e_s, h2d_b, d2h_b = None, 0, 0
    if not a.no_e2e:
        res = None                     # caller-owned result buffers, reused across steps
        for _ in range(max(W, 3)):
            res = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True, out=res)
        barrier()
        t1 = time.perf_counter()
        for _ in range(K):
            res = api.solve_dense(dom, ctx, -1, want_hashes=True, want_ser=True, out=res)
            s2 = api.last_stats()
            h2d_b, d2h_b = int(s2.h2d_bytes), int(s2.d2h_bytes)
        barrier()
        e2e_s = time.perf_counter() - t1
    if rank == 0:
        sampler.stop()

    # max over ranks
    (dev_ms, wall_ms, e2e_max), (pivots_all, elem_all, cells_all) = pdist.reduce_stats(
        [dev_ms, wall_ms, e2e_s or 0.0], [float(pivots_step), float(elem_step), float(cells_step)],
        device="cuda")

    if rank == 0:
        pk, pk_src = peaks()
        total_problems = B * world * K
        value = total_problems / (dev_ms / 1e3)
        ms_per_step = dev_ms / K
        # algorithmic HBM traffic of the solve kernel: every problem's input words are read once
        # and its solution cells written once (the working set itself lives in shared memory)
        mx = max(int(np.abs(dom).max()), int(np.abs(ctx).max()))
        elem = 1 if mx < 127 else 4 if mx < 2 ** 31 - 1 else 8       # the host ships the narrowest width
        in_bytes = (dom.shape[1] * (dom.shape[2] - 1) + ctx.shape[1] * (ctx.shape[2] - 1)) * elem + 32
        alg_bytes = float(B) * (in_bytes + 56) + 8.0 * cells_step
        achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
        traffic = None
        tj = os.path.join(ROOT, "profiles", "r1_solve_kernel_traffic.json")
        if os.path.exists(tj):
            t = json.load(open(tj))
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["problems"] * B
        uniq, counts = np.unique(status, return_counts=True)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
            "pivots_per_sec": pivots_all * K / (dev_ms / 1e3),
            "elem_updates_per_sec": elem_all * K / (dev_ms / 1e3),
            "pivots_per_problem": pivots_all / (B * world),
            "wall_ms_per_step": wall_ms / K,
            "rounds_per_step": rounds,
            "status_counts": {str(int(u)): int(c) for u, c in zip(uniq, counts)},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_src,
                         "algorithmic_bytes": alg_bytes,
                         "note": "class-S kernel keeps the tableau in shared memory: HBM sees only "
                                 "inputs+cells; the binding resource is the INT pipe / issue slots "
                                 "(see elem_updates_per_sec and profiles/)"},
        }
        if traffic is not None and t.get("warp_instructions"):
            # the binding resource of the shared-memory-resident kernel: warp-instruction issue slots.
            # instructions per problem come from the committed ncu capture of this kernel and workload
            # (profiles/), the rate is this run's; peak = SMs x 4 schedulers x 1 instruction/cycle x clock
            ipp = t["warp_instructions"] / t["problems"]
            clk = (line["clocks"].get("sm_mhz") or pk.get("sm_max_mhz") or 1965.0) * 1e6
            peak_issue = 148 * 4 * clk
            line["issue_roofline"] = {"bound": "issue", "achieved": ipp * value / world, "peak": peak_issue,
                                      "unit": "warp-instructions/s per GPU", "frac": ipp * value / world / peak_issue,
                                      "warp_instructions_per_problem": ipp,
                                      "source": "profiles/r1_solve_kernel_traffic.json (ncu smsp__inst_executed.sum)"}
        if e2e_max:
            line["e2e"] = {"value": B * world * K / e2e_max, "unit": UNIT,
                           "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                           "api": "pip_solve_dense_dp (host PolyLib matrices in, serialised quasts out)"}
        if world == 1 and not a.no_large:
            line["config4_large_tableau"] = large_tableau_line(pk, pk_src)
        if world == 1:
            sample = a.cpu_sample or min(B, cores * 2048)
            r = cpu_arm(dom, ctx, sample, cores)
            line["cpu_baseline"] = {"value": sample / r["seconds"], "unit": UNIT, "cores": r["cores"],
                                    "kind": r["kind"],
                                    "sample": "first %d problems of the batch, one process per core, "
                                              "pip_solve loop" % sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
