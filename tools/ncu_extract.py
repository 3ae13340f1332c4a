"""Read one `ncu --set full` capture of the solve kernel (tools/gpu_final.sh) and refresh the files bench.py and
profiles/README.md quote:  python tools/ncu_extract.py <file.ncu-rep> <problems> <label> "<what>"
 - profiles/r2_solve_kernel_details_<label>.csv   the details page
 - profiles/r2_solve_kernel_traffic.json          per-launch DRAM bytes / instructions (bench.py scales them)
 - profiles/r2_solve_kernel_icache.json           one more row of the instruction-cache table
"""
import csv
import io
import json
import subprocess
import sys

rep, problems, label, what = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(k, scale_unit=True):
    u, v = d[k]
    x = float(v.replace(",", ""))
    if scale_unit:
        x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return x


det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
open("profiles/r2_solve_kernel_details_%s.csv" % label, "w").write(det)
kernel = d["Kernel Name"][1]
traffic = {
    "source": "ncu --set full --clock-control none, profiles/r2_solve_kernel_details_%s.csv, %s, one launch "
              "(python tools/ncu_run.py loopnest16x24p3 %d 1)" % (label, kernel, problems),
    "workload": "loopnest16x24p3", "problems": problems,
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "warp_instructions": num("smsp__inst_executed.sum"), "duration_ms": num("gpu__time_duration.sum", False),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "icc_hit_rate_pct": num("sm__icc_request_hit_rate.pct"),
    "sm_active_frac": num("sm__cycles_active.avg") / num("sm__cycles_elapsed.avg"),
    "note": "bench.py scales these per-problem figures to its batch for roofline.traffic and issue_roofline (same "
            "workload only); sm_active_frac < 1 is the tail of the launch: a few very large trees, one warp each",
}
json.dump(traffic, open("profiles/r2_solve_kernel_traffic.json", "w"), indent=1)
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__icc_requests.sum", "sm__icc_request_hit_rate.pct",
        "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread"]
row = {k: num(k, False) for k in keys if k in d}
row["problems"] = problems
row["what"] = what
row["gcc_instruction_requests_pct_of_peak_while_active"] = (
    row["gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"] * row["sm__cycles_elapsed.avg"] / row["sm__cycles_active.avg"])
row["warp_instructions_per_problem"] = row["smsp__inst_executed.sum"] / problems
ic = json.load(open("profiles/r2_solve_kernel_icache.json"))
ic["captures"][label] = row
json.dump(ic, open("profiles/r2_solve_kernel_icache.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
print(label, "instr/problem %.0f" % row["warp_instructions_per_problem"], "gcc while active %.1f %%" % row["gcc_instruction_requests_pct_of_peak_while_active"])
