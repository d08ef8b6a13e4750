"""Summarise an `ncu --set full` report (one kernel launch) into profiles/<name>.json + .md.
Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME "free text note"
Runs here (no GPU needed): `ncu -i ... --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    m = {"kernel": name}
    for k in KEYS:
        if k in d:
            try:
                m[k] = float(d[k].replace(",", ""))
            except ValueError:
                m[k] = d[k]
            m[k + " [unit]"] = units[hdr.index(k)]
    launches.append(m)
summary = {"report": rep, "note": note, "launches": launches}
if launches:
    m = launches[0]
    rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    if rd is not None and wr is not None:
        summary["traffic_bytes_per_launch"] = rd * scale.get(m["dram__bytes_read.sum [unit]"], 1.0) + \
            wr * scale.get(m["dram__bytes_write.sum [unit]"], 1.0)
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write("# ncu summary: %s\n\n%s\n\nsource report: `%s` (not committed: binary, MBs)\n\n" % (out.split("/")[-1], note, rep))
    for m in launches:
        f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % m["kernel"])
        for k in KEYS:
            if k in m:
                f.write("| %s | %s | %s |\n" % (k, m[k], m.get(k + " [unit]", "")))
        f.write("\n")
print("wrote", out + ".json", out + ".md", "traffic", summary.get("traffic_bytes_per_launch"))
