"""Summarise an ncu capture of pomcp_search_kernel into profiles/ (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<how it was captured>" [generate_steps_per_launch]
writes profiles/r01_search_kernel_raw.csv (raw page) and profiles/search_kernel_summary.json (what bench.py quotes)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, how = sys.argv[1], sys.argv[2]
gen_steps = float(sys.argv[3]) if len(sys.argv) > 3 else 104.6e6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
with open(os.path.join(ROOT, "profiles", "r01_search_kernel_raw.csv"), "w") as f:
    f.write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def val(name, scale=1.0):
    v = vals[col[name]].replace(",", "")
    u = units[col[name]]
    x = float(v)
    if u in ("Mbyte",): x *= 1e6
    if u in ("Gbyte",): x *= 1e9
    if u in ("Kbyte",): x *= 1e3
    if u in ("us", "usecond"): x *= 1e-3          # -> ms
    if u in ("ns", "nsecond"): x *= 1e-6
    if u in ("s", "second"): x *= 1e3
    return x * scale


warp_inst = val("smsp__inst_executed.sum")
thr_inst = warp_inst * val("smsp__thread_inst_executed_per_inst_executed.ratio")
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
out = {
    "capture": how,
    "kernel": "pomcp_search_kernel",
    "gpu_time_ms": val("gpu__time_duration.sum"),
    "grid": int(val("launch__grid_size")),
    "block": int(val("launch__block_size")),
    "registers": int(val("launch__registers_per_thread")),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "warp_inst": warp_inst,
    "threads_per_inst": round(thr_inst / warp_inst, 2),
    "thread_inst_per_generate_step": thr_inst / gen_steps,
    "dram_bytes_per_launch": dram,
    "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
    "stall_barrier_per_issue": val("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    "stall_long_scoreboard_per_issue": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
}
with open(os.path.join(ROOT, "profiles", "search_kernel_summary.json"), "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
print(json.dumps(out, indent=1))
