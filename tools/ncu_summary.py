"""Summarise an ncu capture of pomcp_search_kernel into profiles/ (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py REPORT.ncu-rep "<how it was captured>" TAG launch=label[:generate_steps] ...
e.g. python tools/ncu_summary.py gpurun_out/r2_prof.ncu-rep "episode_steplog.py 12, --set full" r02 0=fresh_tree 4=in_episode_chain
writes profiles/<TAG>_search_kernel_raw.csv (raw page, all launches) and profiles/search_kernel_summary.json (what
bench.py quotes under roofline.*, with `from` = the file and the git revision the capture was taken at)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, how, tag = sys.argv[1], sys.argv[2], sys.argv[3]
picks = [a.split("=") for a in sys.argv[4:]] or [["0", "launch0"]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
raw_name = "%s_search_kernel_raw.csv" % tag
with open(os.path.join(ROOT, "profiles", raw_name), "w") as f:
    f.write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()


def summary(vals, gen_steps):
    def val(name):
        if name not in col:
            return None
        x, u = float(vals[col[name]].replace(",", "")), units[col[name]]
        x *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1.0)
        return x
    warp_inst = val("smsp__inst_executed.sum")
    thr_inst = val("smsp__thread_inst_executed.sum") or warp_inst * val("smsp__thread_inst_executed_per_inst_executed.ratio")
    st = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
    out = {
        "gpu_time_ms": val("gpu__time_duration.sum"), "grid": int(val("launch__grid_size")), "block": int(val("launch__block_size")),
        "registers": int(val("launch__registers_per_thread")),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "warp_inst": warp_inst, "threads_per_inst": round(thr_inst / warp_inst, 2),
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
        "l2_atom_sectors": val("lts__t_sectors_op_atom.sum") or val("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum"),
        "l2_red_sectors": val("lts__t_sectors_op_red.sum") or val("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum"),
        "shared_atom_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
        "l1_global_load_sectors": val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
        "l1_local_load_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        "l1_local_store_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
        "shared_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "stalls_per_issue": {k: val(st % k) for k in ("barrier", "long_scoreboard", "short_scoreboard", "wait", "sleeping",
                                                       "no_instruction", "not_selected", "math_pipe_throttle", "branch_resolving")},
    }
    if gen_steps:
        out["thread_inst_per_generate_step"] = thr_inst / gen_steps
    return out


res = {"capture": how, "kernel": "pomcp_search_kernel", "from": "profiles/%s@%s" % (raw_name, sha), "launches": {}}
for p in picks:
    idx, label = int(p[0]), p[1]
    label, _, gs = label.partition(":")
    res["launches"][label] = summary(rows[2 + idx], float(gs) if gs else None)
first = res["launches"][picks[-1][1].partition(":")[0]]          # the last pick is what bench.py quotes
res.update({k: first[k] for k in ("issue_active_pct", "dram_bytes_per_launch", "l2_atom_sectors", "gpu_time_ms") if k in first})
if "thread_inst_per_generate_step" in first:
    res["thread_inst_per_generate_step"] = first["thread_inst_per_generate_step"]
res["quoted_launch"] = picks[-1][1].partition(":")[0]
with open(os.path.join(ROOT, "profiles", "search_kernel_summary.json"), "w") as f:
    json.dump(res, f, indent=1)
    f.write("\n")
print(json.dumps(res, indent=1)[:3000])
