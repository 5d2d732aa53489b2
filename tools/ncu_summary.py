"""Turns an .ncu-rep (ncu --set full) into the small JSON summaries kept under profiles/:
python tools/ncu_summary.py <report.ncu-rep> [kernel-substring] > profiles/<name>.json"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
out = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    name = d.get("Kernel Name", "")
    if want and want not in name:
        continue
    u = dict(zip(hdr, units))
    m = {k: {"value": d[k], "unit": u.get(k, "")} for k in KEYS if k in d}
    stalls = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(v.replace(",", "")) for k, v in d.items()
              if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
    out.append({"kernel": name[:160], "grid": d.get("Grid Size"), "block": d.get("Block Size"), "metrics": m,
                "warp_stall_cycles_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])})
print(json.dumps(out if len(out) != 1 else out[0], indent=1))
