"""profiles/ncu_summary.json from an `ncu --set full` report:  python scripts/ncu_summary.py <report.ncu-rep> [out.json]
Per kernel (averaged over the captured launches): duration, instruction counts, SIMT efficiency, pipe and issue
utilisation, occupancy, cache hit rates, DRAM bytes, main stall reasons; plus DRAM bytes per pass
(paths = setup + beam + trace kernels), which bench.py reports as roofline.traffic."""
import collections, csv, json, subprocess, sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_summary.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = {
    "duration_us": "gpu__time_duration.sum", "warp_inst_executed": "smsp__inst_executed.sum",
    "threads_per_warp_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "dram_read_MB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "issue_active_per_cycle": "smsp__issue_active.avg.per_cycle_active",
    "local_load_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_store_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "local_load_hit_pct": "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
}
scale = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}
acc = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("hdt::", "").replace("HashDagResolvedDevT<(bool)1>", "HashDagPrefixDev").replace("HashDagResolvedDevT<(bool)0>", "HashDagResolvedDev")
    for k, m in M.items():
        if m not in ix or r[ix[m]] in ("", "n/a"):
            continue
        v = float(r[ix[m]].replace(",", ""))
        u = units[ix[m]]
        if k in ("dram_read_MB", "dram_write_MB", "duration_us"):
            v *= scale.get(u, 1.0)
        acc[short][k].append(v)
kernels = {n: {k: round(sum(v) / len(v), 3) for k, v in d.items()} | {"launches_captured": len(d["duration_us"])} for n, d in acc.items()}
per_pass = {}
for p in ("paths", "shadows", "colors"):
    tot = 0.0
    for n, d in kernels.items():
        if p in n and ("HashDagDev" in n or "HashDagResolvedDev" in n or "HashDagPrefixDev" in n or "recorded" in n) or (p in n and "setup_" in n):
            tot += (d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0)) * 1e6
    per_pass[p] = tot
json.dump({"source": rep + " (ncu --set full --clock-control none; per-launch averages; cold caches, kernels serialised)",
           "kernels": kernels, "dram_bytes_per_pass": per_pass}, open(out, "w"), indent=1)
print(json.dumps(per_pass))
