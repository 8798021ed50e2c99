"""Summarise an ncu report (``ncu -i X.ncu-rep --page raw --csv``) to the handful of metrics the
roofline discussion needs.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [pattern]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "dram__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
    print("-- warp stall reasons (per warp active, pct) --")
    st = [(float(d[h].replace(',', '')), h) for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") and d[h]]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"   {h.replace('smsp__average_warp_latency_issue_stalled_', '').replace('smsp__average_warps_issue_stalled_', ''):60s} {v:8.2f}")
    if extra:
        for h in hdr:
            if extra in h:
                print(f"{h:75s} {d[h]:>18s}")
