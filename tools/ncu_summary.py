"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ cite.
usage: python tools/ncu_summary.py <file.ncu-rep> [out.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        # local memory = register spills + per-thread arrays: instructions executed, L1 requests / sectors
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "launch__local_size", "launch__occupancy_limit_warps", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
stall = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled")]
out = []
for v in vals:
    out.append("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w); out.append(f"{w:75s} {v[i]:>22s} {units[i]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try: st.append((float(v[hdr.index(h)].replace(",", "")), h))
            except ValueError: pass
    st.sort(reverse=True)
    out.append("-- top stall reasons (warps stalled per issue-active cycle) --")
    for val, h in st[:8]:
        out.append(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {val:8.3f}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
