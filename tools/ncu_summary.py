"""Scratch: key metrics of every kernel in an .ncu-rep (ncu --page raw --csv) as a small table.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
R = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = R[0], R[1], R[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fp64.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_atom.sum', 'dram__sectors_read.sum', 'dram__sectors_write.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]:", " | ".join(r[i][:48] for r in rows))
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h and 'smsp__average_warps_issue_stalled' in h:
        vals = [float(r[i].replace(',', '')) for r in rows]
        if max(vals) > 0.3:
            print("stall", h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vals)
