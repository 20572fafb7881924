"""Prints the key raw metrics of an .ncu-rep (one block per captured launch)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio','smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed_pipe_xu.sum','smsp__inst_executed_pipe_fma.sum','smsp__inst_executed_pipe_alu.sum',
 'smsp__inst_executed_pipe_lsu.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
 'smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_imc_miss_per_warp_active.pct',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio','local_load','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:60])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w); print(f'  {w:75s} {r[i]:>16s} {units[i]}')
    for i, w in enumerate(hdr):  # warp-state breakdown (names differ between ncu versions)
        if 'issue_stalled' in w and w.endswith('per_issue_active.ratio'):
            print(f'  {w:75s} {r[i]:>16s} {units[i]}')
