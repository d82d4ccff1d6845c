import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__grid_size','launch__block_size','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum']
idx=[hdr.index(w) if w in hdr else -1 for w in want]
for r in rows[2:]:
    print('----')
    for w,i in zip(want,idx):
        if i>=0: print(f"  {w:75s} {r[i][:80]} {units[i]}")
