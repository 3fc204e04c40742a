import csv,sys,subprocess,io
WANT=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active",
"launch__registers_per_thread","launch__block_size","launch__grid_size","launch__shared_mem_per_block_dynamic","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem",
"smsp__inst_executed.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct",
"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
"smsp__inst_executed_op_global_ld.sum","smsp__inst_executed_op_global_st.sum","smsp__inst_executed_op_global_atom.sum","smsp__inst_executed_op_local_ld.sum","smsp__inst_executed_op_local_st.sum",
"l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum","l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
"lts__t_sectors_op_atom.sum","lts__t_sectors_op_red.sum","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_sectors_srcunit_tex_op_write.sum",
"smsp__thread_inst_executed_per_inst_executed.ratio","sm__cycles_active.avg","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
raw=subprocess.run(["ncu","-i",sys.argv[1],"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)));hdr,units=rows[0],rows[1]
sel=sys.argv[2] if len(sys.argv)>2 else ""
for r in rows[2:]:
    d=dict(zip(hdr,r))
    if sel and sel not in d["Kernel Name"]: continue
    print("==",d["ID"],d["Kernel Name"][:80])
    for h,u in zip(hdr,units):
        if h in WANT and d.get(h,"")!="": print("  %-86s %s %s"%(h,d[h],u))
