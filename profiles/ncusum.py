import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
import re
pat=sys.argv[2] if len(sys.argv)>2 else None
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_global_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__cycles_active.avg','launch__grid_size','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    for i,h in enumerate(hdr):
        if h in want or (pat and re.search(pat,h)):
            print(f"  {h}: {r[i]} {units[i]}")
