"""Print the key metrics of every kernel in an ncu report: usage: python scripts/ncu_metrics.py X.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_op_atom.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__lsu_writeback_active_mem_lg.sum', 'smsp__inst_executed_pipe_lsu.sum']
ki = hdr.index('Kernel Name')
for d in data:
    print('----', d[ki][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {units[i]:12s} {d[i][:40]}")
