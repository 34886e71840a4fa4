"""Print the key metrics of an `ncu --page raw --csv` dump (one kernel per row)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
        'gpu__dram_throughput.avg.pct', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64',
        'sm__pipe_fp64_cycles_active.avg.pct', 'smsp__inst_executed_pipe_fp64', 'pipe_fp64',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled', 'sm__cycles_elapsed.avg ', 'sm__cycles_elapsed.max',
        'smsp__inst_executed.sum ', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
        'smsp__issue_active.avg.pct', 'launch__shared_mem_per_block', 'launch__grid_size', 'launch__block_size',
        'lts__t_bytes.sum ', 'sm__inst_executed_pipe_lsu', 'smsp__warps_eligible.avg.per_cycle_active']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, r):
        if any(k.strip() in h for k in keys):
            print(f'  {h} [{u}] = {v}')
