"""Turn an `ncu --page raw --csv` export into the short text summary kept under profiles/."""
import csv, sys
want = ('Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_alu.avg.pct',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct', 'sm__pipe_alu_cycles_active.avg.pct',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'smsp__inst_executed.sum', 'issue_stalled',
        'sm__throughput.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum ',
        'sm__cycles_elapsed.avg ', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__throughput.avg.pct')
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if any(h.startswith(w.strip()) or w in h for w in want):
            if 'per_second' in h or 'peak_sustained' in h and 'pct' not in h:
                continue
            print(f'{h:95s} {r[i]} {units[i]}')
    print('-' * 40)
