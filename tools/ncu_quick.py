"""Key counters + top stall reasons of every kernel in an `ncu --page raw --csv` export.  Usage: ncu_quick.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_issued.sum',
        'sm__icc_requests_lookup_miss.sum', 'sm__icc_requests.sum', 'smsp__warps_eligible.avg.per_cycle_active']
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d['Kernel Name'][:70])
    for k in KEYS:
        if k in d and d[k] != '':
            print('   ', k, d[k])
    st = []
    for k in h:
        if k.startswith('smsp__average_warps_issue_stalled') or k.startswith('smsp__average_warp_latency_issue_stalled'):
            try: st.append((float(d[k].replace(',', '')), k))
            except ValueError: pass
    for v, k in sorted(st, reverse=True)[:9]:
        print('    stall', k.split('issue_stalled_')[1][:45], round(v, 3))
