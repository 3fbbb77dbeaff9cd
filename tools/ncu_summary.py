"""Key counters of every kernel in an ncu report -> text: python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print('-' * 100)
    for k in keys:
        for h in hdr:
            if h == k or h.endswith('.' + k) or (k in h and k == 'Kernel Name'):
                print('%-90s %-10s %s' % (h, u.get(h, ''), d.get(h, '')))
                break
