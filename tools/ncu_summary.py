"""Prints the metrics we care about from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:] 
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '')[:90])
    for w in WANT + extra:
        if w in d:
            print(f'   {w:85s} {d[w]:>16s} {units[hdr.index(w)]}')
