"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py quote.
Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep [longest] > profiles/rNN_xxx.txt   (runs where ncu is installed, no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def main(path, longest_only=False):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, data = rows[0], rows[1], rows[2:]
    name_i = h.index('Kernel Name')
    if longest_only:                 # one entry per kernel name: its longest launch (+ how many launches the capture held)
        ti = h.index('gpu__time_duration.sum')
        best, count = {}, {}
        for r in data:
            k = r[name_i].split('(')[0]
            count[k] = count.get(k, 0) + 1
            if k not in best or float(r[ti].replace(',', '')) > float(best[k][ti].replace(',', '')):
                best[k] = r
        data = list(best.values())
        print('# longest launch of every kernel in the capture; launches captured:', ', '.join(f'{k.split("::")[-1]} x{n}' for k, n in count.items()))
    for li, r in enumerate(data):
        print(f'=== launch {li}: {r[name_i][:90]}')
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f'    {k:85s} {r[i]:>16s} {units[i]}')
    print('\n(source: ncu --set full --clock-control none --import-source on; file', path, ')')


if __name__ == '__main__':
    main(sys.argv[1], len(sys.argv) > 2 and sys.argv[2] == 'longest')
