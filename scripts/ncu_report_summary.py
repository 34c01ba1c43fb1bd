#!/usr/bin/env python
"""Key metrics of an .ncu-rep (ncu -i ... --page raw --csv), one block per profiled launch."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:110])
        for k in hdr:
            base = k.split('.TriageCompute.')[-1]
            if base in KEYS or any(s in base for s in ('warp_issue_stalled', 'pcsamp_warps_issue_stalled')) and 'not_issued' not in base and float((r[hdr.index(k)] or '0').replace(',', '') or 0) > 0:
                print(f'  {base:85s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    main(sys.argv[1])
