#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a compact per-launch table: python scripts/ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__t_bytes.sum', 'l2_bytes'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'hmma%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'ld_sectors'),
        ('smsp__inst_executed.sum', 'inst')]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    cols = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    lines = ['| kernel | ' + ' | '.join(f'{lab} ({units[i]})' for i, lab in cols) + ' |', '|' + '---|' * (len(cols) + 1)]
    for r in body:
        name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[-48:]
        lines.append('| ' + name + ' | ' + ' | '.join(r[i] for i, _ in cols) + ' |')
    text = '\n'.join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
