#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, without a GPU):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
Prints, per captured kernel instance: duration, DRAM bytes read/written, DRAM and tensor-pipe utilisation, issue
activity, achieved occupancy, registers -- the quantities the roofline entries of bench.py / DESIGN.md are judged on."""
import csv
import io
import re
import subprocess
import sys

M = [('gpu__time_duration.sum', 'dur'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
     ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
     ('lts__t_bytes.sum', 'l2_bytes'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
     ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
     ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'),
     ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
     ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
     ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%')]


def short(name):
    m = re.search(r'igemm_\w+_kernel<srl::(\w+)', name)
    if m:
        return m.group(0).split('<srl::')[0].replace('void ', '') + '<' + m.group(1) + '>'
    return name.split('(')[0].replace('void ', '')[:48]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {}
    for key, _ in M:
        for i, h in enumerate(hdr):
            if h == key:
                col[key] = i
    ki = hdr.index('Kernel Name')
    lines = ['| kernel | ' + ' | '.join(lbl for _, lbl in M) + ' |', '|---|' + '---|' * len(M)]
    for r in rows[2:]:
        vals = []
        for key, _ in M:
            i = col.get(key)
            vals.append('n/a' if i is None else f'{r[i]} {units[i]}'.strip())
        lines.append('| ' + short(r[ki]) + ' | ' + ' | '.join(vals) + ' |')
    out = '\n'.join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(f'# ncu summary of {rep}\n\n(ncu --set full --clock-control none; per-launch values, cold-ish caches, ~40 replays)\n\n' + out + '\n')
    print(out)


if __name__ == '__main__':
    main()
