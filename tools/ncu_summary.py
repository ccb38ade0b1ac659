#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` dump of one bench.py pass into the two files bench.py and the docs read:

    python tools/ncu_summary.py gpurun_out/ncu_r1e_raw.csv[,more.csv] r1e "<command the capture ran>"

  profiles/ncu_<tag>_kernels.csv : one row per launch (duration, DRAM bytes, DRAM %, L1 data-pipe wavefronts %, tensor %,
                                   issue %, registers, occupancy, L2 / L1 hit rates)
  profiles/traffic_<tag>.json    : dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per kernel family
                                   under the names bench.py's `rooflines` uses
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size']


def family(short: str) -> str:
    if short.startswith('sep_fused3_kernel'):
        m = re.search(r'<\s*\d+,\s*\d+,\s*(\d+)', short)
        return 'sep_fused3_kernel[layers 8-12, tensor]' if m and m.group(1) == '512' else 'sep_fused3_kernel[layers 3-7, hbm]'
    if short.startswith('logmel2_kernel'):
        return 'logmel_kernel'
    for nm in ('pw_gemm_kernel', 'depthwise_kernel', 'l12_fused2_kernel', 'logmel_kernel', 'conv1_dw2_kernel',
               'pool_head_kernel', 'resample_tc_kernel', 'resample_kernel'):
        if short.startswith(nm):
            return nm
    return short


UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def main():
    srcs, tag = sys.argv[1].split(','), sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ''
    fam = {}
    out = os.path.join(ROOT, 'profiles', f'ncu_{tag}_kernels.csv')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'kernel'] + COLS)
        w.writerow(['', ''] + ['us', 'Mbyte', 'Mbyte'] + ['%'] * 6 + ['register/thread', '%', '%', '%', '', ''])
        n = 0
        for src in srcs:                                  # every capture file carries its own units row
            rows = list(csv.reader(open(src)))
            hdr, units = rows[0], rows[1]
            ix = {h: i for i, h in enumerate(hdr)}
            for r in rows[2:]:
                short = re.sub(r'\(.*', '', r[ix['Kernel Name']])
                short = re.sub(r'^(void )?(bd::)?(<?unnamed>::)?', '', short)
                us = float(r[ix[COLS[0]]]) * UNIT[units[ix[COLS[0]]]]
                rd = float(r[ix[COLS[1]]]) * UNIT[units[ix[COLS[1]]]]
                wr = float(r[ix[COLS[2]]]) * UNIT[units[ix[COLS[2]]]]
                w.writerow([n, short, f'{us:.2f}', f'{rd / 1e6:.2f}', f'{wr / 1e6:.2f}'] + [r[ix[c]] for c in COLS[3:]])
                n += 1
                d = fam.setdefault(family(short), {'bytes': 0.0, 'us': 0.0, 'n': 0})
                d['bytes'] += rd + wr; d['us'] += us; d['n'] += 1
    traffic = {k: {'dram_bytes_per_launch': v['bytes'] / v['n'], 'launches_captured': v['n'],
                   'ncu_us_per_launch': v['us'] / v['n']} for k, v in fam.items()}
    traffic['_source'] = f'ncu --set full --clock-control none, {cmd}: dram__bytes_read.sum + dram__bytes_write.sum ' \
                         f'averaged over the launches of one pass'
    json.dump(traffic, open(os.path.join(ROOT, 'profiles', f'traffic_{tag}.json'), 'w'), indent=1)
    print(out, {k: (round(v['dram_bytes_per_launch'] / 1e6, 1), v['launches_captured'], round(v['ncu_us_per_launch'], 1))
                for k, v in traffic.items() if k != '_source'})


if __name__ == '__main__':
    main()
