#!/usr/bin/env python
"""Summarise ncu output into profiles/ (run here, on the CPU box).

    python scripts/ncu_summary.py launches gpurun_out/launches_c2.csv profiles/r01_launches_c2.md
    python scripts/ncu_summary.py full gpurun_out/prof_shard.ncu-rep profiles/r01_ncu_shard.md [workload]

`launches`: the `--metrics gpu__time_duration.sum` pass -> per-kernel count / total / share.
`full`:     one `--set full` capture -> the metrics DESIGN.md and bench.py's roofline quote
            (duration, DRAM bytes, issue utilisation, stall reasons, registers, occupancy) and
            profiles/roofline_traffic.json[workload][kernel] = DRAM bytes per launch.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'occupancy limit (registers), CTAs/SM'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (shared mem), CTAs/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active %'),
    ('sm__inst_executed_pipe_xu.sum', 'XU (SFU) instructions'),
    ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'stall dispatch / issue'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'eligible warps per scheduler cycle'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle / issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no_instruction / issue'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio_throttle / issue'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle / issue'),
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list — `{os.path.basename(src)}`\n\n')
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: '
                'compare shares, not absolutes).\n\n')
        f.write('| launches | total µs | share | kernel |\n|---:|---:|---:|---|\n')
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'| {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | `{k[:110]}` |\n')
        f.write(f'\ntotal {tot / 1e3:.1f} µs over {sum(a[0] for a in agg.values())} launches\n')
    print(open(dst).read())


def full(src, dst, workload=None):
    if src.endswith('.csv'):       # `ncu -i x.ncu-rep --page raw --csv` already run on the GPU box
        raw = open(src).read()
    else:
        raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full — `{os.path.basename(src)}`\n\n')
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            f.write(f'## `{name}`\n\n| metric | value |\n|---|---:|\n')
            vals = {}
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    vals[key] = (r[i], units[i])
                    f.write(f'| {label} (`{key}`) | {r[i]} {units[i]} |\n')
            try:
                def gb(k):
                    v, u = vals[k]
                    return float(v) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]
                t = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
                base = name.split('<')[0].replace('void ', '').strip()
                targs = [a.strip() for a in name.split('<')[1].split('>')[0].split(',')] if '<' in name else []
                if base == 'hbv_fwd_kernel' and len(targs) >= 3 and targs[2] == '0':
                    base = 'hbv_fwd_warmup'
                if base == 'hbv_fwd_lean_kernel' and len(targs) >= 9 and targs[8] == '0':
                    base = 'hbv_fwd_warmup'
                if base == 'hbv_fwd_pipe_kernel' and len(targs) >= 7 and targs[6] == '0':
                    base = 'hbv_fwd_warmup'
                traffic.setdefault(base, t)
                f.write(f'| **DRAM traffic per launch** | {t / 1e9:.3f} GB |\n')
            except Exception:
                pass
            f.write('\n')
    if workload:
        tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        cur = json.load(open(tp)) if os.path.exists(tp) else {}
        short = {'hbv_fwd_kernel': 'hbv_fwd', 'hbv_bwd_kernel': 'hbv_bwd',
                 'hbv_fwd_lean_kernel': 'hbv_fwd', 'hbv_bwd_lean_kernel': 'hbv_bwd',
                 'hbv_fwd_dense_kernel': 'hbv_fwd', 'hbv_bwd_dense_kernel': 'hbv_bwd',
                 'hbv_fwd_pipe_kernel': 'hbv_fwd', 'hbv_bwd_pipe_kernel': 'hbv_bwd'}
        cur[workload] = {}        # one capture = one consistent set of kernels
        for k, v in traffic.items():
            cur[workload][short.get(k, k)] = v
        json.dump(cur, open(tp, 'w'), indent=1, sort_keys=True)
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:])
