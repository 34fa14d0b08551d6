#!/usr/bin/env python
"""Per-opcode instruction counts and stall shares of a kernel's hot loop from an ncu report.

    python scripts/ncu_hot.py gpurun_out/prof_c3.ncu-rep bwd_dense 8212500 [min_exec]

arg 3 = warp-steps of the launch (warps x time steps) used to normalise counts per warp-step."""
import collections
import csv
import subprocess
import sys

rep, pat, wsteps = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{pat}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# one block per captured launch (each with its own header row); take the one that executed
# the most instructions
blocks, names, hdrs = [], [], []
for r in rows:
    if r and r[0] == 'Kernel Name':
        blocks.append([])
        names.append(r[1])
        hdrs.append(None)
    elif blocks and 'Source' in r and hdrs[-1] is None:
        hdrs[-1] = r
    elif blocks and hdrs[-1] is not None and len(r) == len(hdrs[-1]):
        blocks[-1].append(r)


def col(i, name):
    return hdrs[i].index(name)


k = max(range(len(blocks)), key=lambda i: sum(int(r[col(i, 'Instructions Executed')]) for r in blocks[i]
                                              if r[col(i, 'Instructions Executed')].isdigit()))
hdr = hdrs[k]
iS, iE, iW = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
body = [r for r in blocks[k] if r[iE].isdigit()]
print(names[k])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.45 * wsteps
hot = [r for r in body if int(r[iE]) > thr]
print('total warp instr', sum(int(r[iE]) for r in body), '| hot-loop instr / warp-step',
      round(sum(int(r[iE]) for r in hot) / wsteps, 1), f'({len(hot)} SASS lines)')
ops, stall = collections.Counter(), collections.Counter()
for r in hot:
    f = r[iS].split()
    o = (f[0] if not f[0].startswith('@') else f[1]).split('.')[0].rstrip(';')
    ops[o] += int(r[iE]) / wsteps
    stall[o] += int(r[iW])
ts = sum(stall.values()) or 1
for o, c in ops.most_common(28):
    print(f'{o:10s} {c:7.1f} / warp-step   stall samples {100 * stall[o] / ts:5.1f}%')
print('\ntop stall sites:')
for r in sorted(hot, key=lambda r: -int(r[iW]))[:18]:
    print(f'{int(r[iW]):6d} {r[iS][:100]}')
