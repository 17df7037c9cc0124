"""Executed-instruction breakdown by opcode from an ncu report's source page (SASS view).
usage: python scripts/sass_breakdown.py <report.ncu-rep> <kernel-name regex> [launch index]"""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat, '--launch-skip', idx, '--launch-count', '1'],
                     capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
print(out[start - 1][:160])
rows = list(csv.DictReader(out[start:]))
by_op = collections.Counter()
stall = collections.Counter()
total = 0
for r in rows:
    src = r['Source'].split()
    if not src:
        continue
    op = src[1] if src[0].startswith('@') and len(src) > 1 else src[0]
    op = op.split('.')[0]
    if r['Address'] == 'Address':            # a second view of the same kernel follows: stop after the first
        break
    n = int(r['Instructions Executed'] or 0)
    by_op[op] += n
    total += n
    stall[op] += int(r['# Samples'] or 0)
print('warp instructions executed: %d' % total)
st = sum(stall.values()) or 1
print('%-10s %12s %7s %9s' % ('opcode', 'executed', 'share', 'samples'))
for op, n in by_op.most_common(22):
    print('%-10s %12d %6.1f%% %8.1f%%' % (op, n, 100.0 * n / total, 100.0 * stall[op] / st))
