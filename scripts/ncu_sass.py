"""SASS-level view of an ncu report (first captured launch): regions of 120 instructions with their share of stall samples /
executed instructions and dominant opcodes, then the instructions with the most samples.
usage: python scripts/ncu_sass.py report.ncu-rep [launch_index]"""
import csv, re, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'] + [len(rows)]
hdr = rows[hi[which]]
data = [dict(zip(hdr, r)) for r in rows[hi[which] + 1:hi[which + 1]] if len(r) == len(hdr) and r[0] != 'Address']
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(I(d['# Samples']) for d in data); toti = sum(I(d['Instructions Executed']) for d in data)
print('samples', tot, 'warp-instructions', toti, 'sass', len(data))
for s in range(0, len(data), 120):
    ch = data[s:s + 120]
    smp = sum(I(d['# Samples']) for d in ch); ins = sum(I(d['Instructions Executed']) for d in ch)
    ops = {}
    for d in ch:
        m = re.search(r'(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', d['Source'].strip())
        if m: ops[m.group(1)] = ops.get(m.group(1), 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print('%5d smp %5.1f%% inst %5.1f%% exec/inst %8d  %s' % (s, 100 * smp / tot, 100 * ins / toti, ins / max(len(ch), 1), top))
for idx, d in enumerate(data): d['idx'] = idx
for d in sorted(data, key=lambda d: -I(d['# Samples']))[:25]:
    print('%5d %6.2f%% %9s  %s' % (d['idx'], 100 * I(d['# Samples']) / tot, d['Instructions Executed'], d['Source'][:100]))
