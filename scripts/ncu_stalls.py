"""Per-SASS-region stall-reason breakdown of an ncu report (first or chosen launch).
usage: python scripts/ncu_stalls.py report.ncu-rep [launch_index] [region_size]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; RS = int(sys.argv[3]) if len(sys.argv) > 3 else 240
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'] + [len(rows)]
hdr = rows[hi[which]]
data = [dict(zip(hdr, r)) for r in rows[hi[which] + 1:hi[which + 1]] if len(r) == len(hdr) and r[0] != 'Address']
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = {s: sum(I(d[s]) for d in data) for s in st}
T = sum(tot.values())
print('all: ' + ' '.join('%s %.1f%%' % (s[6:], 100 * v / T) for s, v in sorted(tot.items(), key=lambda x: -x[1]) if v > 0.01 * T))
for s0 in range(0, len(data), RS):
    ch = data[s0:s0 + RS]
    t = {s: sum(I(d[s]) for d in ch) for s in st}
    tt = sum(t.values()); ins = sum(I(d['Instructions Executed']) for d in ch)
    if tt < 0.005 * T: continue
    print('%5d: smp %5.1f%% inst %9d | ' % (s0, 100 * tt / T, ins) + ' '.join('%s %.0f%%' % (s[6:], 100 * v / tt) for s, v in sorted(t.items(), key=lambda x: -x[1])[:5] if v))
