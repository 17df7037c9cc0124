"""Aggregate an ncu report's source page by source line: executed warp-instructions and stall samples.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    try: line = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    try:
        agg[(cur, line)][0] += int(d.get('Instructions Executed') or 0); agg[(cur, line)][1] += int(d.get('# Samples') or 0)
    except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print('total warp-instructions', tot, 'samples', tots)
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items(): byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]): print('%-24s %12d %5.1f%%  samples %5.1f%%' % (f, v[0], 100 * v[0] / tot, 100 * v[1] / max(tots, 1)))
print('top lines by stall samples')
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]: print('%-20s %5d inst %10d %5.1f%%  samples %5.1f%%' % (f, l, v[0], 100 * v[0] / tot, 100 * v[1] / max(tots, 1)))
