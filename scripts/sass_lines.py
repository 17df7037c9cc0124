"""Per-source-line view of one kernel's SASS (nvdisasm -g of the built library): instruction counts, spills (STL/LDL),
selected opcodes.  usage: python scripts/sass_lines.py <kernel-name-substring> [opcode-regex]"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, 'mycroft_precise_b200', 'csrc', 'libprecise_b200.so')
name = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r'\b(STL|LDL)\b')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = None
tot = collections.Counter(); hit = collections.Counter()
for ln in txt:
    if ln.startswith('//---------------------'):
        inside = ('.text.' in ln) and (name in ln); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', ln):
        tot[cur] += 1
        if pat.search(ln): hit[cur] += 1
print('instructions:', sum(tot.values()), 'matching:', sum(hit.values()))
for k, v in sorted(hit.items(), key=lambda x: -x[1])[:40]: print(k, v, '/', tot[k])
