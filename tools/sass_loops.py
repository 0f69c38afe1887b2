"""Developer tool: static loop bodies of one kernel's SASS (backward branches) with their opcode mix.
usage: python tools/sass_loops.py <object or cubin> <substring of the mangled kernel name> [min body size]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
minsz = int(sys.argv[3]) if len(sys.argv) > 3 else 100
names = [l.split('Function : ')[1].strip() for l in subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout.splitlines() if 'Function :' in l]
fn = [n for n in names if pat in n]
assert len(fn) == 1, fn
sass = subprocess.run(['cuobjdump', '-sass', '-fun', fn[0], obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print('%d instructions' % len(ins))
def opc(t):
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', t); return m.group(2) if m else '?'
for a, t in ins:
    m = re.search(r'BRA(\.U)?\s+.*?(0x[0-9a-f]+)', t)
    if m and int(m.group(2), 16) < a:
        lo = int(m.group(2), 16); body = [x for x in ins if lo <= x[0] <= a]
        if len(body) < minsz: continue
        c = collections.Counter(opc(x[1]) for x in body)
        print('loop 0x%x..0x%x: %d instr: ' % (lo, a, len(body)) + ', '.join('%s %d' % kv for kv in c.most_common(18)))
