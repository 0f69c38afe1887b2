"""Developer tool: per-kernel totals and shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_list.py <launches.csv> <out.md> "<title line>" "<bench shares note>" """
import collections, csv, io, sys
src, dst, title, note = sys.argv[1:5]
lines = open(src).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in csv.DictReader(io.StringIO('\n'.join(lines[start:]))):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v, u = float(r['Metric Value'].replace(',', '')), r['Metric Unit']
    ms = v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v if u.startswith('m') else v * 1e3
    tot[r['Kernel Name']] += ms
    cnt[r['Kernel Name']] += 1
T = sum(tot.values())
out = ['# ' + title, '', note, '', '| kernel | launches | total ms | share |', '|---|---:|---:|---:|']
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    out.append('| `%s` | %d | %.2f | %.1f %% |' % (k[:110], cnt[k], v, 100 * v / T))
open(dst, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:10]))
