"""Developer tool: one-screen summary of an ncu report (first kernel): time, pipes, stalls, i-cache, opcode mix."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else None   # units = warp-steps in the launch
raw = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, r = raw[0], raw[2]
get = lambda k: r[hdr.index(k)] if k in hdr else 'n/a'
for w in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct',
          'sm__inst_issued.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sass__inst_executed_register_spilling', 'l1tex__t_sector_hit_rate.pct']:
    print('%-72s %s %s' % (w, get(w), raw[1][hdr.index(w)] if w in hdr else ''))
st = {h.split('stalled_')[1].split('_per_issue')[0]: float(r[i]) for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h}
print('stalls (warps per issue):', ', '.join('%s %.2f' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]; si, ei = h2.index('Source'), h2.index('Instructions Executed')
ops = collections.Counter(); tot = 0; static = 0
first_kernel_done = False
for row in src[2:]:
    if row and row[0] == 'Kernel Name':
        break
    if len(row) < len(h2) or not row[ei].isdigit(): continue
    if row[0].startswith('Kernel Name'): break
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', row[si]); op = m.group(2) if m else '?'
    n = int(row[ei]); ops[op] += n; tot += n; static += 1
print('static SASS lines', static, ' dynamic warp-inst', tot, (' per unit %.0f' % (tot / units)) if units else '')
print(', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in ops.most_common(16)))
if units:
    f = {k: ops.get(k, 0) / units for k in ('DFMA', 'DMUL', 'DADD')}
    print('FP64 arithmetic per unit: DFMA %.0f, DMUL %.0f, DADD %.0f -> executed FLOP per trajectory-step (DFMA = 2) %.0f' % (f['DFMA'], f['DMUL'], f['DADD'], 2 * f['DFMA'] + f['DMUL'] + f['DADD']))
