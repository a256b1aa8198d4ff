#!/usr/bin/env python
"""Digest of the CSV exports made by tools/gpu_ncu_export.sh: headline metrics per captured launch and, for one launch,
the stall-reason and opcode mix of the per-instruction samples.   usage: ncu_digest.py <tag> [launch index] [hot lines]"""
import csv, gzip, sys
from collections import Counter
tag = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else -1; nhot = int(sys.argv[3]) if len(sys.argv) > 3 else 0
base = "gpurun_out/ncu/" + tag if "/" not in tag else tag
rows = list(csv.reader(open(base + ".raw.csv")))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.max',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg']
ki = hdr.index('Kernel Name')
for n, r in enumerate(rows[2:]):
    print("[%d]" % n, r[ki][-70:])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-72s %-10s %s" % (w, units[i], "  ".join(r[i][:14] for r in rows[2:])))
if which < 0:
    sys.exit()
src = list(csv.reader(gzip.open(base + ".src.csv.gz", "rt")))
ks = []; cur = None
for r in src:
    if r and r[0] == "Kernel Name": cur = {'name': r[1], 'rows': []}; ks.append(cur); continue
    if r and r[0] == "Address": cur['hdr'] = r; continue
    if cur is not None and len(r) > 10: cur['rows'].append(r)
k = ks[which]; h = k['hdr']
ns = h.index('# Samples'); ie = h.index('Instructions Executed')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
S = sum(int(r[ns]) for r in k['rows']); I = sum(int(r[ie]) for r in k['rows'])
print(k['name'][-80:], 'samples', S, 'warp-instr', I, 'static', len(k['rows']))
tot = Counter()
for r in k['rows']:
    for i in stall: tot[h[i]] += int(r[i])
print({a: round(100 * b / S, 1) for a, b in tot.most_common() if b * 200 > S})
c = Counter(); cs = Counter()
for r in k['rows']:
    t = r[1].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[ie]); cs[op] += int(r[ns])
print("  ".join("%s %.1f%%/%.1f%%" % (op, 100 * v / I, 100 * cs[op] / S) for op, v in c.most_common(16)), "(instr / samples)")
if nhot:
    top = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][ns]))[:nhot]
    for i in sorted(top):
        r = k['rows'][i]
        why = max(stall, key=lambda j: int(r[j]))
        print("%5d %7s %9s  %-70s %s" % (i, r[ns], r[ie], r[1][:70], h[why]))
