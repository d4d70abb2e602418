#!/usr/bin/env python3
"""per-source-line instruction counts and lane utilisation from a .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_source_lines.py file.ncu-rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
blocks, cur = [], []
fname = ''
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == 'File Path':
        fname = row[1].split('/')[-1]
    elif row and row[0] == 'Function Name':
        cur = []
        blocks.append((row[1], fname, cur))
    else:
        cur.append(row)
# one block per (launch, file); keep the blocks of the requested launch = same function name group
launches = [b for b in blocks if b[1].endswith('.cuh') or b[1].endswith('.cu')]
name, fname, rows = launches[which]
hdr = None
recs = []
for r in rows:
    if r and r[0] == 'Line No':
        hdr = ['Line No', 'Source', 'Address', 'Sass'] + r[4:]
    elif hdr and len(r) == len(hdr) and r[2] == '-':
        d = dict(zip(hdr, r))
        try:
            ie = int(d['Instructions Executed'])
            te = int(d['Thread Instructions Executed'])
        except (KeyError, ValueError):
            continue
        if ie:
            recs.append((ie, te, fname, d['Line No'], d['Source'].strip()[:110], d.get('Warp Stall Sampling (All Samples)', '')))
tot_i = sum(r[0] for r in recs)
tot_t = sum(r[1] for r in recs)
print('kernel:', name[:90])
print('total warp-inst %.3g  thread-inst %.3g  lanes/inst %.2f' % (tot_i, tot_t, tot_t / max(tot_i, 1)))
print('%7s %6s %6s %8s  %s' % ('inst%', 'lanes', 'stall', 'line', 'source'))
for ie, te, f, ln, src, st in sorted(recs, reverse=True)[:top]:
    print('%6.2f%% %6.2f %6s %8s  %s' % (100.0 * ie / tot_i, te / ie, st, f[:8] + ':' + ln, src))
