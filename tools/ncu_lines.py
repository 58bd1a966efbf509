"""Aggregate an ncu SASS source page by CUDA source line.

usage: python tools/ncu_lines.py report.ncu-rep kernel.cubin [top]
NVRTC kernels carry -lineinfo but ncu cannot import the in-memory sources, so the
SASS page is joined with nvdisasm's line markers by instruction offset.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
line_of = {}
cur = ('?', 0)
inl = ''
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ii, it, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), \
    hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = int(rows[2][ia], 16)
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    off = int(r[ia], 16) - base
    key = line_of.get(off, (('?', 0), ''))[0]
    vals = (int(r[ii]), int(r[it]), int(r[isamp]))
    for k in range(3):
        agg[key][k] += vals[k]
        tot[k] += vals[k]
print('total warp-instr %.3e thread-instr %.3e avg lanes %.2f samples %d' % (
    tot[0], tot[1], tot[1]/max(tot[0], 1), tot[2]))
print('%-28s %10s %7s %7s %7s' % ('file:line', 'warp-instr', 'share', 'lanes', 'samp%'))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-28s %10.3e %6.2f%% %7.2f %6.2f%%' % ('%s:%d' % key, v[0], 100*v[0]/tot[0],
          v[1]/max(v[0], 1), 100*v[2]/max(tot[2], 1)))
