"""Annotated SASS of an ncu report: per instruction the CUDA source line, executed
warp-instructions (share), average lanes and stall samples.
usage: python tools/ncu_annot.py report.ncu-rep kernel.cubin [first [last]]"""
import csv, io, re, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
line_of, cur = {}, ('?', 0)
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc = hdr.index('Address'), hdr.index('Source')
ii, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = int(rows[2][ia], 16)
tot_i = sum(int(r[ii]) for r in rows[2:]); tot_s = sum(int(r[isamp]) for r in rows[2:])
for n, r in enumerate(rows[2:]):
    if n < first or n > last:
        continue
    i, t, s = int(r[ii]), int(r[it]), int(r[isamp])
    f, l = line_of.get(int(r[ia], 16) - base, ('?', 0))
    print('%4d %-22s %9.3e %5.2f%% %5.1f %5.2f%%  %s' % (n, '%s:%d' % (f.replace('.cuh', ''), l), i, 100*i/tot_i, t/max(i, 1), 100*s/max(tot_s, 1), r[isrc]))
