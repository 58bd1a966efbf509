"""Per-instruction view of an ncu report: executed warp-instr, avg lanes, samples.
usage: python tools/ncu_sass.py report.ncu-rep [min_share_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc = hdr.index('Address'), hdr.index('Source')
ii, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
tot_i = sum(int(r[ii]) for r in rows[2:]); tot_s = sum(int(r[isamp]) for r in rows[2:])
tot_t = sum(int(r[it]) for r in rows[2:])
print('total warp-instr %.4e thread-instr %.4e lanes %.2f samples %d' % (tot_i, tot_t, tot_t/tot_i, tot_s))
for n, r in enumerate(rows[2:]):
    i, t, s = int(r[ii]), int(r[it]), int(r[isamp])
    print('%4d %6.2f%% %5.1f %6.2f%%  %s' % (n, 100*i/tot_i, t/max(i, 1), 100*s/max(tot_s, 1), r[isrc]))
