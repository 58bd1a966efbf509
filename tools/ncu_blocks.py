"""Block view of an ncu report: consecutive SASS instructions with the same execution count and
lane count are one block (start-end, instructions, share of warp-instr, lanes, stall samples).
A block whose per-instruction share is a multiple of its neighbours' runs several times per
loop trip - e.g. lanes that reach it at different times without reconverging.
usage: python tools/ncu_blocks.py report.ncu-rep [min_total_share_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc = hdr.index('Source')
ii, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
tot_i = sum(int(r[ii]) for r in rows[2:]); tot_s = sum(int(r[isamp]) for r in rows[2:])
tot_t = sum(int(r[it]) for r in rows[2:])
print('total warp-instr %.4e thread-instr %.4e lanes %.2f samples %d' % (tot_i, tot_t, tot_t/tot_i, tot_s))
blocks = []
for n, r in enumerate(rows[2:]):
    i, t, s = int(r[ii]), int(r[it]), int(r[isamp])
    share, lanes = 100*i/tot_i, t/max(i, 1)
    b = blocks[-1] if blocks else None
    if b and abs(b['share'] - share) < 0.011 and abs(b['lanes'] - lanes) < 0.3:
        b['end'] = n; b['n'] += 1; b['samp'] += s
    else:
        blocks.append(dict(start=n, end=n, share=share, lanes=lanes, n=1, samp=s, first=r[isrc][:48]))
for b in blocks:
    if b['share']*b['n'] >= floor:
        print('%4d-%4d n=%3d  %.2f%%/instr  total %5.2f%%  lanes %5.1f  samples %5.2f%%  %s' % (
            b['start'], b['end'], b['n'], b['share'], b['share']*b['n'], b['lanes'],
            100*b['samp']/max(tot_s, 1), b['first']))
