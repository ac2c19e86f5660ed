"""Dynamic instruction census of an ncu source-page CSV (SASS view): executed warp instructions per opcode
and per segment between block barriers (dev tool).  usage: ncu_sass_phases.py src.csv [points]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0
hdr = rows[1]
ia, isrc, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed')
ish, ishi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
ismp = hdr.index('Warp Stall Sampling (All Samples)')
byop = collections.Counter(); segs = []; cur = collections.Counter(); curw = 0; curs = 0; tot = 0; start = 0
wave_tot = wave_ideal = 0
for n, r in enumerate(rows[2:]):
    if len(r) <= iex: continue
    src = r[isrc].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else src.split()[0]
    ex = float(r[iex] or 0)
    base = op.split('.')[0]
    byop[base] += ex; tot += ex
    cur[base] += ex
    w = float(r[ish] or 0); wave_tot += w; wave_ideal += float(r[ishi] or 0); curw += w; curs += float(r[ismp] or 0)
    if base == 'BAR':
        segs.append((start, n, sum(cur.values()), curw, curs, cur)); cur = collections.Counter(); curw = 0; curs = 0; start = n + 1
segs.append((start, len(rows), sum(cur.values()), curw, curs, cur))
print('total warp instructions %.0f = %.0f per point; smem wavefronts %.0f per point (ideal %.0f)' % (tot, tot / npts, wave_tot / npts, wave_ideal / npts))
print('by opcode:')
for op, v in byop.most_common(28): print('  %-10s %6.2f%%  %8.0f /point' % (op, 100 * v / tot, v / npts))
print('segments between barriers (sass rows, warp-inst/point, smem wavefronts/point, stall samples %):')
ts = sum(s[4] for s in segs) or 1
for a, b, v, w, sm, c in segs:
    if v / tot < 0.004: continue
    top = ' '.join('%s:%.0f' % (k, x / npts) for k, x in c.most_common(6))
    print('  rows %5d-%5d  %7.0f  %6.0f  %5.1f%%   %s' % (a, b, v / npts, w / npts, 100 * sm / ts, top))
