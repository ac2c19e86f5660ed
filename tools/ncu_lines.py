"""Per-source-line instruction / stall-sample breakdown of an ncu source-page CSV (dev tool)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []
def fl(x):
    try: return float(x or 0)
    except Exception: return 0.0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if not r or r[0] == '' or hdr is None: continue
    try: ln = int(r[0])
    except Exception: continue
    d = dict(zip(hdr, r)); g = lambda k: fl(d.get(k, '0'))
    agg.append((cur, ln, g('Warp Stall Sampling (All Samples)'), g('Instructions Executed'), g('stall_long_sb'),
                g('stall_short_sb'), g('stall_mio'), g('stall_barrier'), g('stall_wait'), g('stall_math')))
tot = sum(a[2] for a in agg); toti = sum(a[3] for a in agg)
print('total samples', tot, 'inst', toti)
byfile = collections.defaultdict(lambda: [0, 0])
for a in agg: byfile[a[0]][0] += a[2]; byfile[a[0]][1] += a[3]
for f, (s, i) in byfile.items():
    if i / toti > 0.002: print('%-16s samples %5.1f%% inst %5.1f%%' % (f, 100 * s / tot, 100 * i / toti))
src = {}
import os
base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'thepayne_b200', 'csrc')
for f in [x for x in os.listdir(base) if os.path.isfile(os.path.join(base, x))]: src[f] = open(os.path.join(base, f)).read().split('\n')
agg.sort(key=lambda a: -a[2 if (len(sys.argv) > 3 and sys.argv[3] == "samples") else 3])
print('%-14s %4s %6s %6s %6s %6s %6s %6s %6s' % ('file', 'line', 'inst%', 'samp%', 'longsb', 'shrtsb', 'mio', 'bar', 'wait'))
for a in agg[:top]:
    line = src[a[0]][a[1] - 1].strip()[:72] if a[0] in src and a[1] - 1 < len(src[a[0]]) else ''
    print('%-14s %4d %6.2f %6.2f %6.0f %6.0f %6.0f %6.0f %6.0f  %s' % (a[0][:14], a[1], 100 * a[3] / toti, 100 * a[2] / tot, a[4], a[5], a[6], a[7], a[8], line))
