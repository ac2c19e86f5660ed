"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (committed under profiles/).

    python tools/ncu_launches.py gpurun_out/launches.csv profiles/r01_launches_bench.txt "<command line>"
"""
import collections, csv, sys


def main(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[4], [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
    tot = sum(a[1] for a in agg.values())
    out = ['# ncu --metrics gpu__time_duration.sum --clock-control none  %s' % cmd,
           '# per-launch times are cold-cache and serialised: compare SHARES of the step, not absolutes']
    for k, (n, t) in agg.items():
        out.append('%-72s n=%3d total %8.3f ms  avg %.4f ms  share %5.1f%%' % (k[:72], n, t, t / n, 100 * t / tot))
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
