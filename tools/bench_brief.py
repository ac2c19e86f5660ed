"""Dev tool: the few numbers of a bench.py JSON line one looks at between experiments.  usage: bench_brief.py file [label]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
lab = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
oc = d.get('other_configs') or {}
print('%s: value %.4g %s, %.4f ms/step, tail %.4f ms, e2e %.4g, c4 %s, c3 %s, c5 %s' % (
    lab, d['value'], d['unit'], d['ms_per_step'], d['roofline'].get('ms_per_launch', float('nan')), d['e2e']['value'],
    ('%.4g' % oc['c4_monolithic']['value']) if 'c4_monolithic' in oc else '-',
    ('%.4g' % oc['c3_joint']['value']) if 'c3_joint' in oc else '-',
    ('%.4g' % d['c5_sweep']['value']) if d.get('c5_sweep') else '-'))
