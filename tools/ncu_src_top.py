"""Summary of an `ncu --page source --csv --print-source sass` export: executed instructions and stall samples per
SASS instruction, bucketed by execution count (loop nesting level), and the most-sampled instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix['# Samples']]) for r in data); tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
print('total samples', tot_s, 'warp instructions', tot_i)
b = collections.OrderedDict()
for r in data:
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    key = float('%.2g' % n)
    e = b.setdefault(key, [0, 0, 0]); e[0] += 1; e[1] += s; e[2] += n
for k in sorted(b): print('  executed ~%-10g: %4d instrs, %5.1f%% of samples, %5.1f%% of instructions' % (k, b[k][0], 100.0 * b[k][1] / tot_s, 100.0 * b[k][2] / tot_i))
top = sorted(enumerate(data), key=lambda x: -int(x[1][ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for k, r in sorted(top): print(k, r[ix['Source']][:76].ljust(76), r[ix['# Samples']].rjust(7), r[ix['Instructions Executed']].rjust(10))
