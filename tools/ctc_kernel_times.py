"""Per-kernel device times of the CTC launches from an ncu launch list of tools/microbench.py --what ctc
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv).  Prints the last
iteration of every case: scan | lattice | recursion | gradient."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid], {'k': r[ik]})[r[im]] = float(r[iv].replace(',', ''))
seq = [v for v in d.values() if 'ctc' in v['k']]
per_case = int(sys.argv[2]) if len(sys.argv) > 2 else 4  # launches of one case = 4 kernels x (warmup 3 + iters)
for i in range(0, len(seq), 4):
    if (i // 4) % per_case == per_case - 1:
        g = seq[i:i + 4]
        print(' | '.join('%s %.1fus r%.0f w%.0fMB' % (x['k'].replace('vocr::', '').replace('void ', '')[:14],
                                                       x['gpu__time_duration.sum'] / 1e3, x['dram__bytes_read.sum'] / 1e6,
                                                       x['dram__bytes_write.sum'] / 1e6) for x in g),
              '| total %.1fus' % (sum(x['gpu__time_duration.sum'] for x in g) / 1e3))
