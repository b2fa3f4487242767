"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    t, n = collections.Counter(), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[iu].strip(), 1e-6)
        nm = r[ik].replace('void ', '').replace('ehm::', '').split('<')[0].split('(')[0]
        t[nm] += v * scale
        n[nm] += 1
    tot = sum(t.values())
    print('# %s: %d launches, %.3f ms of kernel time (cold-cache, serialised by ncu)' % (path, sum(n.values()), tot))
    for k, v in t.most_common():
        print('%-22s %5d launches %12.3f ms %6.2f %%' % (k, n[k], v, 100 * v / tot))


if __name__ == '__main__':
    main(sys.argv[1])
