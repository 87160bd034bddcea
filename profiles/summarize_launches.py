"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path, detail=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void |cgg::|<unnamed>::|unnamed>::', '', name)[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
        if detail and detail in row['Kernel Name']:
            print('   id %s grid %s  %.1f us' % (row['ID'], row['Grid Size'], v))
    print('launches %d, total %.1f us' % (n, tot))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-62s n=%4d %10.1f us %5.1f%%' % (k, c, t, 100 * t / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
