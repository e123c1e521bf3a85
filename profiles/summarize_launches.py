"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list into a per-kernel table.
usage: python profiles/summarize_launches.py <launches.csv> <steps_in_capture> > summary.md"""
import collections
import csv
import re
import sys


def main(path, steps):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt, byt = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        v, u, m = float(row['Metric Value'].replace(',', '')), row['Metric Unit'], row['Metric Name']
        if m == 'gpu__time_duration.sum':
            tot[name] += v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u == 's' else v
            cnt[name] += 1
        elif m.startswith('dram__bytes'):
            byt[name] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    total = sum(tot.values())
    print('| kernel | ms / step | share | launches / step | DRAM MB / step |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('| `%s` | %.3f | %.1f %% | %.1f | %s |' % (k[:70], v / steps, 100 * v / total, cnt[k] / steps,
                                                        ('%.0f' % (byt[k] / steps / 1e6)) if k in byt else '-'))
    print('| **total** | %.3f | | %.0f | |' % (total / steps, sum(cnt.values()) / steps))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]))
