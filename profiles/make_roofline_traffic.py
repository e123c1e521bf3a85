"""profiles/roofline_traffic.json (read by bench.py: roofline.traffic) from an ncu launch list that carries
dram__bytes_read.sum / dram__bytes_write.sum per launch.
The file is stamped with the sha of the convolution sources (bench.conv_source_sha): bench.py reports the traffic only while the
sources it runs are the ones the capture was taken on - run this script on the SAME tree that produced the launch list.
usage: python profiles/make_roofline_traffic.py <launches.csv> <steps_in_capture> <source note>"""
import collections
import csv
import json
import os
import re
import sys


def main(path, steps, note):
    lines = [l for l in open(path) if not l.startswith('==')]
    byt, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        if not name.startswith('conv_'):
            continue
        m, u = row['Metric Name'], row['Metric Unit']
        if m.startswith('dram__bytes'):
            byt[name] += float(row['Metric Value'].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        elif m == 'gpu__time_duration.sum':
            cnt[name] += 1
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = {'traffic': sum(byt.values()) / steps,
           'conv_source_sha': bench.conv_source_sha(),
           'unit': 'bytes per step (all tensor-core convolution launches of one training step: forward + dgrad + wgrad)',
           'launches_per_step': sum(cnt.values()) / steps,
           'per_kernel_bytes_per_step': {k: v / steps for k, v in sorted(byt.items(), key=lambda kv: -kv[1])},
           'source': note}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'roofline_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else os.path.basename(sys.argv[1]))
