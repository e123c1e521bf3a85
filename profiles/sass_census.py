"""Instruction census of libsaltunet.so (cuobjdump -sass): per kernel, how many tcgen05 / TMA / bulk-copy / TMEM instructions it
carries, plus the MMA issue loop of the row-halo convolution as an excerpt.
usage: python profiles/sass_census.py > profiles/r2_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'open-solution-salt-identification_b200', 'libsaltunet.so')
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'REDG', 'RED.', 'UTMAPF', 'STG.E.128',
        'STG.E.ENL2.256', 'LDS.128', 'LDG.E.128']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1); kernels[cur] = []
        elif cur is not None and re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
            kernels[cur].append(line)
    names = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# instruction census of %s (cuobjdump -sass, sm_100a)' % os.path.basename(LIB))
    print('# columns: ' + ' | '.join(KEYS) + ' | total instructions | kernel')
    tot = collections.Counter()
    for (mangled, lines), name in zip(kernels.items(), names):
        text = '\n'.join(lines)
        c = [len(re.findall(r'\b' + re.escape(k), text)) for k in KEYS]
        for k, v in zip(KEYS, c):
            tot[k] += v
        if any(c[:6]) or 'ring' in name:
            short = re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', '')).replace('void ', '')
            print(' | '.join('%4d' % v for v in c) + ' | %5d | %s' % (len(lines), short[:110]))
    print('# library totals: ' + ', '.join('%s %d' % (k, tot[k]) for k in KEYS))
    # excerpt: the MMA issue loop of the default row-halo convolution (N = 64, 2-CTA multicast clusters)
    for (mangled, lines), name in zip(kernels.items(), names):
        if 'conv_tc_rows_kernel<64, __nv_bfloat16, 2, false>' in name or ('conv_tc_rows_kernel<(int)64, __nv_bfloat16, (int)2, (bool)0>' in name):
            idx = [i for i, l in enumerate(lines) if 'UTCHMMA' in l]
            if idx:
                lo, hi = max(0, idx[0] - 12), min(len(lines), idx[min(len(idx) - 1, 11)] + 8)
                print('\n# excerpt: %s, first MMA group of the issue loop (lines %d-%d of %d)' % (re.sub(r'\(.*', '', name), lo, hi, len(lines)))
                print('\n'.join(lines[lo:hi]))
            break


if __name__ == '__main__':
    main()
