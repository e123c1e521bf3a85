"""Build libsaltunet.so (sm_100a) in-tree with nvcc.  Usage: python build.py [--force] [-v]."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libsaltunet.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False, timing=False):
    """timing=True: a second library, libsaltunet_timing.so, compiled with -DSALT_TC_TIMING (pipeline-stall counters in the row-halo
    convolution, profiles/rows_timing.py); the shipped libsaltunet.so never carries them."""
    global OBJ, LIB, FLAGS
    if timing:
        OBJ, LIB = os.path.join(HERE, 'build_timing'), os.path.join(HERE, 'libsaltunet_timing.so')
        FLAGS = FLAGS + ['-DSALT_TC_TIMING']
    os.makedirs(OBJ, exist_ok=True)
    dep_t = _newest_dep()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= dep_t:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + '.o')
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_t:
            return obj, ''
        cmd = [NVCC] + ARCH + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, _sources()))
    log = '\n'.join(s for _, s in results)
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + [o for o, _ in results] + ['-lcuda', '-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, timing='--timing' in sys.argv))
