"""Build libqce_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libqce_b200.so')
EXTRA = os.environ.get('QCE_NVCC_EXTRA', '').split()
NVCC_FLAGS = EXTRA + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-pthread', '-shared', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into ``libqce_b200.so``; returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libqce_b200.so')
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, 'build'), exist_ok=True)
    for src in sources():
        obj = os.path.join(PKG, 'build', os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != '-shared'] + ['-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    tmp = LIB + '.tmp'
    out = subprocess.run([nvcc, '-shared', '-o', tmp] + objs + ['-lcudart', '-lpthread'], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError('link failed:\n' + out.stdout)
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
