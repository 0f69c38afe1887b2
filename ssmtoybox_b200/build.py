"""Build libssmb200.so (the C-ABI CUDA library, include/ssm_b200.h) in-tree with nvcc for sm_100a.

    python -m ssmtoybox_b200.build [--force]

Objects go to build/, the shared library to ssmtoybox_b200/lib/libssmb200.so (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(ROOT, 'build', 'obj')
LIB = os.path.join(PKG, 'lib', 'libssmb200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--extended-lambda',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hs.append(os.path.join(ROOT, 'include', 'ssm_b200.h'))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    s = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(s), _headers_mtime()):
        return obj, ''
    r = subprocess.run([NVCC] + FLAGS + ['-c', s, '-o', obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}'.format(src, r.stderr[-4000:]))
    with open(obj + '.ptxas.log', 'w') as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    newest = max(os.path.getmtime(o) for o, _ in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        r = subprocess.run([NVCC, '-shared', '-o', LIB] + [o for o, _ in objs] + ['-lcudart'],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stderr[-4000:])
    if verbose:
        for _, log in objs:
            sys.stderr.write(log)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
