"""Build libhopk.so (the sm_100a kernels + C ABI) in-tree with nvcc.

    python -m hop_b200.build            # or: python <pkg>/build.py

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libhopk.so')
SOURCES = ['api.cu', 'gwnet.cu', 'linear.cu', 'xattn.cu', 'xattn_tc.cu', 'gemm_tma.cu', 'gru.cu', 'glue.cu', 'bert.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, '..', 'include', 'hopk.h'), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, 'build', src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get('HOPK_NVCC_EXTRA', '').split(), '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {src}')
    with open(os.path.join(PKG, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-lcudart']
    subprocess.check_call(cmd)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
