"""Build libfreqair.so in-tree with nvcc for sm_100a (the .so is git-ignored but travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, 'libfreqair.so')
SOURCES = ['api.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'norm.cu', 'conv.cu', 'fft_band.cu', 'win_attn.cu', 'joint_attn.cu',
           'elementwise.cu', 'dcn.cu', 'datagen.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(ROOT, 'include', 'freqair.h'))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _newer(s, o) or os.path.getmtime(o) < hdr_time:
            cmd = [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o]
            if verbose:
                cmd.insert(1, '-Xptxas')
                cmd.insert(2, '-v')
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'--- nvcc failed on {src}\n{out}\n')
        elif verbose or out.strip():
            sys.stderr.write(f'--- {src}\n{out}\n')
    if failed:
        raise RuntimeError('nvcc compilation failed')
    if procs or not os.path.exists(LIB):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
