"""ORACLE (test infrastructure only). Builds oracle/_build/liboracle_constitutive.so from
oracle/constitutive.c with gcc (strict IEEE: -ffp-contract=off, no fast-math), OpenMP over cells."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD, 'liboracle_constitutive.so')


def build(force=False):
    from . import phase_layout
    os.makedirs(BUILD, exist_ok=True)
    hdr = os.path.join(HERE, '_gen_layout.h')
    txt = phase_layout.header_text()
    if not os.path.exists(hdr) or open(hdr).read() != txt:
        with open(hdr, 'w') as f:
            f.write(txt)
    src = os.path.join(HERE, 'constitutive.c')
    deps = [src, hdr, os.path.join(HERE, 'phase_layout.py')]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        return LIB
    cmd = ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-std=c11', '-shared', '-fPIC',
           '-Wall', '-o', LIB, src, '-lm']
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
