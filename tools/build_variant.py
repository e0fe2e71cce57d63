"""Builds a tuning variant of the C-ABI library: same sources, extra -D flags for csrc/pheno_fast.cu (the per-point kernel).
  python tools/build_variant.py NAME [-DFLAG=V ...]     ->  tools/_build/libdmk_NAME.so   (use with DMK_LIB_PATH=...)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from solver_cpp_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
B.build()                                        # objects of the unchanged translation units
out_dir = os.path.join(ROOT, 'tools', '_build')
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, 'pheno_fast_%s.o' % name)
k = B.FLAGS.index('-Wall')
base = B.FLAGS[:k - 1] + B.FLAGS[k + 1:]                 # without the pair -Xcompiler -Wall
subprocess.run([B.NVCC] + base + ['-w'] + flags + ['-x', 'cu', '-c', os.path.join(B.CSRC, 'pheno_fast.cu'), '-o', obj], check=True)
objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith('.o') and not f.startswith('pheno_fast.cu')]
lib = os.path.join(out_dir, 'libdmk_%s.so' % name)
subprocess.run([B.NVCC, '-shared', '-o', lib] + objs + [obj, '-lcudart', '-lz'], check=True)
print(lib)
