"""TEST INFRASTRUCTURE: ctypes front-end of the host twin (tests/host_twin/twin.cu)."""
import ctypes as C, os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))

def build():
    lib = os.path.join(HERE, '_build', 'libtwin.so')
    src = os.path.join(HERE, 'twin.cu')
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    deps = [src] + [os.path.join(HERE, '..', '..', 'solver_cpp_b200', 'csrc', f) for f in ('constitutive_dev.cuh', 'tangent_dev.cuh', 'constitutive.cuh', 'machine_dev.cuh', 'pheno_fast_dev.cuh', 'pheno_lanes_dev.cuh')]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(['nvcc', '-O2', '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-I', os.path.join(HERE, '..', '..', 'include'),
                               '-o', lib, src], stderr=subprocess.DEVNULL)
    return C.CDLL(lib)

def P(a):
    return a.ctypes.data_as(C.c_void_p)

def run_point(lib, phase, num, F0, Ft, Fp0, Lp0, Lp, S0, state0, dep, Dt, tangent=False):
    out = dict(P=np.zeros(9), Fp=np.zeros(9), Fe=np.zeros(9), S=np.zeros(6), state=np.zeros(max(1, len(state0))), cnt=np.zeros(6, dtype=np.uint32),
               Lp=np.array(Lp, float).copy(), dep=np.array(dep, float).copy(), dPdF=np.zeros(81))
    F0, Ft, Fp0, Lp0, S0, state0 = [np.ascontiguousarray(x, float) for x in (F0, Ft, Fp0, Lp0, S0, state0)]
    ok = lib.twin_point(C.byref(phase), C.byref(num), P(F0), P(Ft), P(Fp0), P(Lp0), P(out['Lp']), P(S0), P(state0), P(out['dep']), C.c_double(Dt),
                        P(out['P']), P(out['Fp']), P(out['Fe']), P(out['S']), P(out['state']), P(out['cnt']), P(out['dPdF']) if tangent else None)
    out['ok'] = bool(ok)
    return out
