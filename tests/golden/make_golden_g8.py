"""Extracts golden set G8 -- a result file the reference's author wrote with the C++ fork itself (creator " 3.0.0-alpha8",
/root/reference/2x2x2_tensionX_material_numerics.hdf5: increment 1, t = 1 s, of examples/grid/tensionX-like loading on a 2 x 2 x 2
grid whose cells all carry the first material entry of examples/grid/material.yaml, i.e. one cube-oriented Al crystal; numerics
N_iter_min 4 / N_iter_max 40) -- into tests/golden/g8_reference_run_2x2x2.npz: per-cell F, F_e, F_p, L_p, P, O, xi_sl of the converged
increment (identical in all 8 cells), the time, and the material / numerics decks the file carries under /setup.
Read with solver_cpp_b200/hdf5_min.py.  Run in the build container:  python tests/golden/make_golden_g8.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
REF = '/root/reference/2x2x2_tensionX_material_numerics.hdf5'


def main():
    from solver_cpp_b200.hdf5_min import H5File
    f = H5File(REF)
    out = {'t': np.float64(f.attrs('/increment_1')['t/s']), 'creator': np.array(str(f.attrs('/')['creator']))}
    for k in ('F', 'F_e', 'F_p', 'L_p', 'P', 'O', 'xi_sl'):
        out[k] = f.read('/increment_1/phase/Aluminum/mechanical/' + k)
    out['material_yaml'] = np.array(f.read('/setup/material.yaml')[0].decode())
    out['numerics_yaml'] = np.array(f.read('/setup/numerics.yaml')[0].decode())
    np.savez(os.path.join(HERE, 'g8_reference_run_2x2x2.npz'), **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == '__main__':
    main()
