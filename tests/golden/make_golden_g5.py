"""Golden set G5: the FRAME of a real DAMASK result file (12grains6x7x8_tensionY.hdf5, written by DAMASK_grid through
libhdf5): root / geometry attributes, the cell_to mappings, and F_e + O of increment 4 per phase -- what
solver_cpp_b200.dadf5 has to reproduce from the same input deck (tests/test_dadf5_cpu.py).  Also asserts that the strict
group-structure checker used on our own files accepts the files libhdf5 wrote.
Run in the build container:  python tests/golden/make_golden_g5.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from solver_cpp_b200.hdf5_min import H5File, check_group_structure          # noqa: E402

R = '/root/reference/python/tests/resources/Result/'


def main():
    for fn in ('6grains6x7x8_single_phase_tensionY', '12grains6x7x8_tensionY', '4grains2x4x3_compressionY'):
        assert check_group_structure(H5File(R + fn + '.hdf5')) > 100
    # the reading side (dadf5.ResultReader.place) on the real two-phase file reproduces golden set G4b
    from solver_cpp_b200 import dadf5
    rd, g4b = dadf5.ResultReader(R + '12grains6x7x8_tensionY.hdf5'), np.load(os.path.join(HERE, 'g4b_damask_run_two_phase.npz'))
    assert all(np.array_equal(rd.place(k, 4), g4b[k][0]) for k in ('F', 'P')) and rd.cells == (6, 7, 8)
    f = H5File(R + '12grains6x7x8_tensionY.hdf5')
    out = {}
    root, geo = f.attrs('/'), f.attrs('/geometry')
    out['version'] = np.array([root['DADF5_version_major'], root['DADF5_version_minor']])
    for k in ('cells', 'size', 'origin'):
        out[k] = geo[k]
    ph, ho = f.read('/cell_to/phase'), f.read('/cell_to/homogenization')
    out['phase_label'], out['phase_entry'] = ph['label'].reshape(-1), ph['entry'].reshape(-1)
    out['homog_label'], out['homog_entry'] = ho['label'].reshape(-1), ho['entry'].reshape(-1)
    out['t_inc4'] = np.array(f.attrs('/increment_4')['t/s'])
    for p in f.keys('/increment_4/phase'):
        for name in ('F_e', 'O'):
            out['%s_%s' % (name, p)] = f.read('/increment_4/phase/%s/mechanical/%s' % (p, name))
        out['lattice_' + p] = np.array(f.attrs('/increment_4/phase/%s/mechanical/O' % p)['lattice'].encode())
    np.savez_compressed(os.path.join(HERE, 'g5_dadf5_frame.npz'), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == '__main__':
    main()
