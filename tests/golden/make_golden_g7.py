"""Extracts golden set G7 -- the /cell_to mappings of a result file real DAMASK wrote for a job with 8 constituents per cell
(/root/reference/python/tests/resources/Result/4grains2x4x3_compressionY.hdf5: DAMASK_grid 3.0.0-alpha7-366, RGC homogenisation of
the deck 4grains2x4x3.material.yaml on the geometry 4grains2x4x3.vti, which is tests/golden/damask_written_4grains2x4x3.vti) --
into tests/golden/g7_multi_constituent_mapping.npz: the phase label of every constituent of every material entry of the deck, the
file's /cell_to/phase [24, 8] and /cell_to/homogenization [24] (label, entry), and the rows of each phase's datasets.
Read with solver_cpp_b200/hdf5_min.py (no h5py in this image; that reader is pinned on other DAMASK-written files, G4 / G5).
Run in the build container (the reference tree is not present on the GPU box):  python tests/golden/make_golden_g7.py"""
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
REF = '/root/reference/python/tests/resources/Result/'


def main():
    from solver_cpp_b200.hdf5_min import H5File
    f = H5File(REF + '4grains2x4x3_compressionY.hdf5')
    mat = yaml.safe_load(open(REF + '4grains2x4x3.material.yaml'))
    out = {}
    out['material_phase'] = np.array([[c['phase'] for c in m['constituents']] for m in mat['material']])        # [24, 8]
    out['material_v'] = np.array([[c['v'] for c in m['constituents']] for m in mat['material']])
    out['material_homogenization'] = np.array([m['homogenization'] for m in mat['material']])
    ph = f.read('/cell_to/phase')
    out['cell_to_phase_label'] = ph['label'].astype('S8'); out['cell_to_phase_entry'] = ph['entry']
    ho = f.read('/cell_to/homogenization')
    out['cell_to_homogenization_label'] = ho['label'].astype('S8'); out['cell_to_homogenization_entry'] = ho['entry']
    for p in ('A', 'B', 'C'):
        out['rows_' + p] = np.int64(f.read('/increment_10/phase/%s/mechanical/F' % p).shape[0])
    out['creator'] = np.array(str(f.attrs('/')['creator']))
    np.savez(os.path.join(HERE, 'g7_multi_constituent_mapping.npz'), **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == '__main__':
    main()
