"""Golden set G9: a third real DAMASK run shipped with the reference as a DADF5 result file
(python/tests/resources/Result/2phase_irregularGrid_tensionX_2phase_irregularGrid.material.hdf5: DAMASK_grid 3.0.0-alpha7-880,
13 x 8 x 10 cells -- an odd, non-power-of-two grid --, 18 grains of an fcc phenopowerlaw matrix and an elastic bcc precipitate,
tensionX, 40 increments), extracted with the repo's own HDF5 reader into tests/golden/g9_damask_run_irregular.npz: per-cell F and P
and the matrix's slip resistances at increment 40.  The run's inputs are taken from the file's own /setup group and written next
to it (g9_irregular.vti through solver_cpp_b200/vti.py, g9_irregular.material.yaml, g9_irregular.load.yaml through PyYAML).
Run in the build container:  python tests/golden/make_golden_g9.py"""
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
SRC = '/root/reference/python/tests/resources/Result/2phase_irregularGrid_tensionX_2phase_irregularGrid.material.hdf5'


def main():
    from solver_cpp_b200.hdf5_min import H5File
    from solver_cpp_b200 import vti, dadf5
    f = H5File(SRC)
    with tempfile.NamedTemporaryFile(suffix='.vti') as tmp:
        tmp.write(f.read('/setup/2phase_irregularGrid.vti')[0]); tmp.flush()
        geo = vti.read_vti(tmp.name)
    vti.write_vti(os.path.join(HERE, 'g9_irregular.vti'), geo['material'], geo['size'], geo['origin'])
    for src, dst in (('2phase_irregularGrid.material.yaml', 'g9_irregular.material.yaml'), ('tensionX.yaml', 'g9_irregular.load.yaml')):
        d = yaml.safe_load(f.read('/setup/' + src)[0].decode())
        with open(os.path.join(HERE, dst), 'w') as out:
            out.write('# written by tests/golden/make_golden_g9.py from /setup/%s of the reference\'s result file\n' % src)
            yaml.safe_dump(d, out, default_flow_style=None, sort_keys=False)
    r = dadf5.ResultReader(SRC)
    assert r.increments == [0, 40] and r.phases == ['Matrix', 'Precipitate']
    out = {'F': r.place('F', 40), 'P': r.place('P', 40), 'xi_sl': r.place('xi_sl', 40), 't': np.float64(r.times[-1]),
           'phase_of_cell': np.array([str(p) for p in r.phase_label]), 'creator': np.array(str(f.attrs('/')['creator']))}
    np.savez_compressed(os.path.join(HERE, 'g9_damask_run_irregular.npz'), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
