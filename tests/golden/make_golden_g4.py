"""Golden set G4: per-point F and P of a REAL DAMASK run (DAMASK_grid 3.0.0-alpha7-366, 6x7x8 cells, 6 grains, fcc
phenopowerlaw, tensionY, 40 increments) shipped with the reference as a DADF5 result file, extracted with the repo's
own minimal HDF5 reader into tests/golden/g4_damask_run.npz.  The run's inputs are the three fixtures next to it
(6grains6x7x8_single_phase.vti / .material.yaml, 6grains6x7x8_tensionY.load.yaml: copies of the reference's files).
Run in the build container:  python tests/golden/make_golden_g4.py"""
import importlib.util
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/python/tests/resources/Result/6grains6x7x8_single_phase_tensionY.hdf5'
spec = importlib.util.spec_from_file_location('hdf5_min', os.path.join(HERE, '..', '..', 'solver_cpp_b200', 'hdf5_min.py'))
h5 = importlib.util.module_from_spec(spec)
spec.loader.exec_module(h5)


def main():
    f = h5.H5File(SRC)
    raw = f.read('/cell_to/phase').tobytes()
    rec = len(raw) // 336
    entry = np.array([int.from_bytes(raw[rec * i + rec - 8:rec * i + rec], 'little') for i in range(336)])
    assert (entry == np.arange(336)).all()                    # phase entries are stored in cell order (x fastest)
    incs = [4, 8, 20, 40]
    out = {'increments': np.array(incs)}
    for name in ('F', 'P'):
        out[name] = np.stack([f.read('/increment_%d/phase/pheno_fcc/mechanical/%s' % (i, name)) for i in incs])
    for name in ('u_p', 'u_n'):                               # cell-centre / nodal displacements (update_coords)
        out[name] = np.stack([f.read('/increment_%d/geometry/%s' % (i, name)) for i in incs])
    out['xi_sl_final'] = f.read('/increment_40/phase/pheno_fcc/mechanical/xi_sl')
    np.savez_compressed(os.path.join(HERE, 'g4_damask_run.npz'), **out)
    print({k: v.shape for k, v in out.items()})


def main_two_phase():
    """12grains6x7x8_tensionY.hdf5: the same load case on a 12-grain microstructure with an fcc and a bcc phase"""
    f = h5.H5File('/root/reference/python/tests/resources/Result/12grains6x7x8_tensionY.hdf5')
    raw = f.read('/cell_to/phase').tobytes()
    rec = len(raw) // 336
    label = [raw[rec * i:rec * i + rec - 8].split(b'\0')[0].decode() for i in range(336)]
    entry = np.array([int.from_bytes(raw[rec * i + rec - 8:rec * i + rec], 'little') for i in range(336)])
    incs = [4, 20, 40]
    out = {'increments': np.array(incs)}
    for name in ('F', 'P'):
        per_inc = []
        for i in incs:
            a = np.zeros((336, 3, 3))
            for ph in sorted(set(label)):
                d = f.read('/increment_%d/phase/%s/mechanical/%s' % (i, ph, name))
                sel = np.array([l == ph for l in label])
                a[sel] = d[entry[sel]]
            per_inc.append(a)
        out[name] = np.stack(per_inc)
    np.savez_compressed(os.path.join(HERE, 'g4b_damask_run_two_phase.npz'), **out)
    print({k: v.shape for k, v in out.items()}, sorted(set(label)))


def input_decks():
    """the input files of the two runs, re-written from the reference's own (geometry through solver_cpp_b200/vti.py, YAML through
    PyYAML): same content, this repo's writers"""
    import yaml
    spec2 = importlib.util.spec_from_file_location('dmk_vti', os.path.join(HERE, '..', '..', 'solver_cpp_b200', 'vti.py'))
    vti = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(vti)
    R = '/root/reference/python/tests/resources/Result/'
    for src, dst in (('6grains6x7x8_single_phase.vti', '6grains6x7x8_single_phase.vti'), ('12grains6x7x8.vti', '12grains6x7x8.vti')):
        g = vti.read_vti(R + src)
        vti.write_vti(os.path.join(HERE, dst), g['material'], g['size'], g['origin'])
    for src, dst in (('6grains6x7x8_single_phase.material.yaml', '6grains6x7x8_single_phase.material.yaml'),
                     ('12grains6x7x8.material.yaml', '12grains6x7x8.material.yaml'), ('tensionY.yaml', '6grains6x7x8_tensionY.load.yaml')):
        d = yaml.safe_load(open(R + src))
        with open(os.path.join(HERE, dst), 'w') as f:
            f.write('# written by tests/golden/make_golden_g4.py from the reference\'s python/tests/resources/Result/%s\n' % src)
            yaml.safe_dump(d, f, default_flow_style=None, sort_keys=False)


if __name__ == '__main__':
    input_decks()
    main()
    main_two_phase()
