"""CPU tests of the library's own YAML reader and C++ material / numerics set-up (csrc/yaml_min.cpp, csrc/init_dropin.cpp), which
the init symbols of the drop-in boundary use (f_materialpoint_initDamask): against PyYAML and against the Python host's
solver_cpp_b200.material on the same files."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, 'tests', 'abi', '_build')


@pytest.fixture(scope='module')
def yaml_dump():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, 'yaml_dump')
    src = [os.path.join(ROOT, 'tests', 'abi', 'yaml_dump.cpp'), os.path.join(ROOT, 'solver_cpp_b200', 'csrc', 'yaml_min.cpp')]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in src):
        subprocess.run(['g++', '-std=c++17', '-O1', '-o', exe] + src, check=True)
    return exe


def _same(a, b, path=''):
    if isinstance(b, dict):
        assert isinstance(a, dict) and [str(k) for k in b] == list(a), path
        for k in b:
            _same(a[str(k)], b[k], path + '/' + str(k))
    elif isinstance(b, list):
        assert isinstance(a, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, '%s[%d]' % (path, i))
    elif b is None:
        assert a is None, path
    elif isinstance(b, bool):
        assert a.lower() in (('true', 'yes', 'on') if b else ('false', 'no', 'off')), path
    elif isinstance(b, (int, float)):
        assert float(a) == float(b), (path, a, b)
    else:
        assert a == str(b), (path, a, b)


TRICKY = """
# comment line
---
solver: {mechanical: spectral_basic}      # flow mapping, trailing comment
loadstep:
  - boundary_conditions:
      mechanical:
        dot_F: [[1.0e-3, 0, 0],
                [0, x, 0],
                [0, 0, x]]
        P: [ [x, x, x], [x, 0, x], [x, x, 0] ]
    discretization: {t: 5, N: 10}
    f_out: 4
  -   boundary_conditions: {mechanical: {dot_F: [[0, 0, 0], [0, 0, 0], [0, 0, 0]], P: [[x, x, x], [x, x, x], [x, x, x]]}}
      discretization:
        t: 1.5
        N: 3
phase:
  "A phase":
    lattice: cF
    'C_11,T^2': -9.790e3
    references:
      - first line,
        continued line
      - https://doi.org/10.1000/x#y
    empty:
list_at_key_indent:
- 1
- two
matrix_in_block_style:
- - 1.0e-3
  - x
- - -2
  - - nested
    - 4
"""


def test_reader_matches_pyyaml(yaml_dump, golden_dir, tmp_path):
    files = [os.path.join(golden_dir, f) for f in sorted(os.listdir(golden_dir)) if f.endswith('.yaml')]
    p = tmp_path / 'tricky.yaml'
    p.write_text(TRICKY)
    files.append(str(p))
    assert len(files) >= 5
    for f in files:
        r = subprocess.run([yaml_dump, f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stderr)
        _same(json.loads(r.stdout), yaml.safe_load(open(f)), f)


def test_reader_rejects_what_it_does_not_understand(yaml_dump, tmp_path):
    for bad in ('a: [1, 2\n', 'a:\n\t- 1\n', 'a: 1\n  b: 2\n', 'just a line\nanother: 1\n'):
        p = tmp_path / 'bad.yaml'
        p.write_text(bad)
        r = subprocess.run([yaml_dump, str(p)], capture_output=True, text=True)
        assert r.returncode == 1 and 'YAML' in r.stderr, (bad, r.stdout, r.stderr)


def test_cpp_material_setup_equals_python_host(tmp_path, golden_dir):
    """dmk_material_from_yaml (what f_materialpoint_initDamask uses) against solver_cpp_b200.material on the same material.yaml:
    every phase table byte for byte, phase index and orientation matrix of every material entry"""
    from solver_cpp_b200 import lib as L, material as M, workloads as W
    lib = L.load()
    decks = []
    for phase in ('Al_pheno', 'Cu_pheno', 'Al_iso', 'Al_elastic', 'TWIP', 'W_tungsten', 'Steel_kine'):
        material, _, _ = W.polycrystal((4, 4, 4), 5, phase, 1)
        p = tmp_path / ('material_%s.yaml' % phase)
        p.write_text(yaml.safe_dump(material))
        decks.append((str(p), material))
    for name in ('material_20grains_Al.yaml', '12grains6x7x8.material.yaml', '6grains6x7x8_single_phase.material.yaml', 'g9_irregular.material.yaml'):
        f = os.path.join(golden_dir, name)
        if os.path.exists(f):
            decks.append((f, yaml.safe_load(open(f))))
    assert len(decks) >= 11
    for path, material in decks:
        n_mat = len(material['material'])
        phases = (L.Phase * L.MAX_PHASES)()
        nph, nm = C.c_int(), C.c_int()
        ph_of = np.zeros(n_mat, dtype=np.int32)
        O = np.zeros((n_mat, 9))
        rc = lib.dmk_material_from_yaml(path.encode(), C.c_double(M.T_ROOM), phases, C.c_int(L.MAX_PHASES), C.byref(nph),
                                        ph_of.ctypes.data_as(C.c_void_p), O.ctypes.data_as(C.c_void_p), C.c_int(n_mat), C.byref(nm))
        assert rc == 0, lib.dmk_last_error()
        ref = M.Material(material, np.arange(n_mat))
        assert nph.value == len(ref.phases) and nm.value == n_mat
        for k in range(nph.value):
            a, b = bytes(phases[k]), bytes(ref.phases[k])
            if a != b:                                         # name the first field that differs
                for fname, _ in L.Phase._fields_:
                    x, y = np.ctypeslib.as_array(getattr(phases[k], fname)) if hasattr(getattr(phases[k], fname), '_length_') else getattr(phases[k], fname), \
                        np.ctypeslib.as_array(getattr(ref.phases[k], fname)) if hasattr(getattr(ref.phases[k], fname), '_length_') else getattr(ref.phases[k], fname)
                    assert np.array_equal(np.asarray(x), np.asarray(y)), (path, k, fname)
            assert a == b
        assert np.array_equal(ph_of, ref.phase_of_cell)
        assert np.array_equal(O, ref.O_of_cell)


def test_cpp_numerics_equals_python_host(tmp_path):
    from solver_cpp_b200 import lib as L, job
    lib = L.load()
    p = tmp_path / 'numerics.yaml'
    p.write_text(yaml.safe_dump({'solver': {'grid': {'N_cutback_max': 5, 'mechanical': {'N_iter_min': 4, 'N_iter_max': 40, 'eps_abs_div(P)': 1e-2,
                                                                                         'eps_rel_div(P)': 1e-1, 'update_gamma': True},
                                                     'FFT': {'derivative': 'FWBW_difference', 'divergence_correction': 'size'}}},
                                 'phase': {'mechanical': {'r_cutback': 0.5, 'plastic': {'integrator_state': 'RK4', 'N_iter_Lp_max': 25}}}}))
    for path in (None, str(p)):
        num = L.Numerics()
        der, div = C.c_int(), C.c_int()
        rc = lib.dmk_numerics_from_yaml(path.encode() if path else None, C.byref(num), C.byref(der), C.byref(div))
        assert rc == 0, lib.dmk_last_error()
        ref, rder, rdiv, rcut, rpol = job.read_numerics(path)
        assert bytes(num) == bytes(ref) and (der.value, div.value) == (rder, rdiv)
        gn = L.GridNumerics()
        assert lib.dmk_grid_numerics_from_yaml(path.encode() if path else None, C.byref(gn)) == 0, lib.dmk_last_error()
        assert gn.N_cutback_max == rcut == (5 if path else 3) and gn.N_staggered_iter_max == 10
        assert dict(alpha=gn.alpha, beta=gn.beta, eps_abs_curl=gn.eps_abs_curl, eps_rel_curl=gn.eps_rel_curl) == rpol
    # a key the parser does not know is refused by both sides (config.cpp:386-402 knows seven)
    p.write_text(yaml.safe_dump({'solver': {'grid': {'mechanical': {'eps_div': 1e-2}}}}))
    assert lib.dmk_numerics_from_yaml(str(p).encode(), C.byref(num), C.byref(der), C.byref(div)) != 0
    assert b"unknown key 'eps_div'" in lib.dmk_last_error()
    with pytest.raises(L.DmkError):
        job.read_numerics(str(p))
    for deck in ({'phase': {'mechanical': {'plastic': {'N_iter_Lp': 3}}}}, {'phase': {'mechanical': {'eps_state': 1e-3}}},
                 {'phase': {'mechanical': {'eigen': {'N_iter_Li': 3}}}}):
        p.write_text(yaml.safe_dump(deck))
        assert lib.dmk_numerics_from_yaml(str(p).encode(), C.byref(num), C.byref(der), C.byref(div)) != 0 and b'unknown key' in lib.dmk_last_error()
        with pytest.raises(L.DmkError):
            job.read_numerics(str(p))
    # values out of range are refused by both, with the same list (upstream's IO_error 301; Config::parse_numerics does not check)
    p.write_text(yaml.safe_dump({'solver': {'grid': {'mechanical': {'N_iter_min': 5, 'N_iter_max': 3, 'eps_rel_P': 0.0}}},
                                 'phase': {'mechanical': {'r_cutback': -0.5, 'plastic': {'N_iter_Lp_max': 0}}}}))
    assert lib.dmk_numerics_from_yaml(str(p).encode(), C.byref(num), C.byref(der), C.byref(div)) != 0
    cpp = lib.dmk_last_error().decode()
    with pytest.raises(L.DmkError) as err:
        job.read_numerics(str(p))
    assert cpp == 'numerics: value out of range: eps_rel_P, r_cutback, N_iter_Lp_max, N_iter_min' and cpp in str(err.value)
    # the struct's own names and the Li keys are accepted by both
    p.write_text(yaml.safe_dump({'solver': {'grid': {'mechanical': {'itmin': 2, 'itmax': 50, 'eps_abs_div': 1e-3, 'eps_rel_div': 1e-2}}},
                                 'phase': {'mechanical': {'eigen': {'N_iter_Li_max': 30}, 'eps_rel_state': 1e-5}}}))
    assert lib.dmk_numerics_from_yaml(str(p).encode(), C.byref(num), C.byref(der), C.byref(div)) == 0, lib.dmk_last_error()
    assert bytes(num) == bytes(job.read_numerics(str(p))[0]) and (num.itmin, num.itmax, num.eps_abs_div, num.eps_rel_state) == (2, 50, 1e-3, 1e-5)
    p.write_text(yaml.safe_dump({'solver': {'grid': {'N_staggered_iter_max': 7, 'FFT': {'alpha': 1.5, 'beta': 0.5, 'eps_abs_curl(F)': 1e-8, 'eps_rel_curl(F)': 1e-3}}}}))
    assert lib.dmk_grid_numerics_from_yaml(str(p).encode(), C.byref(gn)) == 0
    assert (gn.N_staggered_iter_max, gn.N_cutback_max, gn.alpha, gn.beta, gn.eps_abs_curl, gn.eps_rel_curl) == (7, 3, 1.5, 0.5, 1e-8, 1e-3)
    assert job.read_numerics(str(p))[4] == dict(alpha=1.5, beta=0.5, eps_abs_curl=1e-8, eps_rel_curl=1e-3)


def test_example_phase_configs_of_the_reference_through_both_readers(tmp_path):
    """every plastic parameter set the reference ships (examples/config/phase/mechanical/plastic/*.yaml, 18 files) combined with a
    Hooke set of its material: the Python host and the library's C++ reader must either both build the phase -- byte-identical
    tables -- or both refuse it, for the same reason (lattices hP / tI, the nonlocal law, transformation systems, shear banding
    without slip, cI slip families whose interaction types are not offered).  Reads the reference tree: build container only."""
    import glob
    from solver_cpp_b200 import lib as L, material as M
    base = '/root/reference/examples/config/phase/mechanical/'
    files = sorted(glob.glob(base + 'plastic/*.yaml'))
    if not files:
        pytest.skip('needs the reference tree')
    assert len(files) == 18
    elastic = {'AA6022': 'Al', 'phenopowerlaw_Al': 'Al', 'nonlocal_Al': 'Al', 'Au': 'Au', 'phenopowerlaw_Cu': 'Cu', 'Pt': 'Pt', 'DP-steel': 'Fe',
               'martensite': 'SAE1050-martensite', 'Mg': 'Mg', 'Ti': 'Ti', 'Sn': 'Sn-beta', 'dislotungsten_W': 'W', '317L': 'AISI304',
               'IF-steel': 'Fe', 'TWIP': 'TWIP-steel', 'Brass': 'Cu', 'Ni': 'Ni', 'free-surface': 'Al'}
    lattice = {'DP-steel': 'cI', 'martensite': 'cI', 'Mg': 'hP', 'Ti': 'hP', 'Sn': 'tI', 'dislotungsten_W': 'cI', 'IF-steel': 'cI'}
    lib = L.load()
    built, refused = [], {}
    for f in files:
        name = os.path.basename(f)
        key = [k for k in elastic if k in name]
        assert len(key) == 1, name
        lat = ([v for k, v in lattice.items() if k in name] or ['cF'])[0]
        phase = {'lattice': lat, 'mechanical': {'elastic': yaml.safe_load(open(base + 'elastic/Hooke_%s.yaml' % elastic[key[0]])),
                                                'plastic': yaml.safe_load(open(f))}}
        if lat in ('hP', 'tI'):
            phase['c/a'] = 1.6
        material = {'homogenization': {'SX': {'N_constituents': 1, 'mechanical': {'type': 'pass'}}}, 'phase': {'X': phase},
                    'material': [{'homogenization': 'SX', 'constituents': [{'phase': 'X', 'v': 1.0, 'O': [1.0, 0.0, 0.0, 0.0]}]}]}
        path = tmp_path / 'material.yaml'
        path.write_text(yaml.safe_dump(material))
        try:
            ref, why = M.Material(material, np.arange(1)), None
        except L.DmkError as e:
            ref, why = None, str(e)
        phases = (L.Phase * L.MAX_PHASES)()
        nph, nm = C.c_int(), C.c_int()
        ph_of, O = np.zeros(1, dtype=np.int32), np.zeros((1, 9))
        rc = lib.dmk_material_from_yaml(str(path).encode(), C.c_double(M.T_ROOM), phases, C.c_int(L.MAX_PHASES), C.byref(nph),
                                        ph_of.ctypes.data_as(C.c_void_p), O.ctypes.data_as(C.c_void_p), C.c_int(1), C.byref(nm))
        if ref is None:
            assert rc != 0, name
            cpp = lib.dmk_last_error().decode()
            assert why.split(' (')[0][-40:] in cpp, (name, why, cpp)            # the same reason
            refused[name] = why
        else:
            assert rc == 0, (name, lib.dmk_last_error())
            assert bytes(phases[0]) == bytes(ref.phases[0]), name
            built.append(name)
    assert built == ['dislotungsten_W.yaml', 'isotropic_free-surface.yaml', 'kinehardening_317L.yaml', 'phenopowerlaw_AA6022-T4.yaml',
                     'phenopowerlaw_Al.yaml', 'phenopowerlaw_Au.yaml', 'phenopowerlaw_Cu.yaml', 'phenopowerlaw_Pt-5%Cu.yaml']
    assert 'N_tr' in refused['dislotwin_TWIP-TRIP.yaml'] and 'nonlocal' in refused['nonlocal_Ni.yaml'] and 'hP' in refused['phenopowerlaw_Mg.yaml']


def test_example_elastic_configs_of_the_reference_through_both_readers(tmp_path):
    """every Hooke parameter set the reference ships (examples/config/phase/mechanical/elastic/*.yaml, 19 files, most with
    temperature polynomials C_ij,T / C_ij,T^2) evaluated at room temperature and at 500 K: the stiffness the C++ reader builds
    equals the Python host's bit for bit (phase_mechanical_elastic.f90:74-104, polynomials.f90:65-110).  Build container only."""
    import glob
    from solver_cpp_b200 import lib as L, material as M
    files = sorted(glob.glob('/root/reference/examples/config/phase/mechanical/elastic/*.yaml'))
    if not files:
        pytest.skip('needs the reference tree')
    assert len(files) == 19
    lib = L.load()
    differs = 0
    for f in files:
        for T in (M.T_ROOM, 500.0):
            material = {'homogenization': {'SX': {'N_constituents': 1, 'mechanical': {'type': 'pass'}}},
                        'phase': {'X': {'lattice': 'cF', 'mechanical': {'elastic': yaml.safe_load(open(f))}}},
                        'material': [{'homogenization': 'SX', 'constituents': [{'phase': 'X', 'v': 1.0, 'O': [1.0, 0.0, 0.0, 0.0]}]}]}
            path = tmp_path / 'material.yaml'
            path.write_text(yaml.safe_dump(material))
            ref = M.Material(material, np.arange(1), T=T)
            phases = (L.Phase * L.MAX_PHASES)()
            nph, nm = C.c_int(), C.c_int()
            ph_of, O = np.zeros(1, dtype=np.int32), np.zeros((1, 9))
            rc = lib.dmk_material_from_yaml(str(path).encode(), C.c_double(T), phases, C.c_int(L.MAX_PHASES), C.byref(nph),
                                            ph_of.ctypes.data_as(C.c_void_p), O.ctypes.data_as(C.c_void_p), C.c_int(1), C.byref(nm))
            assert rc == 0, (f, lib.dmk_last_error())
            assert bytes(phases[0]) == bytes(ref.phases[0]), (f, T)
            differs += int(T != M.T_ROOM and bytes(ref.phases[0]) != bytes(M.Material(material, np.arange(1)).phases[0]))
    assert differs >= 10                       # the polynomials are exercised: most sets change with temperature
