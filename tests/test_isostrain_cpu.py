"""CPU tests of the isostrain homogenisation (N_constituents > 1: /root/reference/src/homogenization_mechanical_isostrain.f90:36-45,
mechanical_homogenize /root/reference/src/homogenization_mechanical.f90:125-145, material.f90:135-161): the oracle's restatement
against single-constituent oracles (which G1/G2/G4 pin), the host-side material tables (Python and the library's C++ reader), and
the refusals.  The reference holds no fixture with several constituents: the homogenisation itself is parity-unpinned; what it
averages is pinned."""
import ctypes as C
import os

import numpy as np
import pytest
import yaml

from oracle.constitutive import ConstitutiveOracle
from solver_cpp_b200 import workloads as W, material as M, lib as L


def _single(material, k):
    """the material with only constituent k of every entry (pass homogenisation)"""
    return {'homogenization': {'SX': {'N_constituents': 1, 'mechanical': {'type': 'pass'}}}, 'phase': material['phase'],
            'material': [{'homogenization': 'SX', 'constituents': [dict(e['constituents'][k], v=1.0)]} for e in material['material']]}


def _path(n, rng, steps=3):
    fl = rng.standard_normal((n, 3, 3))
    out = []
    for i in range(1, steps + 1):
        F = np.eye(3) + 4e-4 * i * fl
        F[:, 0, 0] += 1.5e-3 * i
        out.append(F)
    return out


@pytest.mark.parametrize('phases,nc', [(['Al_pheno'], 2), (['Al_pheno', 'Al_iso'], 3)])
def test_oracle_isostrain_is_the_weighted_sum_of_single_constituent_responses(phases, nc):
    cells = (4, 3, 2)
    n = int(np.prod(cells))
    material, material_at, _ = W.taylor_polycrystal(cells, 5, phases, nc, seed=2)
    o = ConstitutiveOracle(material, material_at)
    assert (o.n, o.n_cells, o.n_const) == (n * nc, n, nc)
    singles = [ConstitutiveOracle(_single(material, k), material_at) for k in range(nc)]
    v = np.array([[material['material'][m]['constituents'][k]['v'] for m in material_at] for k in range(nc)])
    assert np.abs(v.sum(0) - 1.0).max() < 1e-12
    rng = np.random.default_rng(0)
    dt = 0.5
    for F in _path(n, rng):
        ill = o.mechanical_response(F.reshape(n, 9), dt)
        P, dPdF = o.mechanical_response2(dt)
        accP = accT = None
        ill_s = False
        for k, s in enumerate(singles):
            ill_s |= s.mechanical_response(F.reshape(n, 9), dt)
            Pk, Tk = s.mechanical_response2(dt)
            # every constituent saw the cell's F and carries its own state: point k * n + cell
            assert np.array_equal(o.a['F'][k * n:(k + 1) * n], s.a['F'])
            assert np.array_equal(o.a['Fp'][k * n:(k + 1) * n], s.a['Fp'])
            assert np.array_equal(o.a['state'][k * n:(k + 1) * n, :s.max_state], s.a['state'])
            wP, wT = Pk * v[k][:, None, None], Tk * v[k][:, None, None, None, None]
            accP = wP if k == 0 else accP + wP
            accT = wT if k == 0 else accT + wT
        assert ill == ill_s
        assert np.array_equal(P, accP) and np.array_equal(dPdF, accT)        # same products, same order of the sum
        o.forward()
        for s in singles:
            s.forward()
    assert np.abs(o.a['Fp'] - np.eye(3).reshape(9)).max() > 1e-5             # the path went plastic


def test_identical_constituents_reproduce_the_single_crystal():
    cells = (3, 3, 2)
    n = int(np.prod(cells))
    material, material_at, _ = W.polycrystal(cells, 4, 'Al_pheno', 1)
    twin = {'homogenization': {'T': {'N_constituents': 2, 'mechanical': {'type': 'isostrain'}}}, 'phase': material['phase'],
            'material': [{'homogenization': 'T', 'constituents': [dict(e['constituents'][0], v=0.25), dict(e['constituents'][0], v=0.75)]}
                         for e in material['material']]}
    o1, o2 = ConstitutiveOracle(material, material_at), ConstitutiveOracle(twin, material_at)
    rng = np.random.default_rng(5)
    for F in _path(n, rng, 2):
        o1.mechanical_response(F.reshape(n, 9), 1.0); o2.mechanical_response(F.reshape(n, 9), 1.0)
        P1, T1 = o1.mechanical_response2(1.0); P2, T2 = o2.mechanical_response2(1.0)
        assert np.abs(P1 - P2).max() <= 4e-16 * np.abs(P1).max()
        assert np.abs(T1 - T2).max() <= 4e-16 * np.abs(T1).max()
        o1.forward(); o2.forward()


def test_material_tables_python_and_cpp_reader_agree(tmp_path):
    cells = (5, 4, 3)
    n = int(np.prod(cells))
    material, material_at, _ = W.taylor_polycrystal(cells, 6, ['Al_pheno', 'Cu_pheno'], 3, seed=4)
    m = M.Material(material, material_at)
    o = ConstitutiveOracle(material, material_at)
    assert m.n_constituents == 3 and m.n_cells == n and m.phase_of_point.shape == (3 * n,)
    assert np.array_equal(m.phase_of_point, o.phase_of) and np.array_equal(m.v_of_point, o.v_of)
    assert np.abs(m.O_of_point - o.a['Fp0']).max() < 1e-15
    # the library's own reader (what f_materialpoint_initDamask uses)
    path = str(tmp_path / 'material.yaml')
    with open(path, 'w') as f:
        yaml.safe_dump(material, f, sort_keys=False)
    lib = L.load()
    nm = len(material['material'])
    phases = (L.Phase * L.MAX_PHASES)()
    nph, nmat = C.c_int(), C.c_int()
    ncs = np.zeros(nm, dtype=np.int32); ph = np.zeros((nm, 4), dtype=np.int32)
    O = np.zeros((nm, 4, 9)); v = np.zeros((nm, 4))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.dmk_material_constituents_from_yaml(path.encode(), C.c_double(M.T_ROOM), phases, C.c_int(L.MAX_PHASES), C.byref(nph), C.c_int(4),
                                                 P(ncs), P(ph), P(O), P(v), C.c_int(nm), C.byref(nmat))
    assert rc == 0, lib.dmk_last_error()
    assert nph.value == 2 and nmat.value == nm and (ncs == 3).all()
    assert np.array_equal(ph[material_at][:, :3].T.reshape(-1), m.phase_of_point)
    assert np.array_equal(v[material_at][:, :3].T.reshape(-1), m.v_of_point)
    assert np.abs(O[material_at][:, :3].transpose(1, 0, 2).reshape(-1, 9) - m.O_of_point).max() < 1e-15
    # the single-constituent reader refuses the file
    ph1 = np.zeros(nm, dtype=np.int32); O1 = np.zeros((nm, 9))
    rc = lib.dmk_material_from_yaml(path.encode(), C.c_double(M.T_ROOM), phases, C.c_int(L.MAX_PHASES), C.byref(nph), P(ph1), P(O1),
                                    C.c_int(nm), C.byref(nmat))
    assert rc != 0 and b'constituents' in lib.dmk_last_error()


def test_refusals():
    cells = (2, 2, 2)
    material, material_at, _ = W.taylor_polycrystal(cells, 2, 'Al_pheno', 2, seed=0)
    import copy
    bad = copy.deepcopy(material)
    bad['material'][0]['constituents'][0]['v'] += 1e-6                      # material.f90:161
    with pytest.raises(L.DmkError):
        M.Material(bad, material_at)
    with pytest.raises(ValueError):
        ConstitutiveOracle(bad, material_at)
    bad = copy.deepcopy(material)
    bad['homogenization']['Taylor']['N_constituents'] = 3                  # material.f90:145
    with pytest.raises(L.DmkError):
        M.Material(bad, material_at)
    bad = copy.deepcopy(material)
    bad['homogenization']['Taylor']['mechanical']['type'] = 'RGC'
    with pytest.raises(L.DmkError):
        M.Material(bad, material_at)
    with pytest.raises(NotImplementedError):
        ConstitutiveOracle(bad, material_at)
    bad = copy.deepcopy(material)
    bad['material'][1]['constituents'] = bad['material'][1]['constituents'][:1]   # mixed counts in one grid
    bad['material'][1]['constituents'][0]['v'] = 1.0
    bad['material'][1]['homogenization'] = 'SX'
    bad['homogenization']['SX'] = {'N_constituents': 1, 'mechanical': {'type': 'pass'}}
    if len(np.unique(material_at)) == 2:
        with pytest.raises(L.DmkError):
            M.Material(bad, material_at)


def test_result_file_with_several_constituents_round_trips(tmp_path):
    """/cell_to/phase [N, C] with the entry counters of material.f90:177-203 (per phase: cell by cell, constituent by
    constituent), per-phase datasets in that order, and the reader's place() back to [N, C, ...]"""
    from solver_cpp_b200.dadf5 import ResultWriter, ResultReader
    cells = (4, 3, 2)
    n, nc = int(np.prod(cells)), 3
    material, material_at, size = W.taylor_polycrystal(cells, 5, ['Al_pheno', 'Cu_pheno'], nc, seed=6)
    w = ResultWriter(cells, size, material, material_at)
    rng = np.random.default_rng(1)
    F = rng.standard_normal((n, nc, 3, 3)); xi = rng.standard_normal((n, nc, 24)); O = rng.standard_normal((n, nc, 4))
    u_p = rng.standard_normal((n, 3)); u_n = rng.standard_normal(((cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1), 3))
    w.add_increment(4, 2.0, {'F': F, 'xi_sl': xi, 'O': O}, u_n=u_n, u_p=u_p)
    with pytest.raises(ValueError):
        w.add_increment(5, 2.5, {'F': F[:, :2]})
    path = str(tmp_path / 'taylor.hdf5')
    w.save(path)
    r = ResultReader(path)
    assert r.n_const == nc and r.n_cells == n and r.increments == [4] and r.times == [2.0]
    # the reference's counting rule, literally
    counter = {}
    lab = np.empty((n, nc), dtype=object); ent = np.zeros((n, nc), dtype=np.int64)
    for ce in range(n):
        for co in range(nc):
            ph = material['material'][material_at[ce]]['constituents'][co]['phase']
            lab[ce, co] = ph
            ent[ce, co] = counter.get(ph, 0)                       # 0-based in the file (result.f90 writes entry - 1)
            counter[ph] = counter.get(ph, 0) + 1
    assert np.array_equal(r.phase_label.reshape(n, nc), lab) and np.array_equal(r.phase_entry.reshape(n, nc), ent)
    assert np.array_equal(r.place('F', 4), F) and np.array_equal(r.place('O', 4), O)
    assert np.array_equal(r.place('xi_sl', 4), xi[..., :12])       # N_sl = 12 of both phases
    per_phase = r.get('F', 4)
    assert {k: v.shape[0] for k, v in per_phase.items()} == counter
    assert np.array_equal(r.geometry('u_p', 4), u_p)
