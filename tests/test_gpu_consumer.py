"""A compiled C++ program (tests/abi/consumer.cpp, built by __graft_entry__.build()) drives the drop-in boundary the way the
reference's driver does -- the init trio (main.cpp:80-86), the Fortran-named f_homogenization_* symbols on callee-owned host
arrays, and the residual through a function of PETSc's DMDA signature -- and must print what the Python host computes through
ctypes on the same input."""
import os
import subprocess

import numpy as np
import pytest
import yaml

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _parse(out):
    d, evals = {}, []
    for line in out.splitlines():
        parts = line.split()
        if not parts:
            continue
        if parts[0] == 'level2_eval':
            evals.append({})
        elif parts[0].startswith('level2_'):
            evals[-1][parts[0]] = np.array([float(v) for v in parts[1:]])
        elif len(parts) > 1:
            d[parts[0]] = np.array([float(v) for v in parts[1:]])
    return d, evals


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize('phase', ['Al_pheno', 'TWIP', 'taylor'])     # taylor: isostrain aggregates of 2 Al crystallites through the init trio
def test_cpp_consumer_matches_python_host(tmp_path, phase):
    import __graft_entry__ as ge
    exe = ge.build_consumer()
    from solver_cpp_b200 import workloads as W, material as M, mech
    from solver_cpp_b200.grid import GridContext, fortran_to_c3333
    nx, ny, nz, n_mat = 8, 6, 4, 7
    if phase == 'taylor':
        material, _, _ = W.taylor_polycrystal((nx, ny, nz), n_mat, ['Al_pheno'], 2, 3)
    else:
        material, _, _ = W.polycrystal((nx, ny, nz), n_mat, phase, 3)
    path = tmp_path / 'material.yaml'
    path.write_text(yaml.safe_dump(material))
    r = subprocess.run([exe, str(path), str(nx), str(ny), str(nz), str(n_mat), str(tmp_path / 'no_numerics.yaml')], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'CONSUMER_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    got, evals = _parse(r.stdout)
    # ---- the same through the Python host
    n = nx * ny * nz
    idx = np.arange(n)
    x, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    material_at = (x // 2 + 3 * (j // 2) + 5 * (k // 2)) % n_mat
    size = (1.0e-5 * nx / ny, 1.0e-5, 1.0e-5 * nz / ny)
    F = np.zeros((n, 3, 3))
    s, c = 1.0e-3 * np.sin(0.7 * idx), 5.0e-4 * np.cos(0.3 * idx)
    F[:, 0, 0] = 1.0 + 1.0e-3 + s; F[:, 0, 1] = c; F[:, 1, 0] = -c; F[:, 1, 1] = 1.0 - 4.0e-4; F[:, 1, 2] = 0.5 * s; F[:, 2, 2] = 1.0 - 4.0e-4 + c
    g = GridContext((nx, ny, nz), size)
    g.material_init(M.Material(material, material_at))
    P = g.to_natural(g.constitutive_response(g.to_device(F.reshape(nz, ny, nx, 3, 3)), 0.5)).reshape(n, 3, 3)
    _, _, _, d = g.tangent_response(0.5, materialise=True)
    assert int(got['level1_ill'][0]) == int(g.terminally_ill())
    assert _rel(got['level1_P_av'].reshape(3, 3).T, P.mean(0)) < 1e-13
    assert abs(got['level1_dPdF_sum'][0] - float(d.sum().item())) <= 1e-12 * abs(float(d.abs().sum().item()))
    # level 2
    sv = mech.make_solver((nx, ny, nz), size, material, material_at)
    sv.g.constitutive_response(sv.g.to_device(F.reshape(nz, ny, nx, 3, 3)), 0.5)       # the consumer's level-1 call left this state behind
    sv.init()
    st = mech.parse_load(W.TENSION_X)[0]
    sv.forward(False, False, 0.5, 1.0, 5.0, st['deformation'], st['stress'], st['R'])
    sv.g.materialpoint_forward()
    sv.prepare_evaluation()
    Fh = sv.field_host('F')
    rh = np.zeros_like(Fh)
    assert len(evals) == 2
    for e in evals:
        sv.form_residual_host(Fh, rh)
        assert _rel(e['level2_P_av'].reshape(3, 3).T, sv.get33('P_av')) < 1e-12
        assert abs(e['level2_err_div'][0] - sv.get('err_div')[0]) <= 1e-12 * sv.get('err_div')[0]
        assert abs(e['level2_err_BC'][0] - sv.get('err_BC')[0]) <= 1e-9 * max(sv.get('err_BC')[0], 1.0)
        assert abs(e['level2_r_sq'][0] - float((rh ** 2).sum())) <= 1e-11 * float((rh ** 2).sum())
        Fh = Fh - rh
    assert int(got['math_invert_err'][0]) == 0 and abs(got['math_invert_00'][0] - np.linalg.inv(np.array([[4., 1, 0], [1, 3, 1], [0, 1, 2]]))[0, 0]) < 1e-15


@pytest.mark.skipif(os.environ.get('DMK_TEST_UNVERIFIED') != '1',
                    reason='tests/abi/native_driver.cpp was written after the GPU budget of round 2 was spent: compiled, linked and run up to '
                           'the device on CPU (tests/test_abi_cpu.py), not yet run on a GPU; set DMK_TEST_UNVERIFIED=1 to run it')
def test_native_cpp_driver_solves_the_reference_deck_like_the_python_host(golden_dir, tmp_path):
    """the whole job in C++ (geometry -> init trio -> dmk_mech_init -> dmk_loop_over_loadcase) on the reference's own 16^3 deck
    against the Python host's loop on the same input: iteration counts equal, P_av and F_aim to 1e-12"""
    import __graft_entry__ as ge
    exe = ge.build_native_driver()
    from solver_cpp_b200 import job
    deck = [os.path.join(golden_dir, f) for f in ('g6_unit_test_geometry.vti', 'material_20grains_Al.yaml', 'load_tensionX.yaml')]
    load = yaml.safe_load(open(deck[2]))
    load['loadstep'][0]['f_out'] = 1
    lp = tmp_path / 'load.yaml'
    lp.write_text(yaml.safe_dump(load, default_flow_style=None))
    r = subprocess.run([exe, deck[0], deck[1], str(lp), '', '3'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'NATIVE_DRIVER_OK 3 increments' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [[float(v) for v in l.split()[1:]] for l in r.stdout.splitlines() if l.startswith('increment ')]
    solver, steps = job.load_job(deck[0], deck[1], str(lp))
    solver.init()
    from solver_cpp_b200 import mech
    recs = mech.loop_over_loadcase(solver, steps, max_increments=3)
    assert len(rows) == len(recs) == 3
    for row, rec in zip(rows, recs):
        assert (int(row[0]), int(row[2])) == (rec['inc'], rec['iterations']) and row[1] == rec['t']
        assert _rel(np.array(row[3:12]).reshape(3, 3), rec['P_av']) < 1e-12
        assert _rel(np.array(row[12:21]).reshape(3, 3), rec['F_aim']) < 1e-12
