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
