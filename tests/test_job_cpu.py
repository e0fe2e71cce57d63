"""CPU tests of the host-side input plumbing that needs no GPU: numerics.yaml -> dmk_numerics, load.yaml -> load steps."""
import os
import sys

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_numerics_defaults_and_overrides(tmp_path):
    """defaults of config.h:50-78 / phase_mechanical.f90:277-294; the reference's examples/grid/numerics.yaml overrides
    N_iter_min / N_iter_max; integrator and FFT keys"""
    from solver_cpp_b200 import job
    num, derivative, divcorr, max_cut_back, pol = job.read_numerics(None)
    assert pol == dict(alpha=1.0, beta=1.0, eps_abs_curl=1.0e-10, eps_rel_curl=5.0e-4)
    assert max_cut_back == 3
    assert (num.itmin, num.itmax, num.N_iter_state_max, num.N_iter_Lp_max, num.integrator_state) == (1, 100, 20, 40, 0)
    assert (num.eps_rel_div, num.eps_abs_div, num.eps_rel_P, num.eps_abs_P) == (5.0e-4, 1.0e-4, 1.0e-3, 1.0e3)
    assert (derivative, divcorr) == (0, 2)
    p = tmp_path / 'numerics.yaml'
    p.write_text(yaml.safe_dump({'solver': {'grid': {'mechanical': {'N_iter_min': 4, 'N_iter_max': 40, 'eps_rel_div': 1e-4},
                                                     'FFT': {'derivative': 'FWBW_difference', 'divergence_correction': 'size'}}},
                                 'phase': {'mechanical': {'r_cutback': 0.5, 'plastic': {'integrator_state': 'RK4', 'N_iter_Lp_max': 25}}}}))
    num, derivative, divcorr, _, _ = job.read_numerics(str(p))
    assert (num.itmin, num.itmax, num.eps_rel_div, num.r_cutback, num.N_iter_Lp_max, num.integrator_state) == (4, 40, 1e-4, 0.5, 25, 3)
    assert (derivative, divcorr) == (2, 1)


def test_numerics_keys_of_the_reference_parser(tmp_path):
    """Config::parse_numerics (config.cpp:383-402) spells the divergence tolerances `eps_abs_div(P)` / `eps_rel_div(P)` and reads
    solver/grid/N_cutback_max; a key it does not know must not be dropped silently"""
    from solver_cpp_b200 import job, lib as L
    p = tmp_path / 'numerics.yaml'
    p.write_text(yaml.safe_dump({'solver': {'grid': {'N_cutback_max': 5, 'mechanical': {'eps_abs_div(P)': 1e-2, 'eps_rel_div(P)': 1e-1,
                                                                                         'eps_abs_P': 2e3, 'eps_rel_P': 2e-3, 'update_gamma': True}}}}))
    num, _, _, max_cut_back, _ = job.read_numerics(str(p))
    assert (num.eps_abs_div, num.eps_rel_div, num.eps_abs_P, num.eps_rel_P, num.update_gamma, max_cut_back) == (1e-2, 1e-1, 2e3, 2e-3, 1, 5)
    p.write_text(yaml.safe_dump({'solver': {'grid': {'mechanical': {'eps_div': 1e-2}}}}))
    with pytest.raises(L.DmkError):
        job.read_numerics(str(p))
    p.write_text(yaml.safe_dump({'phase': {'mechanical': {'plastic': {'N_iter_Lp': 3}}}}))
    with pytest.raises(L.DmkError):
        job.read_numerics(str(p))
    p.write_text(yaml.safe_dump({'solver': {'grid': {'FFT': {'derivative': 'spectral'}}}}))
    with pytest.raises(ValueError):
        job.read_numerics(str(p))


def test_load_file_of_the_golden_deck(golden_dir):
    from solver_cpp_b200 import mech
    steps = mech.parse_load(yaml.safe_load(open(os.path.join(golden_dir, '6grains6x7x8_tensionY.load.yaml'))))
    assert len(steps) == 1 and steps[0]['N'] == 40 and steps[0]['t'] == 20
    d, s = steps[0]['deformation'], steps[0]['stress']
    assert d.type == 'dot_F' and d.values[1][1] == 1.0e-3 and s.type == 'P'
    # dot_F prescribes yy and the shears, P the xx and zz components: the masks are complementary
    assert (d.mask.astype(bool) != s.mask.astype(bool)).all()


def test_unknown_integrator_is_rejected():
    from solver_cpp_b200 import material as M, lib as L
    with pytest.raises(L.DmkError):
        M.make_numerics({'plastic': {'integrator_state': 'Heun'}})


# the input decks of the reference's own parser tests (test_config.cpp:34-58 and :96-118), key for key
_REF_NUMERICS = """
solver:
  grid:
    N_staggered_iter_max: 10
    N_cutback_max:        3

    mechanical:
      N_iter_min: 1
      N_iter_max: 100
      eps_abs_div(P): 1.0e-4
      eps_rel_div(P): 5.0e-4
      eps_abs_P: 1.0e3
      eps_rel_P: 1.0e-3
      update_gamma: false

    FFT:
      memory_efficient: true
      divergence_correction: size+grid
      derivative: continuous
      FFTW_plan_mode: FFTW_MEASURE
      FFTW_timelimit: -1.0
      PETSc_options: -snes_type ngmres -snes_ngmres_anderson
      alpha: 1.0
      beta: 1.0
      eps_abs_curl(F): 1.0e-10
      eps_rel_curl(F): 5.0e-4
"""

_REF_LOAD = """
---

solver:
  mechanical: spectral_basic

loadstep:
  - boundary_conditions:
      mechanical:
        dot_F: [[1.0e-3, 1, 0],
                [0,      x, 0],
                [2,      0, x]]
        P: [[x, x, x],
            [x, 0, x],
            [x, x, 0]]
        R: [1, 2, 3, 4]
    discretization:
      t: 60
      N: 120
      r: 2
    estimate_rate: false
    f_out: 4
    f_restart: 10
"""


def test_numerics_deck_of_the_reference_parser_test(tmp_path):
    """test_config.cpp:33-72 (TestYamlReadSuccess): the same deck gives the same values, through the Python host and through
    the library's own reader (dmk_numerics_from_yaml, what f_materialpoint_initDamask uses)"""
    import ctypes as C
    from solver_cpp_b200 import job, lib as L
    p = tmp_path / 'numerics.yaml'
    p.write_text(_REF_NUMERICS)
    num, derivative, divcorr, max_cut_back, pol = job.read_numerics(str(p))
    assert max_cut_back == 3
    assert (num.itmin, num.itmax, num.update_gamma) == (1, 100, 0)
    assert (num.eps_abs_div, num.eps_rel_div, num.eps_abs_P, num.eps_rel_P) == (1.0e-4, 5.0e-4, 1.0e3, 1.0e-3)
    assert (derivative, divcorr) == (0, 2)
    assert pol == dict(alpha=1.0, beta=1.0, eps_abs_curl=1.0e-10, eps_rel_curl=5.0e-4)
    lib = L.load()
    n2 = L.Numerics()
    der, div = C.c_int(-1), C.c_int(-1)
    assert lib.dmk_numerics_from_yaml(str(p).encode(), C.byref(n2), C.byref(der), C.byref(div)) == 0, lib.dmk_last_error()
    assert (n2.itmin, n2.itmax, n2.update_gamma, der.value, div.value) == (1, 100, 0, 0, 2)
    assert (n2.eps_abs_div, n2.eps_rel_div, n2.eps_abs_P, n2.eps_rel_P) == (1.0e-4, 5.0e-4, 1.0e3, 1.0e-3)


def test_load_deck_of_the_reference_parser_test():
    """test_config.cpp:95-145 (TestYamlReadBasic): values, masks (true where the deck says x), discretisation, flags"""
    import numpy as np
    from solver_cpp_b200 import mech, material as M
    ld = yaml.safe_load(_REF_LOAD)
    assert ld['solver'] == {'mechanical': 'spectral_basic'}
    st = mech.parse_load(ld)[0]
    assert (st['r'], st['t'], st['N'], st['estimate_rate'], st['f_out'], st['f_restart']) == (2, 60, 120, False, 4, 10)
    d, s = st['deformation'], st['stress']
    assert d.type == 'dot_F' and s.type == 'P'
    assert np.array_equal(d.values, [[0.001, 1, 0], [0, 0, 0], [2, 0, 0]])
    assert np.array_equal(d.mask.astype(bool), [[False, False, False], [False, True, False], [False, False, True]])
    assert np.array_equal(s.mask.astype(bool), ~d.mask.astype(bool))
    assert np.array_equal(s.values, np.zeros((3, 3)))
    assert np.array_equal(st['R'], M.qu2om([1, 2, 3, 4]))          # the quaternion (w, x, y, z) of the deck, as a matrix
    # the oracle's parser reads the deck alike
    from oracle.solver import parse_load as oparse
    so = oparse(ld)[0]
    assert (so['r'], so['t'], so['N'], so['estimate_rate'], so['f_out']) == (2, 60, 120, False, 4)
    assert np.array_equal(so['deformation'].values, d.values) and np.array_equal(so['deformation'].mask.astype(bool), d.mask.astype(bool))
