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
