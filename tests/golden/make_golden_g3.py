"""Extracts golden set G3 -- the literal inputs and expected values of the reference's own C++ unit tests of the hot path
(/root/reference/src/grid/test/test_mech_base.cpp, test_spectral.cpp, test_mech_solver_basic.cpp) -- into tests/golden/g3_unit_test_literals.npz.
Run in the build container (the reference tree is not present on the GPU box):  python tests/golden/make_golden_g3.py"""
import os
import re
import numpy as np

REF = '/root/reference/src/grid/test/test_mech_base.cpp'
REF_SPECTRAL = '/root/reference/src/grid/test/test_spectral.cpp'
REF_SOLVER = '/root/reference/src/grid/test/test_mech_solver_basic.cpp'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'g3_unit_test_literals.npz')


def test_body(src, name):
    m = re.search(r'TEST_F\(GridTestSetup,\s*%s\)' % name, src)
    i = m.start()
    j = src.find('TEST_F(', i + 10)
    return src[i:j if j > 0 else len(src)]


def set_values(body, var):
    """the nested initialiser of `var.setValues({...});` as a numpy array (Eigen index order = nesting order)"""
    m = re.search(r'\b%s\.setValues\(\{' % re.escape(var), body)
    k = m.end() - 1
    depth, e = 0, k
    while True:
        depth += {'{': 1, '}': -1}.get(body[e], 0)
        e += 1
        if depth == 0:
            break
    txt = body[k:e]
    txt = re.sub(r'c\(\s*([^,()]+?)\s*,\s*([^,()]+?)\s*\)', r'complex(\1,\2)', txt)
    txt = txt.replace('{', '[').replace('}', ']').replace('true', 'True').replace('false', 'False')
    return np.array(eval(txt))


def main():
    src = open(REF).read()
    out = {}
    b = test_body(src, 'MechBaseTestDivergenceRMS')
    out['div_field'] = set_values(b, 'tensor_field')                # (3, 3, nx, ny, nz), natural (i, j)
    out['div_xi1st'] = set_values(b, 'spectral.xi1st')              # (3, nx/2+1, nz, ny)
    out['div_rms'] = float(re.search(r'ASSERT_EQ\(rms,\s*([0-9.eE+-]+)\)', b).group(1))
    out['div_cells'] = np.array([int(v) for v in re.search(r'MockDiscretizedGrid mock_grid\(std::array<int, 3>\{(\d+),(\d+),(\d+)\}\)', b).groups()])
    b = test_body(src, 'MechBaseTestMaskedCompliance')
    out['mc_C'] = set_values(b, 'C')                                # (3, 3, 3, 3) natural
    out['mc_expected'] = set_values(b, 'expected_masked_compliance')
    m = re.search(r'mask_stress <<([^;]+);', b)
    out['mc_mask_stress'] = np.array([v.strip() == 'true' for v in m.group(1).replace('\n', ' ').split(',')]).reshape(3, 3)
    b = test_body(open(REF_SPECTRAL).read(), 'SpectralTestConstitutiveResponse')
    out['hom_P'] = set_values(b, 'mech2_homogenization_P_result')          # (3, 3, cell)
    out['hom_dPdF'] = set_values(b, 'mech2_homogenization_dPdF_result')    # (3, 3, 3, 3, cell)
    out['hom_P_av'] = set_values(b, 'expected_P_av')
    out['hom_C_vol_avg'] = set_values(b, 'expected_C_vol_avg')
    out['hom_C_minmax_avg'] = set_values(b, 'expected_C_minmax_avg')
    # xi tables of Spectral::init on the 2 x 1 x 1 mock grid (test_spectral.cpp:15-48; geom_size = cells * 1e-5, conftest.h:39)
    b = test_body(open(REF_SPECTRAL).read(), 'SpectralTestInit')
    out['xi_cells'] = np.array([int(v) for v in re.search(r'mock_grid\(std::array<int, 3>\{(\d+),(\d+),(\d+)\}\)', b).groups()])
    out['xi_expected_1st'] = set_values(b, 'expected_xi1st')              # (3, nx/2+1, nz, ny)
    out['xi_expected_2nd'] = set_values(b, 'expected_xi2nd')
    # forward_field / calculate_rate literals (test_mech_base.cpp:199-237, 492-545)
    b = test_body(src, 'MechBaseTestForwardField')
    out['ff_field_last_inc'] = set_values(b, 'field_last_inc')            # (3, 3, nx, ny, nz)
    out['ff_rate'] = set_values(b, 'rate')
    out['ff_aim'] = set_values(b, 'aim')
    out['ff_delta_t'] = float(re.search(r'double delta_t = ([0-9.eE+-]+);', b).group(1))
    out['ff_expected'] = set_values(b, 'expected_forwarded_field')
    b = test_body(src, 'TestCalculateRate')
    out['cr_field'] = set_values(b, 'field')
    out['cr_field_0'] = set_values(b, 'field_0')
    out['cr_av_rate'] = set_values(b, 'avRate')
    out['cr_expected'] = set_values(b, 'expected_rate')
    m = re.search(r'calculate_rate\((true|false), field_0, field, ([0-9.eE+-]+), avRate\)', b)
    out['cr_heterogeneous'] = m.group(1) == 'true'
    out['cr_delta_t'] = float(m.group(2))
    # residual glue with mocked constitutive response, divergence and Gamma (test_mech_solver_basic.cpp:232-376)
    b = test_body(open(REF_SOLVER).read(), 'TestFormResidual')
    out['fr_F'] = set_values(b, 'F')                                      # (3, 3, 2, 1, 1)
    out['fr_S'] = set_values(b, 'mech_basic.S')
    out['fr_P_av_before'] = set_values(b, 'mech_basic.P_av')
    out['fr_P'] = set_values(b, 'mocked_constitutive_response_P')
    out['fr_expected_gamma_aim'] = set_values(b, 'expected_rotated')
    out['fr_expected_F_aim'] = set_values(b, 'expected_F_aim')
    m = re.search(r'mech_basic\.F_aim <<([^;]+);', b)
    out['fr_F_aim'] = np.array([float(v) for v in m.group(1).replace('\n', ' ').split(',')]).reshape(3, 3)
    m = re.search(r'stress_mask <<([^;]+);', b)
    out['fr_stress_mask'] = np.array([v.strip() == 'true' for v in m.group(1).replace('\n', ' ').split(',')]).reshape(3, 3)
    out['fr_err_BC'] = float(re.search(r'EXPECT_EQ\(mech_basic\.err_BC,\s*([0-9.eE+-]+)\)', b).group(1))
    out['fr_err_div'] = float(re.search(r'EXPECT_EQ\(mech_basic\.err_div,\s*([0-9.eE+-]+)\)', b).group(1))
    # Gamma convolution on the 2 x 1 x 1 grid (xi1st = 0: DC path and unnormalised-inverse convention; test_mech_base.cpp:389-490)
    b = test_body(src, 'MechBaseTestGammaConvolution')
    out['gc_field_aim'] = set_values(b, 'field_aim')
    out['gc_C_ref'] = set_values(b, 'mech_base.C_ref')
    out['gc_field'] = set_values(b, 'field')                              # (3, 3, 2, 1, 1)
    out['gc_expected'] = set_values(b, 'expected_gamma_field')
    # update_gamma with xi = 0 -> Gamma_hat = 0 (test_mech_base.cpp:154-197): the stiffness literal
    b = test_body(src, 'TestUpdateGamma')
    out['ug_C'] = set_values(b, 'C_min_max_avg')
    # divergence correction of the geometry (test_mech_base.cpp:54-76)
    b = test_body(src, 'MechBaseTestInitDivergenceCorrection')
    out['dc_geom_size'] = np.array([float(v) for v in re.search(r'geom_size = std::array<double, 3>\{([^}]*)\}', b).group(1).split(',')])
    out['dc_cells'] = np.array([int(v) for v in re.search(r'mock_grid\(std::array<int, 3>\{(\d+),(\d+),(\d+)\}\)', b).groups()])
    out['dc_expected_size'] = np.array([float(v) for v in re.search(r'expected_scaled_geom_size_1\(\{([^}]*)\}\)', b).group(1).split(',')])
    out['dc_expected_size_grid'] = np.array([float(v) for v in re.search(r'expected_scaled_geom_size_2\(\{([^}]*)\}\)', b).group(1).split(',')])
    # update_coords at F = I on the 2 x 1 x 1 grid (test_mech_base.cpp:78-152)
    b = test_body(src, 'MechBaseTestUpdateCoords')
    out['uc_F'] = set_values(b, 'F_')
    out['uc_expected_x_n'] = set_values(b, 'expected_x_n')                # (3, nx+1, ny+1, nz+1)
    out['uc_expected_x_p'] = set_values(b, 'expected_x_p')                # (3, nx, ny, nz)
    np.savez(OUT, **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == '__main__':
    main()
