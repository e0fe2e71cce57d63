"""Extracts golden set G3 -- the literal inputs and expected values of the reference's own C++ unit tests of the hot path
(/root/reference/src/grid/test/test_mech_base.cpp, test_spectral.cpp) -- into tests/golden/g3_unit_test_literals.npz.
Run in the build container (the reference tree is not present on the GPU box):  python tests/golden/make_golden_g3.py"""
import os
import re
import numpy as np

REF = '/root/reference/src/grid/test/test_mech_base.cpp'
REF_SPECTRAL = '/root/reference/src/grid/test/test_spectral.cpp'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'g3_unit_test_literals.npz')


def test_body(src, name):
    i = src.index('TEST_F(GridTestSetup, %s)' % name)
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
    np.savez(OUT, **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == '__main__':
    main()
