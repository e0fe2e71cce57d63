"""Extracts golden set G6 -- the fixtures of the reference's discretisation unit tests
(/root/reference/src/grid/test/test_discretization_grid.cpp: the 16 x 16 x 16 VTI geometry text of VTISetup, :103-124, with the
expectations of TestVTI_readCellsSizeOrigin, :127-139, and the cell-centre / node coordinate literals of the 2 x 1 x 1 grid,
:70-101) -- into tests/golden/g6_unit_test_geometry.vti and g6_discretization_literals.npz.
Run in the build container (the reference tree is not present on the GPU box):  python tests/golden/make_golden_g6.py"""
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_g3 import set_values          # noqa: E402

REF = '/root/reference/src/grid/test/test_discretization_grid.cpp'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(REF).read()
    m = re.search(r'vti_file_content = \(R""""\((.*?)\)""""\)', src, re.S)
    open(os.path.join(HERE, 'g6_unit_test_geometry.vti'), 'w').write(m.group(1))
    out = {}
    i = src.index('TEST_F(VTISetup, TestVTI_readCellsSizeOrigin)')
    b = src[i:src.index('TEST_F(', i + 10)]
    for key in ('cells', 'origin', 'size'):
        vals = re.search(r'expected_%s = \{([^}]*)\}' % key, b).group(1)
        out['vti_expected_' + key] = np.array([float(v) for v in vals.split(',')])
    i = src.index('TEST(CoordCalculationSetup, TestCalculateIpCoords0)')
    b = src[i:src.index('TEST(', i + 10)]
    out['coords_cells'] = np.array([int(v) for v in re.search(r'cells\{(\d+), (\d+), (\d+)\}', b).groups()])
    out['coords_size'] = np.array([float(v) for v in re.search(r'geom_size = \{([^}]*)\}', b).group(1).split(',')])
    out['ip_coords0'] = set_values(b, 'expected_ip_coords0')              # (3, cells)
    i = src.index('TEST(CoordCalculationSetup, TestCalculateNodes0)')
    b = src[i:src.index('class VTISetup', i)]
    out['node_coords0'] = set_values(b, 'expected_node_coords0')          # (3, nodes), x fastest
    np.savez(os.path.join(HERE, 'g6_discretization_literals.npz'), **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == '__main__':
    main()
