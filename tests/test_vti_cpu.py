"""CPU tests of the geometry reader / writer (solver_cpp_b200/vti.py): a file written by upstream DAMASK's own tooling
(tests/golden/damask_written_4grains2x4x3.vti, a 741-byte data fixture taken verbatim from the reference tree,
python/tests/resources/Result/4grains2x4x3.vti: VTK's base64 / zlib block conventions as DAMASK emits them) and the input
decks of golden set G4, which tests/golden/make_golden_g4.py re-writes with this module's writer."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('dmk_vti', os.path.join(HERE, '..', 'solver_cpp_b200', 'vti.py'))
vti = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(vti)          # by path: importing the package would load the CUDA library


def test_reads_a_damask_written_file(golden_dir):
    d = vti.read_vti(os.path.join(golden_dir, 'damask_written_4grains2x4x3.vti'))
    assert tuple(d['cells']) == (2, 4, 3) and np.allclose(d['size'], 1.0) and np.allclose(d['origin'], 0.0)
    assert d['material'].shape == (3, 4, 2) and np.bincount(d['material'].ravel()).tolist() == [6, 11, 2, 5]


def test_reads_the_g4_geometry(golden_dir):
    d = vti.read_vti(os.path.join(golden_dir, '6grains6x7x8_single_phase.vti'))
    assert tuple(d['cells']) == (6, 7, 8)
    assert np.allclose(d['size'], (0.75, 0.875, 1.0)) and np.allclose(d['origin'], 0.0)
    m = d['material']
    assert m.shape == (8, 7, 6) and m.min() == 0 and m.max() == 5          # RangeMin / RangeMax of the file
    assert np.bincount(m.ravel()).tolist() == [49, 58, 51, 56, 60, 62]


def test_round_trip_compressed_and_plain(tmp_path):
    rng = np.random.default_rng(0)
    m = rng.integers(0, 500, size=(9, 17, 33))
    for compress in (True, False):
        p = str(tmp_path / ('g%d.vti' % compress))
        vti.write_vti(p, m, (3.3e-5, 1.7e-5, 0.9e-5), origin=(1e-6, 0, 0), compress=compress, block=4096)
        d = vti.read_vti(p)
        assert (d['material'] == m).all() and tuple(d['cells']) == (33, 17, 9)
        assert np.allclose(d['size'], (3.3e-5, 1.7e-5, 0.9e-5), rtol=1e-14) and np.allclose(d['origin'], (1e-6, 0, 0))


def test_rejects_wrong_cell_count(tmp_path):
    import pytest
    p = str(tmp_path / 'bad.vti')
    vti.write_vti(p, np.zeros((2, 2, 2), int), (1, 1, 1))
    txt = open(p).read().replace('WholeExtent="0 2 0 2 0 2"', 'WholeExtent="0 3 0 2 0 2"')
    open(p, 'w').write(txt)
    with pytest.raises(ValueError):
        vti.read_vti(p)


def test_cpp_reader_symbols_match_the_python_reader(golden_dir, tmp_path):
    """f_VTI_readCellsSizeOrigin / f_VTI_readDataset_int of the C-ABI library (csrc/vti_dropin.cpp: the readers the reference's
    DiscretizationGrid binds, discretization_grid.h:20-22) on DAMASK-written and self-written files, compressed, plain and
    ascii: cells, size, origin and the material array PLUS ONE (VTI.f90:494) must equal what vti.py reads"""
    import ctypes as C
    lib = C.CDLL(os.path.join(HERE, '..', 'solver_cpp_b200', 'libdamask_grid_b200.so'))
    lib.dmk_last_error.restype = C.c_char_p
    rng = np.random.default_rng(1)
    files = [os.path.join(golden_dir, 'damask_written_4grains2x4x3.vti'), os.path.join(golden_dir, '6grains6x7x8_single_phase.vti')]
    m = rng.integers(0, 70000, size=(5, 6, 7))
    for compress in (True, False):
        p = str(tmp_path / ('w%d.vti' % compress))
        vti.write_vti(p, m, (7e-5, 6e-5, 5e-5), origin=(0, 2e-6, 0), compress=compress, block=512)
        files.append(p)
    p = str(tmp_path / 'ascii.vti')
    vti.write_vti(p, m, (7e-5, 6e-5, 5e-5), compress=False)
    txt = open(p).read()
    a, b = txt.index('format="binary"'), txt.index('</DataArray>')
    body_start = txt.index('>', a) + 1
    open(p, 'w').write(txt[:a] + 'format="ascii"' + txt[a + len('format="binary"'):body_start] + '\n' + ' '.join(str(v) for v in m.ravel()) + '\n' + txt[b:])
    files.append(p)
    for f in files:
        ref = vti.read_vti(f)
        content = open(f, 'rb').read()
        cells = (C.c_int * 3)(); size = (C.c_double * 3)(); origin = (C.c_double * 3)()
        lib.f_VTI_readCellsSizeOrigin(content, C.c_int(len(content)), cells, size, origin)
        assert tuple(cells) == tuple(ref['cells']), lib.dmk_last_error()
        assert np.allclose(list(size), ref['size'], rtol=1e-15) and np.allclose(list(origin), ref['origin'])
        n = int(np.prod(ref['cells']))
        data = np.zeros(n, dtype=np.int32)
        nn = C.c_int(n)
        lib.f_VTI_readDataset_int(content, C.c_int(len(content)), b'material', C.c_int(8), data.ctypes.data_as(C.c_void_p), C.byref(nn))
        assert nn.value == n, lib.dmk_last_error()
        assert np.array_equal(data, ref['material'].ravel() + 1)
    bad = open(files[0], 'rb').read().replace(b'ImageData', b'PolyData')
    cells = (C.c_int * 3)(); size = (C.c_double * 3)(); origin = (C.c_double * 3)()
    lib.f_VTI_readCellsSizeOrigin(bad, C.c_int(len(bad)), cells, size, origin)
    assert tuple(cells) == (-1, -1, -1) and b'VTI' in lib.dmk_last_error()


def test_g6_geometry_text_of_the_reference_unit_test(golden_dir):
    """the VTI text embedded in the reference's own reader test (test_discretization_grid.cpp:103-124, extracted by
    tests/golden/make_golden_g6.py) with the expectations of TestVTI_readCellsSizeOrigin (:127-139): cells 16^3, origin 0,
    size 1; TestVTI_readDataset_int (:141-149) reads 16^3 material IDs -- the file declares RangeMin 0 / RangeMax 19 -- into a
    32-bit array, PLUS ONE.  Through the library's f_VTI_* symbols (what DiscretizationGrid binds) and the Python reader."""
    import ctypes as C
    g6 = np.load(os.path.join(golden_dir, 'g6_discretization_literals.npz'))
    path = os.path.join(golden_dir, 'g6_unit_test_geometry.vti')
    content = open(path, 'rb').read()
    lib = C.CDLL(os.path.join(HERE, '..', 'solver_cpp_b200', 'libdamask_grid_b200.so'))
    lib.dmk_last_error.restype = C.c_char_p
    cells = (C.c_int * 3)(); size = (C.c_double * 3)(); origin = (C.c_double * 3)()
    lib.f_VTI_readCellsSizeOrigin(content, C.c_int(len(content)), cells, size, origin)
    assert list(cells) == g6['vti_expected_cells'].tolist(), lib.dmk_last_error()          # EXPECT_EQ: exact
    assert list(size) == g6['vti_expected_size'].tolist() and list(origin) == g6['vti_expected_origin'].tolist()
    n = 16 * 16 * 16
    data = np.zeros(n, dtype=np.int32)
    nn = C.c_int(n)
    lib.f_VTI_readDataset_int(content, C.c_int(len(content)), b'material', C.c_int(8), data.ctypes.data_as(C.c_void_p), C.byref(nn))
    assert nn.value == n, lib.dmk_last_error()
    assert data.min() == 1 and data.max() == 20 and np.unique(data).size == 20
    d = vti.read_vti(path)
    assert tuple(d['cells']) == (16, 16, 16) and list(d['size']) == [1.0, 1.0, 1.0] and list(d['origin']) == [0.0, 0.0, 0.0]
    assert np.array_equal(d['material'].ravel() + 1, data)
    # a Voronoi tessellation: every grain is a sizeable share of the 4096 cells
    assert np.bincount(d['material'].ravel()).min() > 40


def test_every_damask_written_geometry_of_the_reference_tree_reads_alike():
    """all .vti files under the reference tree (34: the fixtures of upstream DAMASK's own Grid / Result tests, written by its
    tooling with and without compression, several data types, non-cubic grids): the Python reader and the library's f_VTI_*
    symbols must agree on cells, size, origin and every material ID, and the IDs' range must be the RangeMin / RangeMax the file
    declares.  Reads the reference tree, so it runs in the build container only."""
    import ctypes as C
    import glob
    import re
    import pytest
    files = sorted(glob.glob('/root/reference/**/*.vti', recursive=True))
    if not files:
        pytest.skip('needs the reference tree')
    assert len(files) >= 30
    lib = C.CDLL(os.path.join(HERE, '..', 'solver_cpp_b200', 'libdamask_grid_b200.so'))
    lib.dmk_last_error.restype = C.c_char_p
    for f in files:
        content = open(f, 'rb').read()
        d = vti.read_vti(f)
        cells = (C.c_int * 3)(); size = (C.c_double * 3)(); origin = (C.c_double * 3)()
        lib.f_VTI_readCellsSizeOrigin(content, C.c_int(len(content)), cells, size, origin)
        assert tuple(cells) == tuple(d['cells']), (f, lib.dmk_last_error())
        assert np.allclose(list(size), d['size'], rtol=1e-15) and np.allclose(list(origin), d['origin'], rtol=1e-15, atol=0), f
        n = int(np.prod(d['cells']))
        data = np.zeros(n, dtype=np.int32)
        nn = C.c_int(n)
        lib.f_VTI_readDataset_int(content, C.c_int(len(content)), b'material', C.c_int(8), data.ctypes.data_as(C.c_void_p), C.byref(nn))
        assert nn.value == n and np.array_equal(data, d['material'].ravel() + 1), (f, lib.dmk_last_error())
        m = re.search(rb'Name="material"[^>]*RangeMin="([^"]+)"[^>]*RangeMax="([^"]+)"', content)
        if m:
            assert (float(m.group(1)), float(m.group(2))) == (d['material'].min(), d['material'].max()), f
