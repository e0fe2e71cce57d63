"""The minimal HDF5 reader against a result file written by real DAMASK through libhdf5
(tests/golden/damask_written_marc_job.hdf5: the reference's python/tests/resources/Result/check_compile_job1.hdf5, a data
fixture, unchanged).  The checks are properties of the mesh the file describes, so they do not depend on the reader
being right: connectivity within the node range, cell centres = mean of the cell's nodes, volumes adding up."""
import os

import numpy as np

from solver_cpp_b200.hdf5_min import H5File, check_group_structure

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'damask_written_marc_job.hdf5')


def test_tree_and_attributes():
    f = H5File(PATH)
    assert f.keys('/') == ['cell_to', 'geometry', 'increment_1', 'increment_2', 'setup']
    assert check_group_structure(f) == 36
    a = f.attrs('/')
    assert a['DADF5_version_major'] == 1 and a['DADF5_version_minor'] == 0
    assert a['created'] == '2023-02-04 13:32:38+0100' and 'DAMASK_Marc' in a['creator']       # variable-length strings
    assert f.attrs('/increment_2')['t/s'] == 100.0
    at = f.attrs('/geometry/v_0')
    assert at['unit'] == 'm³' and at['description'] == 'initial cell volume'                   # UTF-8
    assert f.attrs('/geometry/T_c')['VTK_TYPE'] == 'HEXAHEDRON'


def test_compound_mapping():
    ph = H5File(PATH).read('/cell_to/phase')
    assert ph.dtype.names == ('label', 'entry') and ph.shape == (8, 1)
    assert (ph['label'] == b'Aluminum_J2isotropic').all() and (ph['entry'].reshape(-1) == np.arange(8)).all()


def test_mesh_datasets_are_mutually_consistent():
    f = H5File(PATH)
    T, x_n, x_p, v0 = f.read('/geometry/T_c'), f.read('/geometry/x_n'), f.read('/geometry/x_p'), f.read('/geometry/v_0')
    assert T.shape == (8, 8) and T.dtype == np.int64 and T.min() == 1 and T.max() == x_n.shape[0] == 27
    assert np.abs(x_n[T - 1].mean(1) - x_p).max() < 1e-12 * np.abs(x_n).max()
    ext = x_n.max(0) - x_n.min(0)
    assert abs(v0.sum() - ext.prod()) < 1e-12 * ext.prod()
    n0, a0 = f.read('/geometry/n_0'), f.read('/geometry/a_0')
    assert np.abs(np.linalg.norm(n0, axis=-1) - 1).max() < 1e-12
    assert np.abs((n0 * a0[..., None]).sum(1)).max() < 1e-12 * a0.max()                        # closed cells
    u_n, u_p = f.read('/increment_2/geometry/u_n'), f.read('/increment_2/geometry/u_p')
    assert np.abs(u_n[T - 1].mean(1) - u_p).max() < 1e-9 * max(np.abs(u_n).max(), 1e-300)


def test_chunk_checksums_written_by_libhdf5_verify_and_catch_corruption(tmp_path):
    """every dataset of the fixture is chunked with libhdf5's fletcher32 filter: the reader recomputes the checksum of each
    chunk it decodes (an independent check of the chunk walk), and a flipped data byte is reported"""
    import struct
    from solver_cpp_b200.hdf5_min import fletcher32
    f = H5File(PATH)
    assert f.verify_checksums
    n = 0
    for grp in ('/geometry', '/increment_1/geometry', '/increment_2/geometry', '/setup'):
        for k in f.keys(grp):
            q = grp + '/' + k
            msgs = f._messages(f._resolve(q))
            assert [f._filters(d) for t, _, d in msgs if t == 0x000B] == [[(3, ())]]
            f.read(q)
            n += 1
    assert n == 13
    # corrupt one byte of the x_n chunk (its address from the chunk B-tree)
    lay = [d for t, _, d in f._messages(f._resolve('/geometry/x_n')) if t == 0x0008][0]
    btree = struct.unpack_from('<Q', lay, 3)[0]
    chunk = struct.unpack_from('<Q', f.b, btree + 24 + 8 + 8 * 3)[0]
    blob = bytearray(f.b)
    blob[chunk + 5] ^= 0x10
    bad = H5File.__new__(H5File)
    bad.b = bytes(blob)
    bad.root = bad._symbol_entry(56)
    try:
        bad.read('/geometry/x_n')
        raise AssertionError('corruption not detected')
    except ValueError as e:
        assert 'fletcher32' in str(e)
    bad.verify_checksums = False
    assert bad.read('/geometry/x_n').shape == (27, 3)
    assert fletcher32(b'') == 0 and fletcher32(b'\x01') == 0x01000100 and fletcher32(b'\x00\x01\x00\x02') == 0x00040003


def test_checksummed_datasets_are_stored_like_libhdf5_stores_them(tmp_path):
    """H5Writer.dataset(checksum=True) against the DAMASK-written fixture: writing the fixture's own x_n back gives the same
    fill-value, filter and layout messages and the same chunk B-tree node (header, first key with the chunk size that
    includes the checksum, end key) as libhdf5 produced -- byte for byte, addresses aside."""
    import struct
    from solver_cpp_b200.hdf5_min import H5Writer
    real = H5File(PATH)
    x_n = real.read('/geometry/x_n')
    w = H5Writer()
    w.dataset('/geometry/x_n', x_n, checksum=True)
    w.dataset('/plain', x_n)
    path = str(tmp_path / 'c.hdf5')
    w.save(path)
    mine = H5File(path)
    assert np.array_equal(mine.read('/geometry/x_n'), x_n) and np.array_equal(mine.read('/plain'), x_n)
    m_real = {t: d for t, _, d in real._messages(real._resolve('/geometry/x_n')) if t in (0x5, 0xB, 0x8)}
    m_mine = {t: d for t, _, d in mine._messages(mine._resolve('/geometry/x_n')) if t in (0x5, 0xB, 0x8)}
    assert m_real[0x5] == m_mine[0x5] and m_real[0xB] == m_mine[0xB]
    assert m_real[0x8][:3] == m_mine[0x8][:3] and m_real[0x8][11:] == m_mine[0x8][11:]
    key = 8 + 8 * 3
    node = lambda f, lay: f.b[struct.unpack_from('<Q', lay, 3)[0]:][:24 + key + 8 + key]
    a, b = node(real, m_real[0x8]), node(mine, m_mine[0x8])
    assert a[:24 + key] == b[:24 + key] and a[24 + key + 8:] == b[24 + key + 8:]
    assert struct.unpack_from('<I', b, 24)[0] == x_n.nbytes + 4
    # and the stored checksum is the one libhdf5 computed for the same bytes
    chunk = lambda f, n: f.b[struct.unpack_from('<Q', n, 24 + key)[0]:][:x_n.nbytes + 4]
    assert chunk(real, a) == chunk(mine, b)


def test_attribute_messages_equal_the_ones_libhdf5_wrote():
    """integer and float scalar attributes: the whole attribute message (name, datatype, dataspace, value) byte for byte"""
    from solver_cpp_b200.hdf5_min import H5Writer
    real = H5File(PATH)
    w = H5Writer()
    w.group('/', DADF5_version_major=np.int64(1))
    w.group('/increment_2', **{'t/s': 100.0})
    mine = H5File.__new__(H5File)
    mine.b = w.tobytes()
    mine.root = mine._symbol_entry(56)
    attr = lambda f, p, name: [d for t, _, d in f._messages(f._resolve(p)) if t == 0x000C and name in d][0]
    assert attr(real, '/', b'DADF5_version_major') == attr(mine, '/', b'DADF5_version_major')
    assert attr(real, '/increment_2', b't/s') == attr(mine, '/increment_2', b't/s')
