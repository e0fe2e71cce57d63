"""Result-file output: the HDF5 writer against the reader (which tests/test_hdf5... pins to files written by
real DAMASK), the strict group-structure walk libhdf5 relies on, and the DADF5 frame written from the reference's own
input deck against golden set G5 (the frame of the real result file of that deck).  Host-side only."""
import os

import numpy as np
import pytest
import yaml

from solver_cpp_b200 import dadf5
from solver_cpp_b200.hdf5_min import H5File, H5Writer, check_group_structure
from solver_cpp_b200.vti import read_vti

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_writer_round_trip_all_types_and_a_three_level_group_tree(tmp_path):
    rng = np.random.default_rng(0)
    w = H5Writer()
    w.group('/', version=np.int32(3), creator='someone', ratio=0.25)
    w.group('/geometry', cells=np.array([6, 7, 8], np.int32), size=np.array([.75, .875, 1.]))
    data = {}
    for i in range(700):                                     # 700 links: 88 symbol nodes, 3 leaf B-tree nodes, one root above
        data[i] = rng.standard_normal((5, 3, 3))
        w.dataset('/many/item_%d/mechanical/F' % i, data[i], unit='1', systems=['[0 1 1]', '(1 1 1)x'])
    ct = np.zeros((4, 1), dtype=[('label', 'S9'), ('entry', '<i8')])
    ct['label'], ct['entry'][:, 0] = b'pheno', np.arange(4)
    w.dataset('/cell_to/phase', ct)
    w.dataset('/ints', np.arange(-3, 9, dtype=np.int64).reshape(3, 4))
    w.dataset('/u8', np.arange(7, dtype=np.uint8))
    w.dataset('/f4', np.linspace(0, 1, 5, dtype=np.float32))
    w.dataset('/setup/x.yaml', np.array([b'hello: 1\n']))
    w.dataset('/empty', np.zeros((0, 3)))
    w.group('/hollow')
    path = str(tmp_path / 't.hdf5')
    w.save(path)
    f = H5File(path)
    assert f.keys('/') == sorted(['many', 'cell_to', 'ints', 'u8', 'f4', 'setup', 'empty', 'hollow', 'geometry'])
    assert f.keys('/hollow') == [] and f.is_group('/hollow') and not f.is_group('/ints')
    assert len(f.keys('/many')) == 700
    a = f.attrs('/')
    assert a['version'] == 3 and a['version'].dtype == np.int32 and a['creator'] == 'someone' and a['ratio'] == 0.25
    assert (f.attrs('/geometry')['cells'] == [6, 7, 8]).all() and (f.attrs('/geometry')['size'] == [.75, .875, 1.]).all()
    for i in (0, 1, 99, 350, 699):
        assert np.array_equal(f.read('/many/item_%d/mechanical/F' % i), data[i])
    at = f.attrs('/many/item_5/mechanical/F')
    assert at['unit'] == '1' and list(at['systems']) == ['[0 1 1]', '(1 1 1)x']
    back = f.read('/cell_to/phase')
    assert back.dtype == ct.dtype and (back == ct).all()
    assert np.array_equal(f.read('/ints'), np.arange(-3, 9).reshape(3, 4)) and f.read('/ints').dtype == np.int64
    assert np.array_equal(f.read('/u8'), np.arange(7)) and f.read('/u8').dtype == np.uint8
    assert np.array_equal(f.read('/f4'), np.linspace(0, 1, 5, dtype=np.float32))
    assert f.read('/setup/x.yaml')[0] == b'hello: 1\n' and f.read('/empty').shape == (0, 3)
    assert check_group_structure(f) == 9 + 700 * 3 + 2             # every link found through the B-tree keys


def test_structure_check_catches_a_wrong_key(tmp_path):
    w = H5Writer()
    for k in range(20):
        w.dataset('/d%02d' % k, np.arange(3.0))
    blob = bytearray(w.tobytes())
    f = H5File.__new__(H5File)
    f.b = bytes(blob)
    f.root = f._symbol_entry(56)
    assert check_group_structure(f) == 20
    bt = f.root['btree']
    blob[bt + 24 + 16:bt + 24 + 24] = blob[bt + 24 + 32:bt + 24 + 40]      # key 1 := key 2
    f.b = bytes(blob)
    with pytest.raises(ValueError):
        check_group_structure(f)


def test_writer_refuses_what_it_cannot_encode():
    w = H5Writer()
    with pytest.raises(NotImplementedError):
        w.dataset('/c', np.zeros(3, dtype=complex))
    w = H5Writer()
    w.dataset('/a', np.zeros(3))
    with pytest.raises(KeyError):
        w.dataset('/a', np.zeros(3))
    with pytest.raises(KeyError):
        w.dataset('/a/b', np.zeros(3))


def test_orientation_from_elastic_deformation_matches_damask_g5():
    """crystallite_orientations: O computed from F_e equals the O stored by real DAMASK next to that F_e (both phases)"""
    g5 = np.load(os.path.join(GOLDEN, 'g5_dadf5_frame.npz'))
    for ph in ('pheno_fcc', 'pheno_bcc'):
        q = dadf5.orientation_from_Fe(g5['F_e_' + ph])
        assert np.abs(q - g5['O_' + ph]).max() < 1e-13
        assert (q[:, 0] >= 0).all()


def test_result_frame_from_the_reference_deck_matches_damask_g5(tmp_path):
    """ResultWriter fed the reference's 12-grain two-phase deck writes the same root / geometry attributes and the same
    cell_to (label, entry) mappings as DAMASK did, and per-phase datasets land in DAMASK's per-phase order."""
    g5 = np.load(os.path.join(GOLDEN, 'g5_dadf5_frame.npz'))
    geo = read_vti(os.path.join(GOLDEN, '12grains6x7x8.vti'))
    mat_text = open(os.path.join(GOLDEN, '12grains6x7x8.material.yaml')).read()
    w = dadf5.ResultWriter(geo['cells'], geo['size'], yaml.safe_load(mat_text), geo['material'], origin=geo['origin'],
                           setup={'material.yaml': mat_text})
    n = int(np.prod(geo['cells']))
    # per-cell F_e assembled from the per-phase golden arrays through DAMASK's own mapping
    Fe = np.zeros((n, 3, 3))
    for ph in ('pheno_fcc', 'pheno_bcc'):
        sel = g5['phase_label'] == ph.encode()
        Fe[sel] = g5['F_e_' + ph][g5['phase_entry'][sel]]
    xi = np.arange(n * 12, dtype=float).reshape(n, 12)
    nx, ny, nz = (int(c) for c in geo['cells'])
    w.add_increment(4, float(g5['t_inc4']), {'F_e': Fe, 'O': dadf5.orientation_from_Fe(Fe), 'xi_sl': xi},
                    u_n=np.zeros((nz + 1, ny + 1, nx + 1, 3)), u_p=np.zeros((nz, ny, nx, 3)))
    path = str(tmp_path / 'job.hdf5')
    w.save(path)
    f = H5File(path)
    check_group_structure(f)
    root, g = f.attrs('/'), f.attrs('/geometry')
    assert [root['DADF5_version_major'], root['DADF5_version_minor']] == list(g5['version'])
    assert (g['cells'] == g5['cells']).all() and g['cells'].dtype == np.int32
    assert np.allclose(g['size'], g5['size'], rtol=1e-15) and np.allclose(g['origin'], g5['origin'])
    ph, ho = f.read('/cell_to/phase'), f.read('/cell_to/homogenization')
    assert ph.shape == (n, 1) and ho.shape == (n,)
    assert (ph['label'].reshape(-1) == g5['phase_label']).all() and (ph['entry'].reshape(-1) == g5['phase_entry']).all()
    assert (ho['label'] == g5['homog_label']).all() and (ho['entry'] == g5['homog_entry']).all()
    assert f.attrs('/increment_4')['t/s'] == float(g5['t_inc4'])
    assert f.keys('/increment_4') == ['geometry', 'homogenization', 'phase']
    assert f.keys('/increment_4/homogenization') == ['SX']
    for p in ('pheno_fcc', 'pheno_bcc'):
        base = '/increment_4/phase/%s/mechanical/' % p
        assert np.array_equal(f.read(base + 'F_e'), g5['F_e_' + p])
        assert np.abs(f.read(base + 'O') - g5['O_' + p]).max() < 1e-13
        assert f.attrs(base + 'O')['lattice'] == g5['lattice_' + p].item().decode()
        assert f.attrs(base + 'F_e')['unit'] == '1'
        assert f.read(base + 'xi_sl').shape == (g5['F_e_' + p].shape[0], 12)
    assert f.read('/increment_4/geometry/u_n').shape == ((nx + 1) * (ny + 1) * (nz + 1), 3)
    assert f.read('/setup/material.yaml')[0].decode() == mat_text


# ---- multi-rank gather (gloo, world_size 2): z-slab pieces arrive on rank 0 in whole-grid order
class _FakeLib:
    def dmk_max_state(self, ctx):
        return 24


class _FakeGrid:
    """stands in for GridContext on one z-slab: hands out slices of whole-grid arrays every rank can regenerate"""
    lib, ctx = _FakeLib(), None

    def __init__(self, cells, z0, z1, seed):
        nx, ny, nz = cells
        self.world_size = 1 if (z0, z1) == (0, nz) else nz // (z1 - z0)
        rng = np.random.default_rng(seed)
        n = nx * ny * nz
        self.whole = {k: np.eye(3).reshape(9) + 0.01 * rng.standard_normal((n, 9)) for k in ('F', 'Fp', 'Fe', 'Lp', 'S')}
        self.whole['state'] = rng.random((n, 24))
        self.x_n = rng.standard_normal((nz + 1, ny + 1, nx + 1, 3))
        self.x_p = rng.standard_normal((nz, ny, nx, 3))
        self.z0, self.z1, self.nz_local, self.n_local = z0, z1, z1 - z0, (z1 - z0) * ny * nx
        self.sl = slice(z0 * ny * nx, z1 * ny * nx)

    def get_state(self, k, comps):
        return self.whole[k][self.sl].copy()

    def to_device(self, a):
        return a

    def update_coords(self, F):                       # called once with F = I for the undeformed coordinates
        return np.zeros_like(self.x_n[self.z0:self.z1 + 1]), np.zeros_like(self.x_p[self.z0:self.z1])


class _FakeSolver:
    def __init__(self, g):
        self.g = g

    def update_coords(self):
        g = self.g
        return g.x_n[g.z0:g.z1 + 1].copy(), g.x_p[g.z0:g.z1].copy()


def _writer():
    geo = read_vti(os.path.join(GOLDEN, '12grains6x7x8.vti'))
    mat = yaml.safe_load(open(os.path.join(GOLDEN, '12grains6x7x8.material.yaml')))
    return dadf5.ResultWriter(geo['cells'], geo['size'], mat, geo['material'], origin=geo['origin']), tuple(int(c) for c in geo['cells'])


def _gather_worker(rank, world, port, path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from solver_cpp_b200 import parallel
    w, cells = _writer()
    z0, z1 = parallel.slab_range(cells[2], world, rank)
    w.record(dict(inc=3, t=1.5), _FakeSolver(_FakeGrid(cells, z0, z1, 5)))
    if rank == 0:
        w.save(path)
        # a whole-grid solver recorded by ONE rank inside the initialised group must not enter a collective
        # (tests/mgpu_parity.py compares the slab solver with a single-GPU solver exactly like this)
        w1, _ = _writer()
        w1.record(dict(inc=3, t=1.5), _FakeSolver(_FakeGrid(cells, 0, cells[2], 5)))
        assert w1.increments == [3]
    else:
        assert w.increments == []
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gather_writes_the_same_file_content_as_one_rank(tmp_path):
    import torch.multiprocessing as mp
    path2, path1 = str(tmp_path / 'two.hdf5'), str(tmp_path / 'one.hdf5')
    mp.spawn(_gather_worker, args=(2, 29631, path2), nprocs=2, join=True)
    w, cells = _writer()
    w.record(dict(inc=3, t=1.5), _FakeSolver(_FakeGrid(cells, 0, cells[2], 5)))
    w.save(path1)
    a, b = H5File(path1), H5File(path2)
    check_group_structure(b)
    n_checked = 0
    for grp in ('/increment_3/geometry', '/increment_3/phase/pheno_fcc/mechanical', '/increment_3/phase/pheno_bcc/mechanical'):
        assert a.keys(grp) == b.keys(grp)
        for k in a.keys(grp):
            assert np.array_equal(a.read(grp + '/' + k), b.read(grp + '/' + k)), (grp, k)
            n_checked += 1
    assert n_checked == 2 + 8 + 8
    assert b.read('/increment_3/geometry/u_n').shape == ((cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1), 3)


def test_run_job_writes_the_increments_f_out_selects(tmp_path, monkeypatch):
    """job.run_job(result=...) with the device solver replaced by a stand-in: the file holds the increments the load step's
    f_out selects (the reference's real result of this deck, f_out 4: increments 4, 8, ...), the input decks under /setup"""
    from solver_cpp_b200 import job, mech
    files = [os.path.join(GOLDEN, f) for f in ('12grains6x7x8.vti', '12grains6x7x8.material.yaml', '6grains6x7x8_tensionY.load.yaml')]
    load = yaml.safe_load(open(files[2]))
    assert load['loadstep'][0]['f_out'] == 4
    _, cells = _writer()

    class Stand(_FakeSolver):
        def init(self):
            pass

        def forward(self, *a):
            pass

        def calculate_solution(self):
            return dict(converged=True, stag_converged=True, terminally_ill=False, required_iterations=2)

        def get33(self, name):
            return np.eye(3)

        def get(self, name, n=1):
            return np.zeros(n)

    g = _FakeGrid(cells, 0, cells[2], 9)
    g.materialpoint_forward = lambda: None
    steps = mech.parse_load(load)
    steps[0]['N'], steps[0]['t'] = 9, steps[0]['t'] * 9 / steps[0]['N']
    monkeypatch.setattr(job, 'load_job', lambda *a, **k: (Stand(g), steps))
    path = str(tmp_path / 'out.hdf5')
    recs = job.run_job(*files, result=path)
    assert len(recs) == 9
    f = H5File(path)
    check_group_structure(f)
    assert f.keys('/') == ['cell_to', 'geometry', 'increment_4', 'increment_8', 'setup']
    assert f.keys('/setup') == ['12grains6x7x8.material.yaml', '6grains6x7x8_tensionY.load.yaml']
    assert f.attrs('/increment_8')['t/s'] == pytest.approx(recs[7]['t'])
    assert f.read('/setup/6grains6x7x8_tensionY.load.yaml')[0].decode() == open(files[2]).read()
    assert 'run_job' in f.attrs('/')['call']


def test_reader_places_per_phase_data_back_into_cell_order(tmp_path):
    g5 = np.load(os.path.join(GOLDEN, 'g5_dadf5_frame.npz'))
    w, cells = _writer()
    n = int(np.prod(cells))
    rng = np.random.default_rng(2)
    F = rng.standard_normal((n, 3, 3))
    xi = rng.random((n, 12))
    w.add_increment(4, 2.0, {'F': F, 'xi_sl': xi}, u_p=np.zeros((n, 3)))
    w.add_increment(12, 6.0, {'F': 2 * F})
    path = str(tmp_path / 'r.hdf5')
    w.save(path)
    r = dadf5.ResultReader(path)
    assert r.version == (1, 0) and r.increments == [4, 12] and r.times == [2.0, 6.0] and r.cells == cells
    assert r.phases == ['pheno_bcc', 'pheno_fcc'] and (r.phase_entry == g5['phase_entry']).all()
    assert np.array_equal(r.place('F', 4), F) and np.array_equal(r.place('F', 12), 2 * F)
    assert np.array_equal(r.place('xi_sl', 4), xi)
    assert set(r.get('F', 4)) == {'pheno_bcc', 'pheno_fcc'} and r.get('F', 4)['pheno_fcc'].shape == (167, 3, 3)
    assert r.datasets(4, 'pheno_fcc') == ['F', 'xi_sl'] and r.geometry('u_p', 4).shape == (n, 3)
    with pytest.raises(KeyError):
        r.place('P', 4)


def test_reader_on_a_file_written_by_damask():
    """the DAMASK-written fixture (Marc solver, one phase): same reader, same placement"""
    r = dadf5.ResultReader(os.path.join(GOLDEN, 'damask_written_marc_job.hdf5'))
    assert r.version == (1, 0) and r.increments == [1, 2] and r.times[1] == 100.0 and r.cells is None
    assert r.phases == ['Aluminum_J2isotropic']
    xi = r.place('xi', 2)
    assert xi.shape == (8,) and np.array_equal(xi, r.get('xi', 2)['Aluminum_J2isotropic']) and (xi > 0).all()


def test_streaming_writer_produces_the_same_file_as_the_in_memory_one(tmp_path):
    def fill(w):
        rng = np.random.default_rng(1)
        w.group('/', version=np.int32(1))
        for i in range(40):
            w.dataset('/increment_%d/phase/p/mechanical/F' % i, rng.standard_normal((7, 3, 3)), unit='1')
        w.dataset('/odd', np.arange(5, dtype=np.uint8))                       # 5 bytes: padding between datasets
        w.dataset('/empty', np.zeros((0, 2)))
        w.group('/increment_3', **{'t/s': 1.5})
        return w
    path = str(tmp_path / 's.hdf5')
    ws = fill(H5Writer(path))
    ws.save()
    image = fill(H5Writer()).tobytes()
    assert open(path, 'rb').read() == image
    f = H5File(path)
    check_group_structure(f)
    assert f.read('/odd').tolist() == [0, 1, 2, 3, 4] and f.read('/empty').shape == (0, 2)
    with pytest.raises(ValueError):
        ws.dataset('/late', np.zeros(3))
    with pytest.raises(ValueError):
        H5Writer(str(tmp_path / 'x.hdf5')).tobytes()


def test_g7_multi_constituent_mappings_equal_a_file_damask_wrote(tmp_path):
    """Golden set G7: real DAMASK's /cell_to/phase [24, 8] and /cell_to/homogenization of a job with 8 constituents per cell
    (python/tests/resources/Result/4grains2x4x3_compressionY.hdf5, extracted by tests/golden/make_golden_g7.py).  The writer,
    given the same geometry and the constituents' phases, must produce the same (label, entry) for every constituent of every
    cell -- the entry counters of material.f90:177-203 -- and the reader must hand the constituents back as an axis."""
    from solver_cpp_b200 import vti
    from solver_cpp_b200.hdf5_min import H5File
    g7 = np.load(os.path.join(GOLDEN, 'g7_multi_constituent_mapping.npz'))
    geo = vti.read_vti(os.path.join(GOLDEN, 'damask_written_4grains2x4x3.vti'))
    assert 'DAMASK_grid' in str(g7['creator'])
    material = {'homogenization': {'RGC': {'N_constituents': 8, 'mechanical': {'type': 'RGC'}}},
                'phase': {p: {'lattice': 'cF'} for p in ('A', 'B', 'C')},
                'material': [{'homogenization': str(h), 'constituents': [{'phase': str(p), 'v': float(v), 'O': [1.0, 0.0, 0.0, 0.0]} for p, v in zip(ps, vs)]}
                             for h, ps, vs in zip(g7['material_homogenization'], g7['material_phase'], g7['material_v'])]}
    w = dadf5.ResultWriter(geo['cells'], geo['size'], material, geo['material'], origin=geo['origin'])
    n = 24
    rng = np.random.default_rng(0)
    F = rng.standard_normal((n, 8, 3, 3))
    w.add_increment(10, 5.0, {'F': F})
    path = str(tmp_path / 'rgc_layout.hdf5')
    w.save(path)
    f = H5File(path)
    ph, ho = f.read('/cell_to/phase'), f.read('/cell_to/homogenization')
    assert ph.shape == (24, 8)
    assert np.array_equal(ph['label'].astype('S8'), g7['cell_to_phase_label']) and np.array_equal(ph['entry'], g7['cell_to_phase_entry'])
    assert np.array_equal(ho['label'].astype('S8'), g7['cell_to_homogenization_label']) and np.array_equal(ho['entry'], g7['cell_to_homogenization_entry'])
    for p in ('A', 'B', 'C'):
        assert f.read('/increment_10/phase/%s/mechanical/F' % p).shape[0] == int(g7['rows_' + p])
    r = dadf5.ResultReader(path)
    assert r.n_const == 8 and np.array_equal(r.place('F', 10), F)


REF_RESULT = '/root/reference/python/tests/resources/Result'


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_RESULT, '4grains2x4x3_compressionY.hdf5')),
                    reason='needs the reference tree (present in the build container only; 1 MB result file not copied into the repo)')
def test_reader_places_constituents_like_damask_own_python_library():
    """The reference stores what its own damask.Result.place returns for the 8-constituent result file
    4grains2x4x3_compressionY.hdf5 (python/tests/test_Result.py:530-554, pickles under resources/Result/place/: constituent 1 of
    every cell, masked where a phase lacks the dataset).  ResultReader.place(...)[:, 1] on the same file must be that array:
    exactly at increment 0 -- where F_e of the other seven constituents differs by O(1), so the placement is unambiguous -- and
    within the drift of the stored pickles later on (they stem from a re-run of the job: 5e-15 at increment 2, 2e-9 at 10)."""
    import bz2
    import pickle
    r = dadf5.ResultReader(os.path.join(REF_RESULT, '4grains2x4x3_compressionY.hdf5'))
    assert r.n_const == 8 and r.n_cells == 24 and r.increments == [0, 2, 4, 6, 8, 10]
    with bz2.BZ2File(os.path.join(REF_RESULT, 'place', 'test_place[0].pbz2')) as f:
        ref = pickle.load(f)
    tol = {0: 0.0, 2: 1e-13, 4: 1e-11, 6: 1e-10, 8: 1e-9, 10: 1e-8}
    for inc in r.increments:
        for name in ('F', 'F_e', 'F_p', 'L_p'):
            a = ref['increment_%d' % inc]['phase'][name]
            mask = np.ma.getmaskarray(a)
            ours = r.place(name, inc)
            assert ours.shape == (24, 8, 3, 3)
            assert np.isnan(ours[:, 1][mask]).all() and not np.isnan(ours[:, 1][~mask]).any()      # phase C carries F only
            assert np.abs(ours[:, 1][~mask] - a.data[~mask]).max() <= tol[inc], (inc, name)
            if name == 'F_e' and inc == 0:
                others = [np.nanmax(np.abs(ours[:, k] - a.data)) for k in (0, 3, 4, 6, 7)]
                assert min(others) > 1.0                                                           # any other constituent is O(1) away
    with bz2.BZ2File(os.path.join(REF_RESULT, 'place', 'test_place[7].pbz2')) as f:
        ref = pickle.load(f)
    for inc in r.increments:
        for name in ('u_p', 'u_n'):
            assert np.abs(r.geometry(name, inc) - ref['increment_%d' % inc]['geometry'][name]).max() <= 1e-10


def test_reader_strips_space_padded_labels_of_newer_damask_files(tmp_path):
    """DAMASK 3.0.0-alpha7-880 and later write the (label, entry) mappings from Fortran character arrays: labels of unequal length
    are padded with SPACES (the older files, and this repo's writer, pad with NULs).  libhdf5 strips either; so must the reader --
    found on the reference's 2phase_irregularGrid result file ('Matrix     ' next to 'Precipitate'), here on a file with hand-padded
    labels."""
    from solver_cpp_b200.hdf5_min import H5Writer
    w = H5Writer()
    w.group('/', DADF5_version_major=np.int32(1), DADF5_version_minor=np.int32(0))
    w.group('/geometry', cells=np.array([2, 1, 1], np.int32), size=np.ones(3), origin=np.zeros(3))
    m = np.zeros((2, 1), dtype=[('label', 'S11'), ('entry', '<i8')])
    m['label'][0, 0] = b'Matrix     '; m['label'][1, 0] = b'Precipitate'
    w.dataset('/cell_to/phase', m)
    w.group('/increment_3', **{'t/s': 1.5})
    w.dataset('/increment_3/phase/Matrix/mechanical/F', np.full((1, 3, 3), 1.0))
    w.dataset('/increment_3/phase/Precipitate/mechanical/F', np.full((1, 3, 3), 2.0))
    path = str(tmp_path / 'padded.hdf5')
    w.save(path)
    r = dadf5.ResultReader(path)
    assert r.phases == ['Matrix', 'Precipitate']
    assert np.array_equal(r.place('F', 3)[:, 0, 0], [1.0, 2.0])
