"""Synthetic workloads of BASELINE.json's configs: periodic Voronoi polycrystals, material definitions and load cases.

Voronoi tessellation as damask.Grid.from_Voronoi_tessellation (/root/reference/python/damask/_grid.py:534-574: cell
centres -> nearest seed in a periodic KD tree) with seeds from damask.seeds.from_random
(/root/reference/python/damask/seeds.py:40-43) and uniformly random unit quaternions (w >= 0).
Phase parameters are those of the reference's example configuration files
(/root/reference/examples/config/phase/mechanical/{elastic,plastic}/*.yaml, examples/grid/material.yaml), evaluated
by solver_cpp_b200.material at T = 293.15 K."""
import numpy as np

HOOKE_AL = {'type': 'Hooke', 'C_11': 106.75e9, 'C_12': 60.41e9, 'C_44': 28.34e9}                      # examples/grid/material.yaml
HOOKE_AL_T = {'type': 'Hooke', 'C_11': 106.9e9, 'C_11,T': -3.691e7, 'C_11,T^2': -9.790e3, 'C_12': 60.5e9,
              'C_12,T': -8.412e6, 'C_12,T^2': -3.972e3, 'C_44': 28.4e9, 'C_44,T': -1.413e7, 'C_44,T^2': -3.408e3,
              'T_ref': 293.15}                                                                         # Hooke_Al.yaml
HOOKE_CU = {'type': 'Hooke', 'C_11': 168.9e9, 'C_11,T': -3.332e7, 'C_11,T^2': -1.074e4, 'C_12': 121.8e9,
            'C_12,T': -1.402e7, 'C_12,T^2': -9.003e3, 'C_44': 75.8e9, 'C_44,T': -2.555e7, 'C_44,T^2': -2.188e3,
            'T_ref': 293.15}                                                                           # Hooke_Cu.yaml
HOOKE_TWIP = {'type': 'Hooke', 'C_11': 175.0e9, 'C_12': 115.0e9, 'C_44': 135.0e9}                       # Hooke_TWIP-steel.yaml

PHENO_AL = {'type': 'phenopowerlaw', 'N_sl': [12], 'a_sl': [2.25], 'atol_xi': 1.0, 'dot_gamma_0_sl': [0.001],
            'h_0_sl-sl': [75.0e6], 'h_sl-sl': [1, 1, 1.4, 1.4, 1.4, 1.4, 1.4], 'n_sl': [20], 'xi_0_sl': [31.0e6],
            'xi_inf_sl': [63.0e6]}                                                                     # examples/grid/material.yaml
PHENO_CU = {'type': 'phenopowerlaw', 'N_sl': [12], 'dot_gamma_0_sl': [3.0e-3], 'n_sl': [20], 'a_sl': [0.6],
            'xi_0_sl': [1.6e6], 'xi_inf_sl': [96.4e6], 'h_0_sl-sl': [0.35e9],
            'h_sl-sl': [1, 1, 5.123, 0.574, 1.123, 1.123, 1]}                                          # phenopowerlaw_Cu.yaml
ISOTROPIC = {'type': 'isotropic', 'dot_gamma_0': 0.001, 'n': 20.0, 'xi_0': 0.3e6, 'xi_inf': 0.6e6, 'a': 2.0,
             'h_0': 1.0e6, 'M': 1.0, 'h': 1.0, 'dilatation': False}                                    # isotropic_free-surface.yaml
DISLOTWIN_TWIP = {'type': 'dislotwin', 'N_sl': [12], 'f_edge': [1.0], 'b_sl': [2.56e-10], 'Q_sl': [3.5e-19],
                  'p_sl': [0.325], 'q_sl': [1.55], 'B': [0.001], 'i_sl': [30.0], 'v_0': [1.4e3], 'tau_0': [5.5e8],
                  'D_a': 2.0, 'Q_cl': 3.0e-19, 'rho_mob_0': [5.0e10], 'rho_dip_0': [5.0e10],
                  'h_sl-sl': [0.122, 0.122, 0.625, 0.07, 0.137, 0.137, 0.122],
                  'N_tw': [12], 'b_tw': [1.47e-10], 'L_tw': 1.91e-7, 'i_tw': 10.0, 't_tw': [5.0e-8], 'p_tw': [7],
                  'h_tw-tw': [1.0, 1.0], 'h_sl-tw': [1.0, 1.0, 1.0],
                  'T_ref': 293.15, 'Gamma_sf': 2.833e-2, 'Gamma_sf,T': 1.214e-4, 'Gamma_sf,T^2': 1.473e-7,
                  'x_c': 1.0e-9, 'V_cs': 1.67e-29, 'D': 5.0e-5}                                        # dislotwin_TWIP-TRIP.yaml, N_tr removed

HOOKE_AISI304 = {'type': 'Hooke', 'C_11': 204.6e9, 'C_12': 137.7e9, 'C_44': 126.2e9}                     # Hooke_AISI304.yaml
KINEHARDENING_317L = {'type': 'kinehardening', 'N_sl': [12], 'xi_0': [0.070e9], 'xi_inf': [0.015e9], 'h_0_xi': [0.065e9],
                      'h_inf_xi': [0.045e9], 'chi_inf': [0.027e9], 'h_0_chi': [55e9], 'h_inf_chi': [1.3e9], 'n': [20],
                      'dot_gamma_0': [1e-4], 'h_sl-sl': [1, 1, 1, 1, 1, 1, 1]}                          # kinehardening_317L.yaml
HOOKE_W = {'type': 'Hooke', 'C_11': 524.3e9, 'C_11,T': -4.794e7, 'C_12': 205.1e9, 'C_12,T': -2.836e6, 'C_44': 160.7e9,
           'C_44,T': -1.269e7, 'T_ref': 293.15}                                                        # Hooke_W.yaml
DISLOTUNGSTEN_W = {'type': 'dislotungsten', 'N_sl': [12], 'b_sl': [2.72e-10], 'rho_mob_0': [1.0e9], 'rho_dip_0': [1.0],
                   'Q_s': [2.61154e-19], 'B': [8.3e-5], 'omega': [9.1e11], 'p_sl': [0.86], 'q_sl': [1.69], 'tau_Peierls': [2.03e9],
                   'h': [2.566e-10], 'h_sl-sl': [0.009, 0.72, 0.009, 0.05, 0.05, 0.06, 0.09], 'w': [2.992e-9],
                   'a_nonSchmid_110': [0.0, 0.56, 0.75], 'i_sl': [1], 'D': 1.0e20, 'D_a': 1.0, 'D_0': 0.0, 'f_at': 1,
                   'Q_cl': 1.0}                                                                        # dislotungsten_W.yaml

PHASES = {
    'Al_pheno': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_AL, 'plastic': PHENO_AL}},
    'Cu_pheno': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_CU, 'plastic': PHENO_CU}},
    'Al_iso': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_AL_T, 'plastic': ISOTROPIC}},
    'Al_elastic': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_AL_T}},
    'TWIP': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_TWIP, 'plastic': DISLOTWIN_TWIP}},
    'W_tungsten': {'lattice': 'cI', 'mechanical': {'elastic': HOOKE_W, 'plastic': DISLOTUNGSTEN_W}},
    'Steel_kine': {'lattice': 'cF', 'mechanical': {'elastic': HOOKE_AISI304, 'plastic': KINEHARDENING_317L}},
}

TENSION_X = {'solver': {'mechanical': 'spectral_basic'},
             'loadstep': [{'boundary_conditions': {'mechanical': {
                 'dot_F': [[1.0e-3, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                 'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]]}},
                 'discretization': {'t': 5, 'N': 10}, 'f_out': 4}]}
MIXED_SHEAR = {'solver': {'mechanical': 'spectral_basic'},
               'loadstep': [{'boundary_conditions': {'mechanical': {
                   'dot_F': [[1.0e-3, 5.0e-4, 0], [0, 'x', 0], [0, 0, 'x']],
                   'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]]}},
                   'discretization': {'t': 5, 'N': 10}, 'f_out': 4}]}

# BASELINE.json configs -> (cells, grains, phase, load)
CONFIGS = {
    'C1': dict(cells=(16, 16, 16), grains=20, phase='Al_pheno', load=TENSION_X, name='20grains16x16x16 tensionX phenopowerlaw'),
    'C2': dict(cells=(64, 64, 64), grains=50, phase='Al_iso', load=TENSION_X, name='64^3 50-grain isotropic plasticity tensionX'),
    'C3': dict(cells=(256, 256, 256), grains=500, phase='Cu_pheno', load=TENSION_X, name='256^3 500-grain phenopowerlaw Cu tensionX'),
    'C4': dict(cells=(512, 512, 512), grains=4000, phase='Cu_pheno', load=MIXED_SHEAR, name='512^3 4000-grain phenopowerlaw mixed BC'),
    'C5': dict(cells=(256, 256, 256), grains=500, phase='TWIP', load=TENSION_X, name='256^3 500-grain dislotwin TWIP steel tensionX'),
}


def random_orientations(n, rng):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1.0
    return q


def grid_size(cells):
    c = np.asarray(cells, float)
    return tuple(1.0e-5 * c / c.min())


def voronoi_material_at(cells, size, seeds, z_range=None):
    """grain index of every cell of z-slab z_range=(z0, z1) (default: whole grid), cell order x fastest"""
    from scipy.spatial import cKDTree
    nx, ny, nz = cells
    z0, z1 = (0, nz) if z_range is None else z_range
    size = np.asarray(size, float)
    tree = cKDTree(np.mod(seeds, size), boxsize=size)
    x = (np.arange(nx) + 0.5) * size[0] / nx
    y = (np.arange(ny) + 0.5) * size[1] / ny
    out = np.empty((z1 - z0, ny, nx), dtype=np.int32)
    for k in range(z0, z1):                                   # slab by slab keeps the coordinate array small
        zz = (k + 0.5) * size[2] / nz
        X, Y = np.meshgrid(x, y, indexing='xy')
        pts = np.stack([X, Y, np.full((ny, nx), zz)], axis=-1).reshape(-1, 3)
        out[k - z0] = tree.query(pts, workers=-1)[1].reshape(ny, nx)
    return out.reshape(-1)


def polycrystal(cells, grains, phase, seed=0, z_range=None):
    """-> (material dict in material.yaml form, material_at for the slab, size)"""
    rng = np.random.default_rng(seed)
    size = grid_size(cells)
    seeds = rng.random((grains, 3)) * np.asarray(size)
    quats = random_orientations(grains, rng)
    material = {'homogenization': {'SX': {'N_constituents': 1, 'mechanical': {'type': 'pass'}}},
                'phase': {phase: PHASES[phase]},
                'material': [{'homogenization': 'SX', 'constituents': [{'phase': phase, 'v': 1.0, 'O': [float(v) for v in q]}]}
                             for q in quats]}
    return material, voronoi_material_at(cells, size, seeds, z_range), size


def taylor_polycrystal(cells, grains, phases, n_constituents, seed=0, z_range=None):
    """polycrystal whose every material entry is a Taylor (isostrain) aggregate of n_constituents crystallites: random
    orientations, random volume fractions summing to 1, the phases taken in turn from `phases` (names in PHASES)
    -> (material dict in material.yaml form, material_at for the slab, size)"""
    rng = np.random.default_rng(seed)
    size = grid_size(cells)
    seeds = rng.random((grains, 3)) * np.asarray(size)
    phases = [phases] if isinstance(phases, str) else list(phases)
    mats = []
    for _ in range(grains):
        quats = random_orientations(n_constituents, rng)
        v = 0.2 + rng.random(n_constituents)
        v = v / v.sum()
        v[-1] = 1.0 - v[:-1].sum()
        mats.append({'homogenization': 'Taylor',
                     'constituents': [{'phase': phases[k % len(phases)], 'v': float(v[k]), 'O': [float(x) for x in quats[k]]}
                                      for k in range(n_constituents)]})
    material = {'homogenization': {'Taylor': {'N_constituents': n_constituents, 'mechanical': {'type': 'isostrain'}}},
                'phase': {ph: PHASES[ph] for ph in dict.fromkeys(phases)},
                'material': mats}
    return material, voronoi_material_at(cells, size, seeds, z_range), size


def config(name, seed=0, z_range=None, cells=None):
    c = dict(CONFIGS[name])
    if cells is not None:
        c['cells'] = tuple(cells)
    material, material_at, size = polycrystal(c['cells'], c['grains'], c['phase'], seed, z_range)
    c.update(material=material, material_at=material_at, size=size)
    return c
