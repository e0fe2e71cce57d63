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


# The reference's own 16^3 deck (BASELINE.json configs[0], "20grains16x16x16 tensionX"): seed positions of
# examples/grid/20grains.seeds (in units of the box; the geometry is their periodic Voronoi tessellation on 16^3 cells of a unit
# cube) and the 20 orientations of examples/grid/material.yaml (unit quaternions; its phase is PHASES['Al_pheno']).  Pinned:
# tests/test_oracle_golden.py compares the tessellation with the 16^3 VTI text the reference's own reader test embeds
# (test_discretization_grid.cpp:103-124 -> tests/golden/g6_unit_test_geometry.vti: all 4096 cells equal) and the tables with
# the decks under tests/golden.
REFERENCE_20GRAINS_SEEDS = (
    (0.375488, 0.161813, 0.891040),
    (0.187988, 0.849313, 0.953540),
    (0.750488, 0.349313, 0.953540),
    (0.312988, 0.099313, 0.891040),
    (0.562988, 0.536813, 0.703540),
    (0.750488, 0.036813, 0.641040),
    (0.812988, 0.724313, 0.016040),
    (0.187988, 0.161813, 0.266040),
    (0.562988, 0.286813, 0.203540),
    (0.937988, 0.599313, 0.953540),
    (0.937988, 0.349313, 0.578540),
    (0.812988, 0.661813, 0.453540),
    (0.125488, 0.411813, 0.203540),
    (0.625488, 0.224313, 0.578540),
    (0.687988, 0.161813, 0.891040),
    (0.625488, 0.161813, 0.766040),
    (0.437988, 0.724313, 0.266040),
    (0.437988, 0.099313, 0.641040),
    (0.937988, 0.911813, 0.828540),
    (0.437988, 0.411813, 0.516040),
)
REFERENCE_20GRAINS_O = (
    (1.0, 0.0, 0.0, 0.0),
    (0.7936696712125002, -0.28765777461664166, -0.3436487135089419, 0.4113964260949434),
    (0.3986143167493579, -0.7014883552495493, 0.2154871765709027, 0.5500781677772945),
    (0.28645844315788244, -0.022571491243423537, -0.467933059311115, -0.8357456192708106),
    (0.33012772942625784, -0.6781865350268957, 0.6494525351030648, 0.09638521992649676),
    (0.43596817439583935, -0.5982537129781701, 0.046599032277502436, 0.6707106499919265),
    (0.169734823419553, -0.699615227367322, -0.6059581215838098, -0.33844257746495854),
    (0.9698864809294915, 0.1729052643205874, -0.15948307917616958, 0.06315956884687175),
    (0.46205660912967883, 0.3105054068891252, -0.617849551030653, 0.555294529545738),
    (0.4512443497461787, -0.7636045534540555, -0.04739348426715133, -0.45939142396805815),
    (0.2161856212656443, -0.6581450184826598, -0.5498086209601588, 0.4667112513346289),
    (0.8753220715350803, -0.4561599367657419, -0.13298279533852678, -0.08969369719975541),
    (0.11908260752431069, 0.18266024809834172, -0.7144822594012615, -0.664807992845101),
    (0.751104669484278, 0.5585633382623958, -0.34579336397009175, 0.06538900566860861),
    (0.08740438971703973, 0.8991264096610437, -0.4156704205935976, 0.10559485570696363),
    (0.5584325870096193, 0.6016408353068798, -0.14280340445801173, 0.5529814994483859),
    (0.4052725440888093, 0.25253073423599154, 0.5693263597910454, -0.669215876471182),
    (0.7570164606888676, 0.15265448024694664, -0.5998021466848317, 0.20942796551297105),
    (0.6987659297138081, -0.132172211261028, -0.19693254724422338, 0.6748883269678543),
    (0.7729330445886478, 0.21682179052722322, -0.5207379472917645, 0.2905078484066341),
)


def reference_20grains(cells=(16, 16, 16), z_range=None):
    """-> (material dict, material_at of the slab, size) of the reference's example deck examples/grid/{20grains.seeds,
    material.yaml}; other `cells` re-discretise the same 20 seeds"""
    size = (1.0, 1.0, 1.0)
    material = {'homogenization': {'SX': {'N_constituents': 1, 'mechanical': {'type': 'pass'}}},
                'phase': {'Aluminum': PHASES['Al_pheno']},
                'material': [{'homogenization': 'SX', 'constituents': [{'phase': 'Aluminum', 'v': 1.0, 'O': list(q)}]}
                             for q in REFERENCE_20GRAINS_O]}
    seeds = np.asarray(REFERENCE_20GRAINS_SEEDS, float) * np.asarray(size)
    return material, voronoi_material_at(cells, size, seeds, z_range), size


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
