"""ORACLE (test infrastructure only -- never imported by the product path).

Crystallographic tables for cF (fcc) and cI (bcc, slip/twin geometry only): slip / twin systems,
Schmid matrices, interaction matrices, forest projections, twin stiffness rotation, isotropic moduli.

Restates (does not copy) /root/reference/src/crystal.f90:
  systems            crystal.f90:41-95 (cF), :118-199 (cI)
  coordinate system  crystal.f90:1897-1961  (direction, normal, direction x normal; all normalised)
  Schmid slip/twin   crystal.f90:1353-1478
  interactions       crystal.f90:602-648 (cF slip-slip), :920-945 (twin-twin), :1061-1100 (slip-twin),
                     :1274-1290 (twin-slip == type 1), :1019-1055 (trans-trans), :1221-1270 (slip-trans)
  buildInteraction   crystal.f90:1853-1890:  M[reacting, acting] = values[type_table[reacting][acting]-1]
                     (type tables below are written with row = reacting, column = acting, i.e. exactly as the
                      reference prints them; the Fortran array is the transpose because of column-major reshape)
  forest projection  crystal.f90:1764-1806
  char. shear        crystal.f90:466-473  (cF/cI: sqrt(2)/2)
  C66 twin           crystal.f90:500-531  (180 deg rotation about the twin-plane normal)
  symmetrize_C66     crystal.f90:1680-1720
  isotropic mu / nu  crystal.f90:2150-2220
"""
import numpy as np

# ---------------------------------------------------------------------------------------------
# slip / twin systems: rows = [direction(3), plane normal(3)]
def _sys(txt):
    a = np.array(txt.split(), dtype=float)
    return a.reshape(-1, 6)

CF_NSLIP = [12, 6]
CF_SLIP = _sys("""
 0  1 -1  1  1  1   -1  0  1  1  1  1    1 -1  0  1  1  1
 0 -1 -1 -1 -1  1    1  0  1 -1 -1  1   -1  1  0 -1 -1  1
 0 -1  1  1 -1 -1   -1  0 -1  1 -1 -1    1  1  0  1 -1 -1
 0  1  1 -1  1 -1    1  0 -1 -1  1 -1   -1 -1  0 -1  1 -1
 1  1  0  1 -1  0    1 -1  0  1  1  0    1  0  1  1  0 -1
 1  0 -1  1  0  1    0  1  1  0  1 -1    0  1 -1  0  1  1
""")
CF_NTWIN = [12]
CF_TWIN = _sys("""
-2  1  1  1  1  1    1 -2  1  1  1  1    1  1 -2  1  1  1
 2 -1  1 -1 -1  1   -1  2  1 -1 -1  1   -1 -1 -2 -1 -1  1
-2 -1 -1  1 -1 -1    1  2 -1  1 -1 -1    1 -1  2  1 -1 -1
 2  1 -1 -1  1 -1   -1 -2 -1 -1  1 -1   -1  1  2 -1  1 -1
""")
# slip pair (1-based slip indices) that nucleates twin/trans system i   (crystal.f90:82-95)
CF_TWIN_NUCLEATION_SLIP_PAIR = np.array(
    [[2, 3], [1, 3], [1, 2], [5, 6], [4, 6], [4, 5], [8, 9], [7, 9], [7, 8], [11, 12], [10, 12], [10, 11]])

CI_NSLIP = [12, 12, 24]
CI_SLIP = _sys("""
 1 -1  1  0  1  1   -1 -1  1  0 -1 -1    1  1  1  0  1 -1   -1  1  1  0 -1  1
-1  1  1 -1  0 -1   -1 -1  1  1  0  1    1  1  1 -1  0  1    1 -1  1  1  0 -1
-1  1  1  1  1  0    1 -1  1 -1 -1  0    1  1  1  1 -1  0   -1 -1  1 -1  1  0
-1  1  1  2  1  1    1  1  1 -2  1  1    1  1 -1  2 -1  1    1 -1  1  2  1 -1
 1 -1  1  1  2  1    1  1 -1 -1  2  1    1  1  1  1 -2  1   -1  1  1  1  2 -1
 1  1 -1  1  1  2    1 -1  1 -1  1  2   -1  1  1  1 -1  2    1  1  1  1  1 -2
 1  1 -1  1  2  3    1 -1  1 -1  2  3   -1  1  1  1 -2  3    1  1  1  1  2 -3
 1 -1  1  1  3  2    1  1 -1 -1  3  2    1  1  1  1 -3  2   -1  1  1  1  3 -2
 1  1 -1  2  1  3    1 -1  1 -2  1  3   -1  1  1  2 -1  3    1  1  1  2  1 -3
 1 -1  1  2  3  1    1  1 -1 -2  3  1    1  1  1  2 -3  1   -1  1  1  2  3 -1
-1  1  1  3  1  2    1  1  1 -3  1  2    1  1 -1  3 -1  2    1 -1  1  3  1 -2
-1  1  1  3  2  1    1  1  1 -3  2  1    1  1 -1  3 -2  1    1 -1  1  3  2 -1
""")
CI_NTWIN = [12]
CI_TWIN = CI_SLIP[12:24].copy()

# ---------------------------------------------------------------------------------------------
# interaction TYPE tables, row = reacting system, column = acting system (as printed in the reference)
def _tab(txt, ncol):
    return np.array(txt.split(), dtype=int).reshape(-1, ncol)

CF_SLIP_SLIP = _tab("""
 1  2  2  4  7  5  3  5  5  4  6  7 10 11 10 11 12 13
 2  1  2  7  4  5  6  4  7  5  3  5 10 11 12 13 10 11
 2  2  1  5  5  3  6  7  4  7  6  4 12 13 10 11 10 11
 4  7  6  1  2  2  4  6  7  3  5  5 10 11 11 10 13 12
 7  4  6  2  1  2  5  3  5  6  4  7 10 11 13 12 11 10
 5  5  3  2  2  1  7  6  4  6  7  4 12 13 11 10 11 10
 3  5  5  4  6  7  1  2  2  4  7  6 11 10 11 10 12 13
 6  4  7  5  3  5  2  1  2  7  4  6 11 10 13 12 10 11
 6  7  4  7  6  4  2  2  1  5  5  3 13 12 11 10 10 11
 4  6  7  3  5  5  4  7  6  1  2  2 11 10 10 11 13 12
 5  3  5  6  4  7  7  4  6  2  1  2 11 10 12 13 11 10
 7  6  4  6  7  4  5  5  3  2  2  1 13 12 10 11 11 10
10 10 12 10 10 12 11 11 13 11 11 13  1  8  9  9  9  9
11 11 13 11 11 13 10 10 12 10 10 12  8  1  9  9  9  9
10 12 10 11 13 11 11 13 11 10 12 10  9  9  1  8  9  9
11 13 11 10 12 10 10 12 10 11 13 11  9  9  8  1  9  9
12 10 10 13 11 11 12 10 10 13 11 11  9  9  9  9  1  8
13 11 11 12 10 10 13 11 11 12 10 10  9  9  9  9  8  1
""", 18)

def _block_table(n, blk, same, other):
    t = np.full((n, n), other, dtype=int)
    for b in range(n // blk):
        t[b*blk:(b+1)*blk, b*blk:(b+1)*blk] = same
    return t

CF_TWIN_TWIN = _block_table(12, 3, 1, 2)      # coplanar twins -> 1, else 2
CF_TRANS_TRANS = _block_table(12, 3, 1, 2)
# slip (reacting, 18 rows) x twin (acting, 12 cols)
CF_SLIP_TWIN = np.vstack([_tab("""
1 1 1 3 3 3 2 2 2 3 3 3
1 1 1 3 3 3 3 3 3 2 2 2
1 1 1 2 2 2 3 3 3 3 3 3
3 3 3 1 1 1 3 3 3 2 2 2
3 3 3 1 1 1 2 2 2 3 3 3
2 2 2 1 1 1 3 3 3 3 3 3
2 2 2 3 3 3 1 1 1 3 3 3
3 3 3 2 2 2 1 1 1 3 3 3
3 3 3 3 3 3 1 1 1 2 2 2
3 3 3 2 2 2 3 3 3 1 1 1
2 2 2 3 3 3 3 3 3 1 1 1
3 3 3 3 3 3 2 2 2 1 1 1
""", 12), np.full((6, 12), 4, dtype=int)])
CF_SLIP_TRANS = CF_SLIP_TWIN.copy()
CF_TWIN_SLIP = np.ones((12, 18), dtype=int)

# cI slip-slip interaction types (Madec & Kubin 2017 labelling, crystal.f90:648-701), <111>{110} family only (the first 12 of
# the 48 systems); 0 = not offered here (the {112} and {123} families)
CI_SLIP_SLIP = np.zeros((48, 48), dtype=int)
CI_SLIP_SLIP[:12, :12] = _tab("""
 1  3  6  6  7  5  4  2  4  2  7  5
 3  1  6  6  4  2  7  5  7  5  4  2
 6  6  1  3  5  7  2  4  5  7  2  4
 6  6  3  1  2  4  5  7  2  4  5  7
 7  5  4  2  1  3  6  6  2  4  7  5
 4  2  7  5  3  1  6  6  5  7  4  2
 5  7  2  4  6  6  1  3  7  5  2  4
 2  4  5  7  6  6  3  1  4  2  5  7
 5  7  4  2  2  4  7  5  1  3  6  6
 2  4  7  5  5  7  4  2  3  1  6  6
 7  5  2  4  7  5  2  4  6  6  1  3
 4  2  5  7  4  2  5  7  6  6  3  1
""", 12)

_LATT = {
    'cF': dict(nslip=CF_NSLIP, slip=CF_SLIP, ntwin=CF_NTWIN, twin=CF_TWIN, ntrans=[12],
               slip_slip=CF_SLIP_SLIP, twin_twin=CF_TWIN_TWIN, slip_twin=CF_SLIP_TWIN,
               twin_slip=CF_TWIN_SLIP, trans_trans=CF_TRANS_TRANS, slip_trans=CF_SLIP_TRANS),
    'cI': dict(nslip=CI_NSLIP, slip=CI_SLIP, ntwin=CI_NTWIN, twin=CI_TWIN, slip_slip=CI_SLIP_SLIP),
}


def _active_rows(active, potential):
    """indices into the potential-system table of the active systems (crystal.f90:1924-1929)."""
    idx = []
    for f, n in enumerate(active):
        if n < 0 or n > potential[f]:
            raise ValueError('invalid number of active systems')
        start = int(np.sum(potential[:f]))
        idx += list(range(start, start + n))
    return np.array(idx, dtype=int)


def coordinate_system(active, potential, systems):
    """-> (n,3,3): [:,0]=direction, [:,1]=normal, [:,2]=direction x normal (all unit)."""
    rows = systems[_active_rows(active, potential)]
    d = rows[:, :3] / np.linalg.norm(rows[:, :3], axis=1, keepdims=True)
    n = rows[:, 3:] / np.linalg.norm(rows[:, 3:], axis=1, keepdims=True)
    return np.stack([d, n, np.cross(d, n)], axis=1)


def _rot_axis_angle(axis, angle_rad):
    """active rotation matrix of a vector about `axis` (Rodrigues)."""
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * K @ K


def schmid_slip(N_sl, lattice, non_schmid=None, sense=+1):
    """(n,3,3) Schmid matrices d (x) n  (+ cI non-Schmid terms, crystal.f90:1410-1428)."""
    L = _LATT[lattice]
    cs = coordinate_system(N_sl, L['nslip'], L['slip'])
    d = cs[:, 0] * float(sense)
    n = cs[:, 1]
    P = np.einsum('ai,aj->aij', d, n)
    if non_schmid is not None and lattice == 'cI' and len(non_schmid) > 0:
        coeff = np.zeros(6)
        coeff[:len(non_schmid)] = non_schmid
        fam = np.repeat(np.arange(len(N_sl)), N_sl)
        for i in range(P.shape[0]):
            if fam[i] != 0:
                continue
            # reference: R%fromAxisAngle([direction,60 deg],P=1); np = R%rotate(normal)  (passive by default,
            # with P=+1 input convention the net effect is the active rotation by -60 deg about d ... the
            # reference comment says n1 = Rot(-m,60deg) @ n)
            npv = _rot_axis_angle(-d[i], np.deg2rad(60.0)) @ n[i]
            nxd = np.cross(n[i], d[i])
            P[i] += (coeff[0] * np.outer(d[i], npv) + coeff[1] * np.outer(nxd, n[i])
                     + coeff[2] * np.outer(np.cross(npv, d[i]), npv) + coeff[3] * np.outer(n[i], n[i])
                     + coeff[4] * np.outer(nxd, nxd) + coeff[5] * np.outer(d[i], d[i]))
    return P


def schmid_twin(N_tw, lattice):
    L = _LATT[lattice]
    cs = coordinate_system(N_tw, L['ntwin'], L['twin'])
    return np.einsum('ai,aj->aij', cs[:, 0], cs[:, 1])


def slip_normal(N_sl, lattice):
    L = _LATT[lattice]
    return coordinate_system(N_sl, L['nslip'], L['slip'])[:, 1]


def characteristic_shear_twin(N_tw, lattice):
    if lattice not in ('cF', 'cI'):
        raise NotImplementedError(lattice)
    return np.full(int(np.sum(N_tw)), 0.5 * np.sqrt(2.0))


def _interaction(react_used, act_used, react_max, act_max, values, table):
    values = np.asarray(values, float)
    r = _active_rows(react_used, react_max)
    a = _active_rows(act_used, act_max)
    types = table[np.ix_(r, a)]
    if types.size and types.min() < 1:
        raise NotImplementedError('interaction types of these system families are not offered')
    if types.size and types.max() > len(values):
        raise ValueError('not enough interaction values')
    return values[types - 1]


def interaction_slip_slip(N_sl, values, lattice):
    L = _LATT[lattice]
    if 'slip_slip' not in L:
        raise NotImplementedError('slip-slip interaction for ' + lattice)
    return _interaction(N_sl, N_sl, L['nslip'], L['nslip'], values, L['slip_slip'])


def interaction_twin_twin(N_tw, values, lattice):
    L = _LATT[lattice]
    return _interaction(N_tw, N_tw, L['ntwin'], L['ntwin'], values, L['twin_twin'])


def interaction_slip_twin(N_sl, N_tw, values, lattice):
    L = _LATT[lattice]
    return _interaction(N_sl, N_tw, L['nslip'], L['ntwin'], values, L['slip_twin'])


def interaction_twin_slip(N_tw, N_sl, values, lattice):
    L = _LATT[lattice]
    return _interaction(N_tw, N_sl, L['ntwin'], L['nslip'], values, L['twin_slip'])


def interaction_trans_trans(N_tr, values, lattice):
    L = _LATT[lattice]
    return _interaction(N_tr, N_tr, L['ntrans'], L['ntrans'], values, L['trans_trans'])


def interaction_slip_trans(N_sl, N_tr, values, lattice):
    L = _LATT[lattice]
    return _interaction(N_sl, N_tr, L['nslip'], L['ntrans'], values, L['slip_trans'])


def forest_projection(N_sl, lattice, f_edge):
    """prm%forestProjection (dislotwin.f90:229-232): f_edge * |n_i . t_j| + (1-f_edge) * |n_i . d_j|,
    with f_edge spread along dim 1, i.e. indexed by the SECOND (acting, j) index."""
    L = _LATT[lattice]
    cs = coordinate_system(N_sl, L['nslip'], L['slip'])
    d, n, t = cs[:, 0], cs[:, 1], cs[:, 2]
    f_edge = np.asarray(f_edge, float)
    edge = np.abs(n @ t.T)
    screw = np.abs(n @ d.T)
    return f_edge[None, :] * edge + (1.0 - f_edge)[None, :] * screw


# ---------------------------------------------------------------------------------------------
# elasticity helpers
VOIGT = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]


def symmetrize_C66(C11, C12, C44):
    C = np.zeros((6, 6))
    for i in range(3):
        C[i, i] = C11
        C[3 + i, 3 + i] = C44
        for j in range(3):
            if i != j:
                C[i, j] = C12
    return C


def voigt66_to_3333(C66):
    C = np.zeros((3, 3, 3, 3))
    for i, (a, b) in enumerate(VOIGT):
        for j, (c, d) in enumerate(VOIGT):
            C[a, b, c, d] = C[b, a, c, d] = C[a, b, d, c] = C[b, a, d, c] = C66[i, j]
    return C


def c3333_to_voigt66(C):
    return np.array([[C[a, b, c, d] for (c, d) in VOIGT] for (a, b) in VOIGT])


def C66_twin(N_tw, C66, lattice):
    """stiffness rotated by 180 deg about each twin-plane normal (crystal.f90:526-529)."""
    L = _LATT[lattice]
    cs = coordinate_system(N_tw, L['ntwin'], L['twin'])
    C = voigt66_to_3333(C66)
    out = []
    for i in range(cs.shape[0]):
        n = cs[i, 1]
        R = 2.0 * np.outer(n, n) - np.eye(3)
        Cr = np.einsum('im,jn,ko,lp,mnop->ijkl', R, R, R, R, C)
        out.append(c3333_to_voigt66(Cr))
    return np.array(out).reshape(-1, 6, 6)


def isotropic_mu(C66, assumption='isostrain'):
    a = assumption.lower()
    if a == 'isostrain':
        return (C66[0, 0] - C66[0, 1] + C66[3, 3] * 3.0) / 5.0
    if a == 'isostress':
        return 5.0 / (4.0 / (C66[0, 0] - C66[0, 1]) + 3.0 / C66[3, 3])
    raise ValueError(assumption)


def isotropic_nu(C66, assumption='isostrain'):
    a = assumption.lower()
    if a == 'isostrain':
        K = np.sum(C66[:3, :3]) / 9.0
    elif a == 'isostress':
        K = 1.0 / np.sum(np.linalg.inv(C66)[:3, :3])
    else:
        raise ValueError(assumption)
    mu = isotropic_mu(C66, assumption)
    return (1.5 * K - mu) / (3.0 * K + mu)


def qu2om(q):
    """unit quaternion (w,x,y,z) -> rotation matrix with DAMASK's P=-1 convention (rotations.f90:423-447)."""
    w, x, y, z = [float(v) for v in q]
    qq = w * w - (x * x + y * y + z * z)
    om = np.empty((3, 3))
    om[0, 0] = qq + 2.0 * x * x
    om[1, 1] = qq + 2.0 * y * y
    om[2, 2] = qq + 2.0 * z * z
    om[0, 1] = 2.0 * (x * y - w * z)
    om[1, 2] = 2.0 * (y * z - w * x)
    om[2, 0] = 2.0 * (z * x - w * y)
    om[1, 0] = 2.0 * (y * x + w * z)
    om[2, 1] = 2.0 * (z * y + w * x)
    om[0, 2] = 2.0 * (x * z + w * y)
    om = om.T                                  # P = -1
    return om / np.cbrt(np.linalg.det(om))
