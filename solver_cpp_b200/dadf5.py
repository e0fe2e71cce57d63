"""DADF5 result files (the reference's HDF5 output: /root/reference/src/result.f90:75-155 file frame and increments,
src/discretization.f90:100-116 displacements, src/phase_mechanical.f90:917-965 per-phase tensors, src/material.f90 cell_to
mappings; layout as read by /root/reference/python/damask/_result.py:100-125).  Host-side output plumbing
on top of hdf5_min.H5Writer; nothing here is on the hot path.

    w = ResultWriter(cells, size, material, material_at, setup={'material.yaml': text, ...})
    mech.loop_over_loadcase(solver, steps, callback=w.callback(every=4))
    w.save('job.hdf5')

Layout written (everything damask.Result reads for a grid job; C = constituents per cell, 1 unless isostrain):
    / attrs DADF5_version_major/minor, creator, created, call
    /geometry attrs cells, size, origin
    /cell_to/phase [N,C], /cell_to/homogenization [N]: compound (label, entry); entries of a phase count its material points
        cell by cell, constituent by constituent (material.f90:177-203)
    /setup/<file name>: the input decks as strings
    /increment_<n> attr t/s
        geometry/u_n [(nx+1)(ny+1)(nz+1),3], u_p [N,3]
        phase/<label>/mechanical/F, F_e, F_p, L_p, P, S [n_ph,3,3], O [n_ph,4], xi_sl (phenopowerlaw) [n_ph,N_sl]
        homogenization/<label>/mechanical/ (empty, like the reference for 'pass')
Strings are fixed-length (the reference writes variable-length ones); h5py returns them as bytes."""
import time

import numpy as np

from .hdf5_min import H5Writer

VERSION = 'solver_cpp_b200'

_DESCR = {                                                             # phase_mechanical.f90:933-958
    'F': ('deformation gradient', '1'), 'F_e': ('elastic deformation gradient', '1'), 'F_p': ('plastic deformation gradient', '1'),
    'L_p': ('plastic velocity gradient', '1/s'), 'P': ('first Piola-Kirchhoff stress', 'Pa'),
    'S': ('second Piola-Kirchhoff stress', 'Pa'), 'O': ('crystal orientation as quaternion q_0 (q_1 q_2 q_3)', 'q_0 (q_1 q_2 q_3)'),
    'xi_sl': ('resistance against plastic slip', 'Pa'), 'gamma_sl': ('plastic shear', '1'),
    'u_n': ('displacements of the nodes', 'm'), 'u_p': ('displacements of the materialpoints (cell centers)', 'm'),
}


def rotational_part(F):
    """R of F = R U for a stack of 3x3 matrices (math.f90:1117-1156 computes the same R in closed form)"""
    U, _, Vt = np.linalg.svd(F)
    return U @ Vt


def om2qu(om):
    """passive rotation matrices [n,3,3] -> unit quaternions [n,4], q_0 >= 0, P = -1 (rotations.f90 convention)"""
    om = np.asarray(om, float)
    t = np.trace(om, axis1=-2, axis2=-1)
    q = np.empty(om.shape[:-2] + (4,))
    q[..., 0] = 0.5 * np.sqrt(np.maximum(0.0, 1.0 + t))
    d = [om[..., 0, 0], om[..., 1, 1], om[..., 2, 2]]
    q[..., 1] = 0.5 * np.sqrt(np.maximum(0.0, 1.0 + d[0] - d[1] - d[2]))
    q[..., 2] = 0.5 * np.sqrt(np.maximum(0.0, 1.0 - d[0] + d[1] - d[2]))
    q[..., 3] = 0.5 * np.sqrt(np.maximum(0.0, 1.0 - d[0] - d[1] + d[2]))
    q[..., 1] *= np.where(om[..., 2, 1] < om[..., 1, 2], -1.0, 1.0)
    q[..., 2] *= np.where(om[..., 0, 2] < om[..., 2, 0], -1.0, 1.0)
    q[..., 3] *= np.where(om[..., 1, 0] < om[..., 0, 1], -1.0, 1.0)
    q[..., 1:] *= -1.0                                                  # P = -1
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def orientation_from_Fe(Fe):
    """crystallite_orientations (phase.f90:577-590): O = fromMatrix(transpose(rotationalPart(F_e)))"""
    return om2qu(np.swapaxes(rotational_part(Fe), -1, -2))


def gather_slabs(local, group=None, device=None):
    """z-slab pieces of per-cell / per-node arrays -> whole-grid arrays on rank 0 (None elsewhere).  `local`: name -> array
    whose first axis runs over this rank's cells (x fastest, z slowest: slabs concatenate; equal slabs, parallel.slab_range);
    the entry 'u_n', if present, is [nz_local + 1, ny + 1, nx + 1, 3]: rank 0 keeps the lower nz_local node layers of every
    rank and the top layer of the last one.  One all_gather of equally shaped tensors per array (on `device` for an NCCL
    group, host memory for gloo) -- result output is not on the hot path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == 'nccl':
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    else:
        dev = torch.device('cpu')
    whole = {}
    for k in sorted(local):
        t = torch.from_numpy(np.ascontiguousarray(local[k], dtype=np.float64)).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        if rank == 0:
            parts = [p.cpu().numpy() for p in parts]
            if k == 'u_n':
                parts = [p[:-1] for p in parts[:-1]] + [parts[-1]]
            whole[k] = np.concatenate(parts, axis=0)
    return whole if rank == 0 else None


class ResultWriter:
    def __init__(self, cells, size, material, material_at, origin=(0.0, 0.0, 0.0), setup=None, call='solver_cpp_b200', path=None):
        """material: parsed material.yaml; material_at: material index of every cell of the WHOLE grid, x fastest.
        path: stream the datasets to that file as the increments arrive (save() finishes it) instead of holding the
        result in host memory until save(path)"""
        self.cells = tuple(int(c) for c in cells)
        self.size = np.asarray(size, float)
        n = int(np.prod(self.cells))
        material_at = np.asarray(material_at).reshape(-1)
        if material_at.size != n:
            raise ValueError('material_at has %d entries for %d cells' % (material_at.size, n))
        mats = material['material']
        counts = {len(mats[m]['constituents']) for m in np.unique(material_at)}
        if len(counts) != 1:
            raise NotImplementedError('DADF5 output of materials with different numbers of constituents in one grid')
        self.n_const = nc = counts.pop()
        # phase label of every material point, [N, C]: constituent co of cell ce
        lab = np.array([[c['phase'] for c in m['constituents']][:nc] + [''] * max(0, nc - len(m['constituents'])) for m in mats], dtype=object)
        self.phase_label = lab.reshape(len(mats), nc)[material_at]
        homog_label = np.array([m['homogenization'] for m in mats], dtype=object)[material_at]
        self.phases = list(material['phase'].keys())
        self.lattice = {k: (v or {}).get('lattice') for k, v in material['phase'].items()}
        self.n_sl = {}
        for k, v in material['phase'].items():
            pl = ((v or {}).get('mechanical') or {}).get('plastic') or {}
            if pl.get('type') == 'phenopowerlaw':
                self.n_sl[k] = int(np.sum(pl.get('N_sl', [0])))
        self.homogs = sorted(set(homog_label))
        self.w = w = H5Writer(path)
        now = time.strftime('%Y-%m-%d %H:%M:%S%z')
        self._stamp = dict(creator=VERSION, created=now)
        w.group('/', DADF5_version_major=np.int32(1), DADF5_version_minor=np.int32(0), creator=call + ' ' + VERSION, created=now,
                call=call)
        w.group('/geometry', cells=np.asarray(self.cells, np.int32), size=self.size, origin=np.asarray(origin, float))
        w.group('/cell_to', description='mappings to place data in space')
        w.group('/setup', description='input data used to run the simulation')
        self.cells_of_phase = {}
        # material.f90:177-203 counts a phase's entries over the cells and, within a cell, over its constituents: the flat
        # (cell, constituent) order; cells_of_phase holds each phase's flat indices ce * C + co
        ph_map = self._mapping(self.phase_label.reshape(-1), self.cells_of_phase)
        w.dataset('/cell_to/phase', ph_map.reshape(n, nc), description='cell ID and constituent ID to phase results', **self._stamp)
        w.dataset('/cell_to/homogenization', self._mapping(homog_label, {}), description='cell ID to homogenization results',
                  **self._stamp)
        for name, text in (setup or {}).items():
            w.dataset('/setup/' + name, np.array([text.encode() if isinstance(text, str) else text]), **self._stamp)
        self.increments = []

    @staticmethod
    def _mapping(labels, members):
        """(label, entry) per cell: entry = running index of the cell among the cells with the same label (material.f90)"""
        width = max(len(l.encode()) for l in labels)
        out = np.zeros(len(labels), dtype=[('label', 'S%d' % width), ('entry', '<i8')])
        for l in sorted(set(labels)):
            sel = np.nonzero(labels == l)[0]
            members[l] = sel
            out['label'][sel] = l.encode()
            out['entry'][sel] = np.arange(sel.size)
        return out

    def add_increment(self, inc, t, fields, u_n=None, u_p=None):
        """fields: name -> [N, ...] array in cell order (x fastest), tensors as natural [N,3,3]; split per phase here.  With C > 1
        constituents per cell: [N, C, ...], constituent co of cell ce at [ce, co].
        xi_sl / gamma_sl may be wider than a phase's N_sl (padded state arrays): the phase's leading columns are kept."""
        n = int(np.prod(self.cells))
        nc = self.n_const
        for name, a in fields.items():                                     # refuse before anything of the increment is written
            a = np.asarray(a)
            if a.shape[0] != n or (nc > 1 and (a.ndim < 2 or a.shape[1] != nc)):
                raise ValueError('%s: shape %s for %d cells with %d constituent(s)' % (name, a.shape, n, nc))
        base = '/increment_%d' % inc
        self.w.group(base, **{'t/s': float(t)})
        self.w.group(base + '/geometry')
        for name, a in (('u_n', u_n), ('u_p', u_p)):
            if a is not None:
                self._dataset(base + '/geometry/' + name, np.asarray(a, float).reshape(-1, 3), name)
        for ph in self.phases:
            sel = self.cells_of_phase.get(ph)
            if sel is None:
                continue
            grp = '%s/phase/%s/mechanical' % (base, ph)
            self.w.group(grp)
            for name, a in fields.items():
                a = np.asarray(a)
                if nc > 1:
                    a = a.reshape((n * nc,) + a.shape[2:])                 # flat (cell, constituent) order
                extra = {}
                if name in ('xi_sl', 'gamma_sl'):
                    if ph not in self.n_sl:
                        continue
                    a = a[:, :self.n_sl[ph]]
                if name == 'O' and self.lattice.get(ph):
                    extra['lattice'] = self.lattice[ph]
                self._dataset(grp + '/' + name, a[sel], name, **extra)
        for h in self.homogs:
            self.w.group('%s/homogenization/%s/mechanical' % (base, h))
        self.increments.append(int(inc))

    def _dataset(self, path, a, name, **extra):
        d = _DESCR.get(name)
        attrs = dict(self._stamp)
        if d:
            attrs.update(description=d[0], unit=d[1])
        attrs.update(extra)
        self.w.dataset(path, a, checksum=True, **attrs)                 # chunked + fletcher32, like the reference's datasets

    # ---- from the device-resident solver
    def record(self, rec, solver, group=None):
        """one converged increment of mech.loop_over_loadcase.  For a z-slab solver (solver.g.world_size > 1) a collective
        call: every rank passes its solver, rank 0 (the only rank that needs to save) receives the slabs.  A whole-grid
        solver records locally even inside an initialised process group."""
        local = self.local_fields(solver)
        if getattr(solver.g, 'world_size', 1) > 1:              # the SOLVER's decomposition decides, not the process group's size
            whole = gather_slabs(local, group, getattr(solver.g, 'device', None))
        else:
            whole = local
        if whole is None:
            return
        u_n, u_p = whole.pop('u_n'), whole.pop('u_p')
        self.add_increment(rec['inc'], rec['t'], whole, u_n=u_n, u_p=u_p)

    def local_fields(self, solver):
        """this rank's cells: the per-point arrays of the constitutive update and the displacements"""
        g = solver.g
        n = g.n_local
        nc = getattr(g, 'n_constituents', 1)
        if nc != self.n_const:
            raise ValueError('solver has %d constituent(s) per cell, the result file was laid out for %d' % (nc, self.n_const))
        # per-point arrays: point k * n + cell is constituent k of the cell -> [cell, constituent, ...] (slabs concatenate)
        cellwise = (lambda a: a) if nc == 1 else (lambda a: np.ascontiguousarray(np.swapaxes(a.reshape((nc, n) + a.shape[1:]), 0, 1)))
        get = lambda k: g.get_state(k, 9).reshape(n * nc, 3, 3)
        F, Fp, Fe, Lp, S = get('F'), get('Fp'), get('Fe'), get('Lp'), get('S')
        # phase_mechanical.f90:595: P = F Fp^-1 S Fp^-T with the Fp BEFORE its determinant is normalised (:600), i.e.
        # Fp^-1 = F^-1 F_e (F_i = I); the stored, normalised F_p differs from it by ~1e-7 (so does DAMASK's own file)
        P = Fe @ S @ np.swapaxes(np.linalg.solve(F, Fe), -1, -2)
        fields = {'F': F, 'F_e': Fe, 'F_p': Fp, 'L_p': Lp, 'P': P, 'S': S, 'O': orientation_from_Fe(Fe)}
        ms = int(g.lib.dmk_max_state(g.ctx))
        if self.n_sl and ms > 0:
            fields['xi_sl'] = g.get_state('state', ms)[:, :max(self.n_sl.values())]
        fields = {k: cellwise(v) for k, v in fields.items()}
        if not hasattr(self, '_x0'):                                       # undeformed coordinates: the same routine on F = I
            nx, ny, _ = self.cells
            self._x0 = g.update_coords(g.to_device(np.broadcast_to(np.eye(3), (g.nz_local, ny, nx, 3, 3)).copy()))
        x_n, x_p = solver.update_coords()
        fields['u_n'] = x_n - self._x0[0]
        fields['u_p'] = (x_p - self._x0[1]).reshape(n, 3)
        return fields

    def callback(self, every=1, group=None):
        def cb(rec, solver):
            if rec['inc'] % every == 0:
                self.record(rec, solver, group)
        return cb

    def save(self, path=None):
        self.w.save(path)


class ResultReader:
    """The reading side, for the same files (ours or DAMASK's own): increments, times, grid, and `place` -- per-phase
    datasets put back into cell order through /cell_to/phase, what damask.Result.place does
    (/root/reference/python/damask/_result.py); several constituents per cell come back as an extra axis."""

    def __init__(self, path):
        from .hdf5_min import H5File
        self.f = f = H5File(path)
        root = f.attrs('/')
        self.version = (int(root['DADF5_version_major']), int(root['DADF5_version_minor']))
        self.increments = sorted(int(k.split('_')[1]) for k in f.keys('/') if k.startswith('increment_'))
        self.times = [float(f.attrs('/increment_%d' % i)['t/s']) for i in self.increments]
        geo = f.attrs('/geometry')
        self.cells = tuple(int(c) for c in geo['cells']) if 'cells' in geo else None        # grid solver only
        self.size = np.asarray(geo['size'], float) if 'size' in geo else None
        self.origin = np.asarray(geo['origin'], float) if 'origin' in geo else None
        m = f.read('/cell_to/phase')
        self.n_const = int(m.shape[1]) if m.ndim == 2 else 1
        n_cells = m.shape[0]
        m = m.reshape(-1)
        # flat (cell, constituent) order; with one constituent per cell simply the cells
        # labels are fixed-length strings: padded with NULs (libhdf5's C default, this repo's writer) or with spaces (Fortran's
        # character arrays, DAMASK 3.0.0-alpha7-880 and later); libhdf5 strips either on conversion to a memory type
        self.phase_label = np.array([l.split(b'\0')[0].decode().rstrip(' ') for l in m['label']], dtype=object)
        self.phase_entry = m['entry'].astype(np.int64)
        self.n_cells = n_cells
        self.phases = sorted(set(self.phase_label))

    def datasets(self, inc, phase=None, field='mechanical'):
        phase = phase or self.phases[0]
        return self.f.keys('/increment_%d/phase/%s/%s' % (inc, phase, field))

    def get(self, name, inc, field='mechanical'):
        """phase label -> dataset as stored (the phase's own order); phases without that dataset are left out"""
        out = {}
        for ph in self.phases:
            grp = '/increment_%d/phase/%s/%s' % (inc, ph, field)
            if name in self.f.keys(grp):
                out[ph] = self.f.read(grp + '/' + name)
        if not out:
            raise KeyError('%s in increment %d' % (name, inc))
        return out

    def place(self, name, inc, field='mechanical', fill=np.nan):
        """dataset of every cell in cell order (x fastest), [N, ...]; with C > 1 constituents per cell [N, C, ...]; material
        points of phases that do not have it hold `fill`"""
        per_phase = self.get(name, inc, field)
        first = next(iter(per_phase.values()))
        out = np.full((self.phase_label.size,) + first.shape[1:], fill, dtype=first.dtype if first.dtype.kind == 'f' else float)
        for ph, a in per_phase.items():
            sel = self.phase_label == ph
            out[sel] = a[self.phase_entry[sel]]
        if self.n_const > 1:
            out = out.reshape((self.n_cells, self.n_const) + out.shape[1:])
        return out

    def geometry(self, name, inc):
        return self.f.read('/increment_%d/geometry/%s' % (inc, name))
