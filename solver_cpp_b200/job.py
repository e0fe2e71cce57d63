"""A DAMASK grid job from the reference's own input files: geometry (.vti), material.yaml, load.yaml, numerics.yaml --
what `DAMASK_grid -g .. -m .. -l .. -n ..` parses in main.cpp:31-101, config.cpp:247-430 and discretization_grid.cpp:8-76
of the reference -- handed to the device-resident solver, optionally with a DADF5 result file.  Host-side input / output plumbing."""
import numpy as np
import yaml

from . import material as M
from . import mech, parallel
from .vti import read_vti

_DERIVATIVE = {'continuous': 0, 'central_difference': 1, 'FWBW_difference': 2}
_DIVCORR = {'none': 0, 'size': 1, 'size+grid': 2, 'grid+size': 2}


def read_numerics(path=None):
    """numerics.yaml -> (lib.Numerics, derivative id, divergence-correction id, N_cutback_max, polarisation numerics); keys and
    defaults: Config::parse_numerics, config.cpp:360-463 and config.h:50-78 of the reference (alpha, beta, eps_abs_curl(F),
    eps_rel_curl(F) of the FFT section are what the polarisation scheme uses)"""
    d = yaml.safe_load(open(path)) if path else {}
    d = d or {}
    grid = ((d.get('solver') or {}).get('grid') or {})
    fft = grid.get('FFT') or {}
    num = M.make_numerics(((d.get('phase') or {}).get('mechanical')), grid.get('mechanical'))
    for key, table in (('derivative', _DERIVATIVE), ('divergence_correction', _DIVCORR)):
        if key in fft and str(fft[key]) not in table:
            raise ValueError('numerics: FFT.%s must be one of %s' % (key, ', '.join(table)))
    pol = dict(alpha=float(fft.get('alpha', 1.0)), beta=float(fft.get('beta', 1.0)), eps_abs_curl=float(fft.get('eps_abs_curl(F)', 1.0e-10)),
               eps_rel_curl=float(fft.get('eps_rel_curl(F)', 5.0e-4)))
    return (num, _DERIVATIVE[fft.get('derivative', 'continuous')], _DIVCORR[str(fft.get('divergence_correction', 'size+grid'))],
            int(grid.get('N_cutback_max', 3)), pol)


def load_job(geometry, material, load, numerics=None, world_size=1, rank=0, comm=None):
    """-> (MechSolverBasic on this rank's z-slab, parsed load steps).  `material` / `load` may be paths or parsed dicts.
    The solver carries `max_cut_back` (numerics: solver/grid/N_cutback_max) for loop_over_loadcase."""
    geo = read_vti(geometry)
    mat = yaml.safe_load(open(material)) if isinstance(material, str) else material
    ld = yaml.safe_load(open(load)) if isinstance(load, str) else load
    num, derivative, divcorr, max_cut_back, pol = read_numerics(numerics)
    scheme = ((ld.get('solver') or {}).get('mechanical')) or 'spectral_basic'             # damask_grid.cpp:8-12 reads the same key
    cells = tuple(int(c) for c in geo['cells'])
    if int(geo['material'].max()) >= len(mat['material']) or int(geo['material'].min()) < 0:
        raise ValueError('geometry refers to material %d, material.yaml has %d entries' % (int(geo['material'].max()), len(mat['material'])))
    z0, z1 = parallel.slab_range(cells[2], world_size, rank)
    solver = mech.make_solver(cells, tuple(float(s) for s in geo['size']), mat, np.ascontiguousarray(geo['material'][z0:z1]), numerics=num,
                              derivative=derivative, divergence_correction=divcorr, world_size=world_size, rank=rank, comm=comm,
                              scheme=scheme, **(pol if scheme == 'spectral_polarisation' else {}))
    solver.max_cut_back = max_cut_back
    return solver, mech.parse_load(ld)


def run_job(geometry, material, load, numerics=None, result=None, **kw):
    """the whole load case; one record per converged increment (mech.loop_over_loadcase).  `result`: path of a DADF5 result
    file to write (solver_cpp_b200.dadf5): the increments each load step's `f_out` selects, like DAMASK_grid.f90 does
    (every f_out-th increment of the step, labelled with the running increment number); rank 0 saves it."""
    solver, steps = load_job(geometry, material, load, numerics, **kw)
    solver.init()
    callback = writer = None
    if result is not None:
        from . import dadf5
        geo = read_vti(geometry)
        mat = yaml.safe_load(open(material)) if isinstance(material, str) else material
        setup = {}
        for path in (material, load, numerics):
            if isinstance(path, str):
                setup[path.rsplit('/', 1)[-1]] = open(path).read()
        writer = dadf5.ResultWriter(geo['cells'], geo['size'], mat, geo['material'], origin=geo['origin'], setup=setup,
                                    path=result if kw.get('rank', 0) == 0 else None,
                                    call='solver_cpp_b200.job.run_job -g %s -l %s' % (geometry, load if isinstance(load, str) else '-'))

        def callback(rec, s):
            if rec['inc_in_step'] % max(1, rec['f_out']) == 0:
                writer.record(rec, s, getattr(kw.get('comm'), 'group', None))
    # the reference finishes its result file after every increment, so what was written before a failure (no solution after
    # the last cut-back, a CUDA error) stays readable: the file's metadata and superblock are written in any case
    try:
        records = mech.loop_over_loadcase(solver, steps, max_cut_back=getattr(solver, 'max_cut_back', 3), callback=callback)
    finally:
        if writer is not None and kw.get('rank', 0) == 0:
            writer.save()
    return records
