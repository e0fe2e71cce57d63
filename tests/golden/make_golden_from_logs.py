#!/usr/bin/env python
"""Extracts the golden vectors G1/G2 from the reference's recorded run logs into
small compressed fixtures. Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden_from_logs.py

Outputs (committed):
  tests/golden/g1_fortran_run.npz   every residual evaluation of /root/reference/output_damask_f.sh:
      F   (108,243)  'FR F1' lines   -- the F field handed to form_residual, Fortran memory order
      r   (108,243)  'FR r2' lines   -- the post-Gamma residual field returned
      P_av_MPa (108,3,3)             -- printed homogenised P (4 decimals, natural [i][j] as printed)
      inc (108,)                     -- 1-based increment each evaluation belongs to
      err_div, err_div_tol, err_BC, err_BC_tol (59,)  and report_inc (59,)  -- the 'reporting' blocks
      total_iter_at_convergence (10,)
  tests/golden/g2_constitutive_known_answer.npz  from /root/reference/output_damask_cpp.sh:290-303:
      F (243,), P (243,)  (Fortran memory order), P_av (9,), rms_raw, rms_scaled
      grain_map (27,)     -- recovered, see SURVEY.md 0.2 V5 (0-based index into material.yaml `material:`)
"""
import os
import re
import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def floats(line):
    return np.array([float(x) for x in line.split()])


def g1():
    F, r, Pav, inc_of_eval = [], [], [], []
    ediv, edtol, ebc, ebtol, rep_inc = [], [], [], [], []
    conv_iter = []
    inc = 0
    last_iter = None
    with open(os.path.join(REF, 'output_damask_f.sh'), errors='replace') as f:
        lines = f.readlines()
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r'\s*Time .*Increment (\d+)/(\d+)-', ln)
        if m:
            inc = int(m.group(1))
        if ln.startswith(' FR F1'):
            F.append(floats(ln[6:])); inc_of_eval.append(inc)
        elif ln.startswith(' FR r2'):
            r.append(floats(ln[6:]))
        elif 'Piola--Kirchhoff stress       / MPa' in ln and inc > 0:
            Pav.append(np.array([floats(lines[i + k]) for k in (1, 2, 3)]))
        elif '@ Iteration' in ln:
            m2 = re.search(r'≤(\d+)≤', ln)
            last_iter = int(m2.group(1))
        elif ln.startswith(' error divergence'):
            m3 = re.search(r'\(([-+0-9.E]+) / m, tol =\s*([-+0-9.E]+)\)', ln)
            ediv.append(float(m3.group(1))); edtol.append(float(m3.group(2))); rep_inc.append(inc)
        elif ln.startswith(' error stress BC'):
            m3 = re.search(r'\(([-+0-9.E]+) Pa,\s*tol =\s*([-+0-9.E]+)\)', ln)
            ebc.append(float(m3.group(1))); ebtol.append(float(m3.group(2)))
        elif re.match(r'\s*increment \d+ converged', ln):
            conv_iter.append(last_iter)
        i += 1
    F = np.array(F); r = np.array(r); Pav = np.array(Pav)
    assert F.shape == (108, 243) and r.shape == (108, 243), (F.shape, r.shape)
    assert Pav.shape[0] == 108, Pav.shape
    np.savez_compressed(os.path.join(HERE, 'g1_fortran_run.npz'), F=F, r=r, P_av_MPa=Pav,
                        inc=np.array(inc_of_eval), err_div=np.array(ediv), err_div_tol=np.array(edtol),
                        err_BC=np.array(ebc), err_BC_tol=np.array(ebtol), report_inc=np.array(rep_inc),
                        total_iter_at_convergence=np.array(conv_iter))
    print('G1:', F.shape, r.shape, Pav.shape, len(ediv), conv_iter)


def g2():
    with open(os.path.join(REF, 'output_damask_cpp.sh'), errors='replace') as f:
        lines = f.readlines()
    out = {}
    for k in range(286, 310):
        ln = lines[k]
        if ln.startswith(' FF ') and 'F' not in out:
            out['F'] = floats(ln[4:])
        elif ln.startswith(' PP ') and 'P' not in out:
            out['P'] = floats(ln[4:])
        elif ln.startswith('rms1') and 'rms_raw' not in out:
            out['rms_raw'] = float(ln[4:])
        elif ln.startswith('rms2') and 'rms_scaled' not in out:
            out['rms_scaled'] = float(ln[4:])
        elif ln.startswith(' P_av2') and 'P_av' not in out:
            out['P_av'] = floats(ln[7:])
    assert out['F'].shape == (243,) and out['P'].shape == (243,)
    out['grain_map'] = np.array([19, 9, 17, 16, 6, 7, 18, 4, 3, 11, 13, 14, 16, 5, 10, 11, 5, 3, 0, 13, 1, 15, 8,
                                 10, 12, 2, 2])
    np.savez_compressed(os.path.join(HERE, 'g2_constitutive_known_answer.npz'), **out)
    print('G2:', {k: np.shape(v) for k, v in out.items()})




def inputs():
    """The run's input deck is data, not source: keep a copy of the 20-orientation material file and the
    tensionX load case next to the vectors they produced (examples/grid/{material,tensionX}.yaml)."""
    import shutil
    shutil.copy(os.path.join(REF, 'examples/grid/material.yaml'), os.path.join(HERE, 'material_20grains_Al.yaml'))
    shutil.copy(os.path.join(REF, 'examples/grid/tensionX.yaml'), os.path.join(HERE, 'load_tensionX.yaml'))
    shutil.copy(os.path.join(REF, 'examples/grid/20grains.seeds'), os.path.join(HERE, '20grains.seeds'))


if __name__ == '__main__':
    g1()
    g2()
    inputs()
