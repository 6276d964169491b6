"""Writes complete LFM case directories (dictionaries + polyMesh + initial fields) for synthetic meshes.

Dictionary keys are the ones CInputReader reads (reference: src/inputReader.cpp:13-64); default values
mirror examples/cylinder_vortex and examples/3D_Cylinder_Re3900/S.
"""
from __future__ import annotations

import os
import numpy as np

from . import meshgen

_HDR = "FoamFile\n{{\n    version     2.0;\n    format      ascii;\n    class       dictionary;\n    object      {obj};\n}}\n\n"

DEFAULTS = dict(
    commType=2, haloCommType=1, doublePrecision=True,
    haveAverage=False, haveForces=False, haveResiduals=True, saveForcesStep=1, printInfoFreq=1,
    tStartAverage=0.0,
    startTime=0.0, endTime=1.0, deltaT=1e-3, writeInterval=1000000, adjustTimeStep=False, maxCo=1.0,
    solver=0, dimension=2, rkOrder=5, minmodExists=False, writeFormat="ascii",
    pinf=1.0, Tinf=1.0, Uinf=(0.2, 0.0, 0.0), Ls=0.0, M=0.2,
    Cp=2.5, molWeight=11640.3, mu=0.0018667, Pr=0.75, simulationType="laminar",
)


def _b(x):
    return "true" if x else "false"


def write_dicts(case_dir, **kw):
    o = dict(DEFAULTS)
    o.update(kw)
    os.makedirs(os.path.join(case_dir, "system"), exist_ok=True)
    os.makedirs(os.path.join(case_dir, "constant"), exist_ok=True)
    with open(os.path.join(case_dir, "system", "controlDict"), "w") as f:
        f.write(_HDR.format(obj="controlDict"))
        f.write("lfm\n{\n")
        f.write(f"    commType        {o['commType']};\n    haloCommType    {o['haloCommType']};\n")
        f.write(f"    doublePrecision {_b(o['doublePrecision'])};\n    post\n    {{\n")
        f.write(f"        haveProbes      false;\n        haveSampling    false;\n")
        f.write(f"        haveAverage     {_b(o['haveAverage'])};\n        tStartAverage   {o['tStartAverage']!r};\n")
        f.write(f"        haveForces      {_b(o['haveForces'])};\n        haveResiduals   {_b(o['haveResiduals'])};\n")
        f.write(f"        saveForcesStep  {o['saveForcesStep']};\n        printInfoFreq   {o['printInfoFreq']};\n")
        f.write("        saveResiduals   false;\n        saveBlendFactor false;\n        saveRank        false;\n    }\n}\n\n")
        f.write(f"startFrom       startTime;\nstartTime       {o['startTime']!r};\nstopAt          endTime;\n")
        f.write(f"endTime         {o['endTime']!r};\ndeltaT          {o['deltaT']!r};\nwriteControl    timeStep;\n")
        f.write(f"writeInterval   {o['writeInterval']};\npurgeWrite      0;\nwriteFormat     {o['writeFormat']};\nwritePrecision  17;\n")
        f.write("timeFormat      general;\ntimePrecision   12;\n")
        f.write(f"adjustTimeStep  {'yes' if o['adjustTimeStep'] else 'no'};\nmaxCo           {o['maxCo']!r};\n")
    with open(os.path.join(case_dir, "system", "fvSchemes"), "w") as f:
        f.write(_HDR.format(obj="fvSchemes"))
        f.write("lfm\n{\n")
        f.write(f"    solver           {o['solver']};\n    dimension        {o['dimension']};\n    rkOrder          {o['rkOrder']};\n")
        f.write(f"    minmodExists     {_b(o['minmodExists'])};\n    constantTimeStep true;\n}}\n")
    with open(os.path.join(case_dir, "constant", "spongeDict"), "w") as f:
        f.write(_HDR.format(obj="spongeDict"))
        f.write(f"Ls      {o['Ls']!r};\nM       {o['M']!r};\npinf    {o['pinf']!r};\nTinf    {o['Tinf']!r};\n")
        f.write(f"Uinf_x  {o['Uinf'][0]!r};\nUinf_y  {o['Uinf'][1]!r};\nUinf_z  {o['Uinf'][2]!r};\n")
    with open(os.path.join(case_dir, "constant", "thermophysicalProperties"), "w") as f:
        f.write(_HDR.format(obj="thermophysicalProperties"))
        f.write("mixture\n{\n    specie\n    {\n        nMoles 1;\n")
        f.write(f"        molWeight {o['molWeight']!r};\n    }}\n    thermodynamics\n    {{\n        Cp {o['Cp']!r};\n        Hf 0;\n    }}\n")
        f.write(f"    transport\n    {{\n        mu {o['mu']!r};\n        Pr {o['Pr']!r};\n    }}\n}}\n")
    with open(os.path.join(case_dir, "constant", "turbulenceProperties"), "w") as f:
        f.write(_HDR.format(obj="turbulenceProperties"))
        f.write(f"simulationType  {o['simulationType']};\n")
    return o


def synthetic_fields(m, xc=None, amp=1.0, Uinf=(0.2, 0.0, 0.0)):
    """Deterministic smooth initial fields:
    p = 1 + 0.01 sin(2 pi x/Lx) cos(2 pi y/Ly), T = 1, U = (0.2(1+0.05 sin(2 pi y/Ly)), 0.01 sin(2 pi x/Lx), 0.01 sin(2 pi z/Lz))."""
    if xc is None:
        xc = meshgen.cell_centres_estimate(m)
    lo = m["points"].min(0)
    L = np.maximum(m["points"].max(0) - lo, 1e-30)
    s = 2.0 * np.pi * (xc - lo) / L
    p = 1.0 + amp * 0.01 * np.sin(s[:, 0]) * np.cos(s[:, 1])
    T = np.ones(len(xc))
    U = np.stack([Uinf[0] * (1.0 + amp * 0.05 * np.sin(s[:, 1])), Uinf[1] + amp * 0.01 * np.sin(s[:, 0]),
                  Uinf[2] + amp * 0.01 * np.sin(s[:, 2])], axis=1)
    return dict(p=p, T=T, U=U)


def distance_alpha(m, xc, names=("inlet", "outlet", "Inlet", "Outlet", "inflow", "outflow", "Inflow", "Outflow")):
    """Geometric stand-in for the eikonal `alpha` field: distance from the cell centre to the nearest
    face centre of the far-field patches (reference: eikonal/eikonal-OF_v2112/eikonal.cpp:30-41)."""
    from scipy.spatial import cKDTree
    pts = []
    for p in m["patches"]:
        if p["name"] in names and p["nFaces"] > 0:
            f = m["faces"][p["startFace"]:p["startFace"] + p["nFaces"]]
            valid = f >= 0
            pts.append((m["points"][np.maximum(f, 0)] * valid[..., None]).sum(1) / valid.sum(1)[:, None])
    if not pts:
        return np.full(len(xc), 1e30)
    d, _ = cKDTree(np.concatenate(pts)).query(xc)
    return d


def write_case(case_dir, m, fields=None, ranks=None, cell_rank=None, two_d=False, **dict_kw):
    """Writes dictionaries, the serial mesh + 0/ fields and (if cell_rank is given) processorN directories.
    Returns the option dict used."""
    xc = meshgen.cell_centres_estimate(m)
    if fields is None:
        fields = synthetic_fields(m, xc, Uinf=dict_kw.get("Uinf", DEFAULTS["Uinf"]))
    fields = dict(fields)
    if two_d:
        fields["U"] = fields["U"] * np.array([1.0, 1.0, 0.0])
    if "alpha" not in fields:
        fields["alpha"] = distance_alpha(m, xc)
    opts = write_dicts(case_dir, **dict_kw)
    t0 = _time_name(opts["startTime"])
    meshgen.write_polymesh(m, os.path.join(case_dir, "constant", "polyMesh"))
    os.makedirs(os.path.join(case_dir, t0), exist_ok=True)
    for name, v in fields.items():
        meshgen.write_field(os.path.join(case_dir, t0, name), name, m, v)
    if cell_rank is not None:
        parts = meshgen.decompose(m, cell_rank)
        for r, pm in enumerate(parts):
            pdir = os.path.join(case_dir, f"processor{r}")
            meshgen.write_polymesh(pm, os.path.join(pdir, "constant", "polyMesh"))
            os.makedirs(os.path.join(pdir, t0), exist_ok=True)
            for name, v in fields.items():
                meshgen.write_field(os.path.join(pdir, t0, name), name, pm, np.asarray(v)[pm["cellProcAddressing"]])
        opts["n_ranks"] = len(parts)
    return opts


def _time_name(t):
    s = "%.12g" % t
    return s


def set_end_time(case_dir, end_time):
    """Rewrites `endTime` in system/controlDict of an existing case (the same case is advanced for a different number of steps)."""
    import re
    p = os.path.join(case_dir, "system", "controlDict")
    s = open(p).read()
    s, n = re.subn(r"^endTime\s+\S+;", f"endTime         {end_time!r};", s, flags=re.M)
    if n != 1:
        raise RuntimeError(f"{p}: no endTime entry")
    open(p, "w").write(s)
