"""Reads an existing ASCII OpenFOAM case (the reference's examples/) into the mesh dict the generators use, and
decomposes it into processorN directories (decomposePar is not available in the image): the pre-processing side of
BASELINE configs 2-4 (reference: README.md:158-163 `decomposePar` + `mpirun -np N lfm_solve -p`)."""
from __future__ import annotations

import os
import re

import numpy as np

from . import meshgen


def _body(path):
    s = open(path).read()
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    i = s.index("}", s.index("FoamFile")) + 1          # behind the header dictionary
    m = re.search(r"(\d+)\s*\(", s[i:])
    n = int(m.group(1))
    a = i + m.end()
    b = s.rindex(")")
    return n, s[a:b]


def _numbers(text, dtype):
    return np.array(text.replace("(", " ").replace(")", " ").split(), dtype=dtype)


def read_polymesh(mesh_dir):
    n, t = _body(os.path.join(mesh_dir, "points"))
    points = _numbers(t, np.float64).reshape(n, 3)
    n, t = _body(os.path.join(mesh_dir, "owner"))
    owner = _numbers(t, np.int64)
    assert len(owner) == n
    n, t = _body(os.path.join(mesh_dir, "neighbour"))
    neighbour = _numbers(t, np.int64)
    nf, t = _body(os.path.join(mesh_dir, "faces"))
    tok = _numbers(t, np.int64)
    faces = np.full((nf, 4), -1, dtype=np.int64)
    if len(tok) == 5 * nf and (tok[::5] == 4).all():
        faces[:] = tok.reshape(nf, 5)[:, 1:]
    else:
        pos = 0
        for f in range(nf):
            k = int(tok[pos])
            assert k in (3, 4), "only triangles and quadrilaterals"
            faces[f, :k] = tok[pos + 1:pos + 1 + k]
            pos += 1 + k
    s = open(os.path.join(mesh_dir, "boundary")).read()
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = s[s.index("}", s.index("FoamFile")) + 1:]
    patches = []
    for name, body in re.findall(r"(\w+)\s*\{([^{}]*)\}", s):
        d = dict(re.findall(r"(\w+)\s+([^;]+);", body))
        p = dict(name=name, type=d["type"].strip(), nFaces=int(d["nFaces"]), startFace=int(d["startFace"]))
        if "neighbourPatch" in d:
            p["neighbourPatch"] = d["neighbourPatch"].strip()
        patches.append(p)
    return dict(points=points, faces=faces.astype(np.int32), owner=owner.astype(np.int32), neighbour=neighbour.astype(np.int32),
                patches=patches, nCells=int(owner.max()) + 1)


def read_internal_field(path, n_cells):
    """internalField of a vol field: uniform or nonuniform, scalar or vector -> [n_cells] or [n_cells, 3]."""
    s = open(path).read()
    i = s.index("internalField")
    head = s[i:s.index("\n", i)] if "nonuniform" not in s[i:i + 200] else None
    if head is not None and "uniform" in head:
        v = head.split("uniform", 1)[1].strip().rstrip(";").strip()
        if v.startswith("("):
            return np.tile(np.array(v.strip("()").split(), dtype=np.float64), (n_cells, 1))
        return np.full(n_cells, float(v))
    j = s.index("boundaryField")
    vec = "List<vector>" in s[i:i + 200]
    body = s[s.index("(", i) + 1:s.rindex(")", i, j)]
    a = _numbers(body, np.float64)
    return a.reshape(-1, 3) if vec else a


def decompose_case(case_dir, n_ranks, time_name="0", fields=("p", "T", "U", "alpha")):
    """Strips of (nearly) equal cell count along x (cell centre estimate), written as processorN/{constant/polyMesh,<time>}."""
    m = read_polymesh(os.path.join(case_dir, "constant", "polyMesh"))
    xc = meshgen.cell_centres_estimate(m)
    order = np.argsort(xc[:, 0], kind="stable")
    cell_rank = np.empty(m["nCells"], dtype=np.int32)
    cell_rank[order] = (np.arange(m["nCells"]) * n_ranks // m["nCells"]).astype(np.int32)
    vals = {}
    for name in fields:
        p = os.path.join(case_dir, time_name, name)
        if os.path.exists(p):
            vals[name] = read_internal_field(p, m["nCells"])
    parts = meshgen.decompose(m, cell_rank)
    for r, pm in enumerate(parts):
        pdir = os.path.join(case_dir, f"processor{r}")
        meshgen.write_polymesh(pm, os.path.join(pdir, "constant", "polyMesh"))
        os.makedirs(os.path.join(pdir, time_name), exist_ok=True)
        for name, v in vals.items():
            meshgen.write_field(os.path.join(pdir, time_name, name), name, pm, np.asarray(v)[pm["cellProcAddressing"]])
    return m, cell_rank, parts
