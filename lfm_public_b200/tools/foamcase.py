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


# ---------------------------------------------------------------------------------------------------------------
# OpenFOAM binary stream format (`writeFormat binary`, what decomposePar writes for the reference's 3D cases): tokens stay
# text, the contents of contiguous lists are raw little-endian elements between `(` and `)`.
# ---------------------------------------------------------------------------------------------------------------
def _bin_header(cls, obj, label_bytes):
    return ("FoamFile\n{\n    version     2.0;\n    format      binary;\n"
            f"    arch        \"LSB;label={8 * label_bytes};scalar=64\";\n    class       {cls};\n    object      {obj};\n}}\n"
            "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n").encode()


def _bin_list(arr):
    return str(len(arr)).encode() + b"\n(" + np.ascontiguousarray(arr).tobytes() + b")\n"


def to_binary(case_dir, label_bytes=4, time_name="0", fields=("p", "T", "U", "alpha")):
    """Rewrites the polyMesh lists and the time-directory fields of an ASCII case (serial and processorN parts) in
    OpenFOAM's binary format, in place (faces become a faceCompactList, as OpenFOAM writes them in binary)."""
    lab = np.int32 if label_bytes == 4 else np.int64
    roots = [case_dir] + sorted(os.path.join(case_dir, d) for d in os.listdir(case_dir) if d.startswith("processor"))
    for root in roots:
        md = os.path.join(root, "constant", "polyMesh")
        if not os.path.isdir(md):
            continue
        n, t = _body(os.path.join(md, "points"))
        pts = _numbers(t, np.float64).reshape(n, 3)
        with open(os.path.join(md, "points"), "wb") as f:
            f.write(_bin_header("vectorField", "points", label_bytes) + _bin_list(pts))
        nf, t = _body(os.path.join(md, "faces"))
        tok = _numbers(t, np.int64)
        offs, labels, pos = [0], [], 0
        for _ in range(nf):
            k = int(tok[pos])
            labels.extend(tok[pos + 1:pos + 1 + k].tolist())
            offs.append(len(labels))
            pos += 1 + k
        with open(os.path.join(md, "faces"), "wb") as f:
            f.write(_bin_header("faceCompactList", "faces", label_bytes) + _bin_list(np.array(offs, dtype=lab)) + b"\n" + _bin_list(np.array(labels, dtype=lab)))
        n_cells = 0
        for name in ("owner", "neighbour", "faceProcAddressing", "cellProcAddressing", "pointProcAddressing", "boundaryProcAddressing", "cellSubmesh"):
            p = os.path.join(md, name)
            if not os.path.exists(p):
                continue
            n, t = _body(p)
            a = _numbers(t, np.int64)
            assert len(a) == n
            if name == "owner":
                n_cells = int(a.max()) + 1
            with open(p, "wb") as f:
                f.write(_bin_header("labelList", name, label_bytes) + _bin_list(a.astype(lab)))
        # polyBoundaryMesh stays a text dictionary list under a binary header, as OpenFOAM writes it
        p = os.path.join(md, "boundary")
        s = open(p).read().replace("format      ascii;", "format      binary;")
        open(p, "w").write(s)
        for name in fields:
            p = os.path.join(root, time_name, name)
            if not os.path.exists(p):
                continue
            s = open(p).read()
            i = s.index("internalField")
            if "nonuniform" not in s[i:i + 200]:
                open(p, "w").write(s.replace("format      ascii;", "format      binary;"))
                continue
            v = read_internal_field(p, n_cells)
            vec = v.ndim == 2
            tail = s[s.index("boundaryField"):]
            with open(p, "wb") as f:
                f.write(_bin_header("volVectorField" if vec else "volScalarField", name, label_bytes))
                f.write(b"dimensions      [0 0 0 0 0 0 0];\n\n")
                f.write(f"internalField   nonuniform List<{'vector' if vec else 'scalar'}> \n".encode())
                f.write(str(len(v)).encode() + b"\n(" + np.ascontiguousarray(v, dtype=np.float64).tobytes() + b");\n\n")
                f.write(tail.encode())
