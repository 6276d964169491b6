#!/usr/bin/env python
"""Standalone cell renumbering of an OpenFOAM/LFM case: the pre-processing step the reference does with OpenFOAM's
`renumberMesh` (examples/*/constant/renumberMeshDict: CuthillMcKee, or the hpathRenumber plugin in 2D), without OpenFOAM.

    python -m lfm_public_b200.tools.renumber <case_in> <case_out> [--method morton|hilbert|rcm|hpath]

Writes <case_out> with the permuted polyMesh (faces re-sorted into OpenFOAM's upper-triangular order) and the permuted
fields of the start time directory; dictionaries are copied.  Methods:
  rcm      reverse Cuthill-McKee (scipy), what the examples' renumberMeshDict asks for: minimises bandwidth, but a
           wavefront numbering keeps a cell's neighbours ~one front width away -- poor for shared-memory tiles;
  morton   cells along a Z-order curve through their centres: consecutive ids form compact patches, which is what the
           tile kernels (and hpath's "boundary cells first, then a path through the interior") want;
  hilbert  2D Hilbert curve (x, y), slightly more compact than Z-order for extruded 2D meshes;
  hpath    the reference's own hpathRenumber plugin restated (lfm_public_b200/host/hpath.cpp): boundary submesh first,
           then a Hamiltonian-like path through the interior (2D meshes).
"""
from __future__ import annotations

import argparse
import os
import shutil

import numpy as np

from . import foamcase, meshgen


def _hilbert2(x, y, bits=16):
    """Hilbert index of integer (x, y) < 2**bits (vectorised)."""
    x = x.astype(np.int64).copy()
    y = y.astype(np.int64).copy()
    d = np.zeros_like(x)
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        swap = ry == 0
        flip = swap & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def order_cells(m, method):
    """new_of_old permutation."""
    if method == "rcm":
        return meshgen.rcm_order(m)
    xc = meshgen.cell_centres_estimate(m)
    span = np.maximum(np.ptp(xc, axis=0), 1e-300)
    if method == "hilbert":
        q = ((xc[:, :2] - xc[:, :2].min(0)) / span[:2] * (2 ** 16 - 1)).astype(np.int64)
        key = _hilbert2(q[:, 0], q[:, 1], 16)
    else:
        q = ((xc - xc.min(0)) / span * (2 ** 20 - 1)).astype(np.int64)
        key = meshgen._morton3(q[:, 0], q[:, 1], q[:, 2])
    order = np.argsort(key, kind="stable")
    new_of_old = np.empty(m["nCells"], dtype=np.int64)
    new_of_old[order] = np.arange(m["nCells"])
    return new_of_old


def locality(m, new_of_old=None):
    nif = len(m["neighbour"])
    o, n = m["owner"][:nif].astype(np.int64), m["neighbour"].astype(np.int64)
    if new_of_old is not None:
        o, n = new_of_old[o], new_of_old[n]
    d = np.abs(o - n)
    return dict(mean=float(d.mean()), median=float(np.median(d)), p90=float(np.percentile(d, 90)), max=int(d.max()))


def hpath_order(case_in):
    """new_of_old of the reference's hpath numbering, computed by liblfmhost.so on the case's polyMesh (+ the walk statistics)."""
    from .. import host_api
    case = host_api.Case.open(case_in)
    order, stats = case.hpath_order()          # order[new] = old
    case.close()
    new_of_old = np.empty(len(order), dtype=np.int64)
    new_of_old[order] = np.arange(len(order))
    return new_of_old, stats


def renumber_case(case_in, case_out, method="morton", time_name="0", fields=("p", "T", "U", "alpha")):
    m = foamcase.read_polymesh(os.path.join(case_in, "constant", "polyMesh"))
    new_of_old = hpath_order(case_in)[0] if method == "hpath" else order_cells(m, method)
    out = meshgen.renumber_cells(m, new_of_old)
    if os.path.exists(case_out):
        shutil.rmtree(case_out)
    os.makedirs(case_out)
    for sub in ("system", "constant"):
        shutil.copytree(os.path.join(case_in, sub), os.path.join(case_out, sub), ignore=shutil.ignore_patterns("polyMesh"))
    meshgen.write_polymesh(out, os.path.join(case_out, "constant", "polyMesh"))
    os.makedirs(os.path.join(case_out, time_name))
    for name in fields:
        p = os.path.join(case_in, time_name, name)
        if not os.path.exists(p):
            continue
        v = foamcase.read_internal_field(p, m["nCells"])
        w = np.empty_like(v)
        w[new_of_old] = v
        meshgen.write_field(os.path.join(case_out, time_name, name), name, out, w)
    return m, out, new_of_old


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case_in")
    ap.add_argument("case_out")
    ap.add_argument("--method", default="morton", choices=["morton", "hilbert", "rcm", "hpath"])
    args = ap.parse_args()
    m, out, p = renumber_case(args.case_in, args.case_out, args.method)
    print("before:", locality(m))
    print("after: ", locality(out))


if __name__ == "__main__":
    main()
