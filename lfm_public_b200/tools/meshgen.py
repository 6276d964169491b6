"""Synthetic polyMesh generators + block decomposer (test/bench tooling, numpy-vectorised).

The reference's inputs are OpenFOAM polyMesh cases (reference: README.md:150-168,
examples/*/constant/polyMesh); blockMesh/decomposePar are not available here, so these generators
write the same on-disk format for structured hex boxes (3D), one-cell-thick
quad boxes (2D, front/back `empty`), O-grid cylinders and triangulated (prism) 2D meshes, and split
them into processorN meshes with `processor` patches + faceProcAddressing the way decomposePar does.

A mesh is a dict:
  points [np,3] f8 | faces [nf,4] i4 (-1 padded for triangles) | owner [nf] | neighbour [nif]
  patches: list of dict(name,type,nFaces,startFace[,neighbourPatch,myProcNo,neighbProcNo]) | nCells
  optional faceProcAddressing / cellProcAddressing / pointProcAddressing / boundaryProcAddressing
"""
from __future__ import annotations

import os
import numpy as np

# ------------------------------------------------------------------------------------------------
# generic assembly: arbitrary faces -> OpenFOAM canonical order
# ------------------------------------------------------------------------------------------------

def _flip(faces: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Reverse the point order of the selected faces (keeps point 0, handles -1 padded triangles)."""
    out = faces.copy()
    if not mask.any():
        return out
    sel = faces[mask]
    tri = sel[:, 3] < 0
    quad = ~tri
    rev = sel.copy()
    rev[quad] = sel[quad][:, [0, 3, 2, 1]]
    rev[tri] = sel[tri][:, [0, 2, 1, 3]]
    out[mask] = rev
    return out


def assemble(points, faces, cell_a, cell_b, patch_id, patch_defs, n_cells, cyclic_keys=None):
    """faces: [nf,4]; normal of each face points from cell_a to cell_b (cell_b == -1: boundary face of
    cell_a with outward normal, patch_id gives the patch).  Returns a mesh dict in canonical order:
    internal faces sorted by (owner, neighbour) (upper-triangular), then patches in patch_defs order.
    cyclic_keys: optional per-face sort key inside a patch so that cyclic twins share the same offset."""
    faces = np.asarray(faces, dtype=np.int32)
    cell_a = np.asarray(cell_a, dtype=np.int64)
    cell_b = np.asarray(cell_b, dtype=np.int64)
    internal = cell_b >= 0
    own = np.where(internal, np.minimum(cell_a, cell_b), cell_a)
    nei = np.where(internal, np.maximum(cell_a, cell_b), -1)
    faces = _flip(faces, internal & (cell_a > cell_b))
    idx_int = np.nonzero(internal)[0]
    order_int = idx_int[np.lexsort((nei[idx_int], own[idx_int]))]
    idx_bnd = np.nonzero(~internal)[0]
    if cyclic_keys is None:
        key = idx_bnd
    else:
        key = np.asarray(cyclic_keys)[idx_bnd]
    order_bnd = idx_bnd[np.lexsort((key, patch_id[idx_bnd]))]
    order = np.concatenate([order_int, order_bnd])
    pid_sorted = patch_id[order_bnd]
    patches = []
    start = len(order_int)
    for i, pd in enumerate(patch_defs):
        n = int(np.count_nonzero(pid_sorted == i))
        d = dict(pd)
        d["nFaces"] = n
        d["startFace"] = start
        start += n
        patches.append(d)
    return dict(points=np.asarray(points, dtype=np.float64), faces=faces[order], owner=own[order].astype(np.int32),
                neighbour=nei[order_int].astype(np.int32), patches=patches, nCells=int(n_cells))


# ------------------------------------------------------------------------------------------------
# structured (i,j,k) topology with optional periodic j (O-grid) -> faces
# ------------------------------------------------------------------------------------------------

def _structured(ni, nj, nk, periodic_j=False):
    """Logical hex block.  Returns (faces, cell_a, cell_b, side) where side is 0..5 for boundary faces
    (imin,imax,jmin,jmax,kmin,kmax) and -1 for internal ones.  Point index p(i,j,k) with j wrapped if
    periodic."""
    npj = nj if periodic_j else nj + 1

    def P(i, j, k):
        return (i + (ni + 1) * ((j % npj) + npj * k)).astype(np.int32)

    def C(i, j, k):
        return i + ni * ((j % nj) + nj * k)

    out_f, out_a, out_b, out_s = [], [], [], []
    # i-faces (normal +i)
    i, j, k = np.meshgrid(np.arange(ni + 1), np.arange(nj), np.arange(nk), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    f = np.stack([P(i, j, k), P(i, j + 1, k), P(i, j + 1, k + 1), P(i, j, k + 1)], axis=1)
    a = np.where(i > 0, C(np.maximum(i - 1, 0), j, k), -1)
    b = np.where(i < ni, C(np.minimum(i, ni - 1), j, k), -1)
    out_f.append(f); out_a.append(a); out_b.append(b); out_s.append(np.where(i == 0, 0, np.where(i == ni, 1, -1)))
    # j-faces (normal +j)
    i, j, k = np.meshgrid(np.arange(ni), np.arange(nj if periodic_j else nj + 1), np.arange(nk), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    f = np.stack([P(i, j, k), P(i, j, k + 1), P(i + 1, j, k + 1), P(i + 1, j, k)], axis=1)
    if periodic_j:
        a = C(i, j - 1, k)
        b = C(i, j, k)
        s = np.full(len(i), -1)
    else:
        a = np.where(j > 0, C(i, np.maximum(j - 1, 0), k), -1)
        b = np.where(j < nj, C(i, np.minimum(j, nj - 1), k), -1)
        s = np.where(j == 0, 2, np.where(j == nj, 3, -1))
    out_f.append(f); out_a.append(a); out_b.append(b); out_s.append(s)
    # k-faces (normal +k)
    i, j, k = np.meshgrid(np.arange(ni), np.arange(nj), np.arange(nk + 1), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    f = np.stack([P(i, j, k), P(i + 1, j, k), P(i + 1, j + 1, k), P(i, j + 1, k)], axis=1)
    a = np.where(k > 0, C(i, j, np.maximum(k - 1, 0)), -1)
    b = np.where(k < nk, C(i, j, np.minimum(k, nk - 1)), -1)
    out_f.append(f); out_a.append(a); out_b.append(b); out_s.append(np.where(k == 0, 4, np.where(k == nk, 5, -1)))
    faces = np.concatenate(out_f)
    a = np.concatenate(out_a)
    b = np.concatenate(out_b)
    side = np.concatenate(out_s)
    # boundary faces: make cell_a the existing cell with outward normal
    low = (a < 0)
    faces = _flip(faces, low)
    a2 = np.where(low, b, a)
    b2 = np.where(low | (b < 0), -1, b)
    return faces, a2, b2, side, npj


def _finish(points, faces, a, b, side, side_patch, patch_defs, n_cells, cyclic_pairs=()):
    """side_patch[s] = patch index for logical side s."""
    patch_id = np.full(len(side), -1, dtype=np.int64)
    for s in range(6):
        patch_id[side == s] = side_patch[s]
    keep = (b >= 0) | (patch_id >= 0)
    assert keep.all(), "boundary side without a patch"
    return assemble(points, faces, a, b, patch_id, patch_defs, n_cells)


def hex_box(nx, ny, nz, lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), two_d=False, z_cyclic=False,
            grading=None):
    """3D box nx*ny*nz hexes.  Patches: inlet (x-), outlet (x+), walls (y-, y+ : type wall) and for z
    either `frontAndBack` empty (two_d, requires nz == 1), a cyclic pair periodic_m/periodic_p, or walls.
    grading: optional callable (axis, u in [0,1]) -> [0,1] to stretch the grid."""
    if two_d:
        assert nz == 1
    faces, a, b, side, _ = _structured(nx, ny, nz)
    xs = np.linspace(0.0, 1.0, nx + 1)
    ys = np.linspace(0.0, 1.0, ny + 1)
    zs = np.linspace(0.0, 1.0, nz + 1)
    if grading is not None:
        xs, ys, zs = grading(0, xs), grading(1, ys), grading(2, zs)
    X, Y, Z = np.meshgrid(origin[0] + lengths[0] * xs, origin[1] + lengths[1] * ys, origin[2] + lengths[2] * zs, indexing="ij")
    # point index = i + (nx+1)*(j + (ny+1)*k)  -> order k slowest
    points = np.stack([X.transpose(2, 1, 0).ravel(), Y.transpose(2, 1, 0).ravel(), Z.transpose(2, 1, 0).ravel()], axis=1)
    patch_defs = [dict(name="inlet", type="patch"), dict(name="outlet", type="patch"), dict(name="walls", type="wall")]
    side_patch = [0, 1, 2, 2, -1, -1]
    if two_d:
        patch_defs.append(dict(name="frontAndBack", type="empty"))
        side_patch[4] = side_patch[5] = 3
    elif z_cyclic:
        patch_defs.append(dict(name="periodic_m", type="cyclic", neighbourPatch="periodic_p"))
        patch_defs.append(dict(name="periodic_p", type="cyclic", neighbourPatch="periodic_m"))
        side_patch[4], side_patch[5] = 3, 4
    else:
        side_patch[4] = side_patch[5] = 2
    m = _finish(points, faces, a, b, side, side_patch, patch_defs, nx * ny * nz)
    m["logical"] = (nx, ny, nz)
    return m


def ogrid_cylinder(nr, nth, nz, r_in=0.5, r_out=20.0, span=np.pi, two_d=False, stretch=1.0):
    """O-grid around a wall cylinder (patch `Cylinder`, type wall); outer boundary split into `Inlet`
    (x<0) and `Outlet` (x>=0) patches (names from examples/3D_Cylinder_Re3900/S/0/U); z: empty (2D)
    or cyclic periodic_m/periodic_p (examples/3D_Cylinder_Re3900/S/system/createPatchDict)."""
    assert nth % 2 == 0
    faces, a, b, side, npj = _structured(nr, nth, nz, periodic_j=True)
    u = np.linspace(0.0, 1.0, nr + 1)
    if stretch != 1.0:
        u = (np.power(stretch, u) - 1.0) / (stretch - 1.0)
    rr = r_in + (r_out - r_in) * u
    # theta offset so that the inlet/outlet split (x = 0) falls on cell faces
    th = np.pi / 2 + 2.0 * np.pi * np.arange(nth) / nth
    zz = span * np.linspace(0.0, 1.0, nz + 1) if not two_d else np.array([0.0, 1.0])
    R, T, Z = np.meshgrid(rr, th, zz, indexing="ij")
    X = R * np.cos(T)
    Y = R * np.sin(T)
    points = np.stack([X.transpose(2, 1, 0).ravel(), Y.transpose(2, 1, 0).ravel(), Z.transpose(2, 1, 0).ravel()], axis=1)
    patch_defs = [dict(name="Cylinder", type="wall"), dict(name="Inlet", type="patch"), dict(name="Outlet", type="patch")]
    if two_d:
        patch_defs.append(dict(name="frontAndBack", type="empty"))
    else:
        patch_defs.append(dict(name="periodic_m", type="cyclic", neighbourPatch="periodic_p"))
        patch_defs.append(dict(name="periodic_p", type="cyclic", neighbourPatch="periodic_m"))
    patch_id = np.full(len(side), -1, dtype=np.int64)
    patch_id[side == 0] = 0
    outer = side == 1
    # face centre x of the outer faces decides inlet / outlet
    fc_x = points[faces[:, :4].clip(0), 0].mean(axis=1)
    patch_id[outer & (fc_x < 0)] = 1
    patch_id[outer & (fc_x >= 0)] = 2
    patch_id[side == 4] = 3
    patch_id[side == 5] = 3 if two_d else 4
    m = assemble(points, faces, a, b, patch_id, patch_defs, nr * nth * nz)
    m["logical"] = (nr, nth, nz)
    return m


def tri_prism_box(nx, ny, lengths=(1.0, 1.0), shuffle_seed=None):
    """2D unstructured-like mesh: every quad of an nx*ny box split into two triangle prisms
    (3 valid faces per cell + 2 `empty`), optionally with randomly permuted cell numbering (what a
    mesher produces before renumberMesh)."""
    npx, npy = nx + 1, ny + 1
    xs = np.linspace(0.0, lengths[0], npx)
    ys = np.linspace(0.0, lengths[1], npy)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    # jitter interior points a little so the triangles are not all congruent
    rng = np.random.default_rng(1234)
    jit = np.zeros((npx, npy, 2))
    jit[1:-1, 1:-1] = 0.15 * rng.uniform(-1, 1, size=(npx - 2, npy - 2, 2)) * np.array([lengths[0] / nx, lengths[1] / ny])
    X = X + jit[..., 0]
    Y = Y + jit[..., 1]
    n2 = npx * npy
    p2 = lambda i, j: i + npx * j
    points = np.concatenate([np.stack([X.T.ravel(), Y.T.ravel(), np.zeros(n2)], 1),
                             np.stack([X.T.ravel(), Y.T.ravel(), np.ones(n2)], 1)])
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i, j = i.ravel(), j.ravel()
    q = i + nx * j
    cl, cu = 2 * q, 2 * q + 1        # lower-right triangle (p00,p10,p11), upper-left (p00,p11,p01)
    p00, p10, p11, p01 = p2(i, j), p2(i + 1, j), p2(i + 1, j + 1), p2(i, j + 1)
    F, A, B, PID = [], [], [], []

    def side_quad(pa, pb):
        # edge pa->pb (counter-clockwise for the cell on its left); normal points to the right of the edge (outward)
        return np.stack([pa, pb, pb + n2, pa + n2], axis=1)

    def add(f, a, b, pid):
        F.append(f); A.append(a); B.append(b); PID.append(pid)

    # diagonal p00->p11: left cell = cu, right = cl ; normal of side_quad(p00,p11) points right: from cu to cl
    add(side_quad(p00, p11), cu, cl, np.full(len(q), -1))
    # bottom edge p00->p10 belongs to cl (ccw); neighbour below is cu of (i, j-1)
    below = np.where(j > 0, 2 * (i + nx * (j - 1)) + 1, -1)
    add(side_quad(p00, p10), cl, below, np.where(j > 0, -1, 2))
    # right edge p10->p11 belongs to cl; neighbour right is cu of (i+1, j)
    right = np.where(i < nx - 1, 2 * ((i + 1) + nx * j) + 1, -1)
    add(side_quad(p10, p11), cl, right, np.where(i < nx - 1, -1, 1))
    # top edge p11->p01 belongs to cu; boundary only (interior handled as "below" of the upper quad)
    top = j == ny - 1
    add(side_quad(p11[top], p01[top]), cu[top], np.full(top.sum(), -1), np.full(top.sum(), 2))
    # left edge p01->p00 belongs to cu; boundary only
    left = i == 0
    add(side_quad(p01[left], p00[left]), cu[left], np.full(left.sum(), -1), np.full(left.sum(), 0))
    # front (z=0, outward -z) and back (z=1, outward +z) triangles: empty
    m1 = np.full(len(q), -1, dtype=np.int64)
    add(np.stack([p00, p11, p10, m1], 1), cl, m1, np.full(len(q), 3))
    add(np.stack([p00, p01, p11, m1], 1), cu, m1, np.full(len(q), 3))
    add(np.stack([p00 + n2, p10 + n2, p11 + n2, m1], 1), cl, m1, np.full(len(q), 3))
    add(np.stack([p00 + n2, p11 + n2, p01 + n2, m1], 1), cu, m1, np.full(len(q), 3))
    faces = np.concatenate(F).astype(np.int32)
    a = np.concatenate(A).astype(np.int64)
    b = np.concatenate(B).astype(np.int64)
    pid = np.concatenate(PID).astype(np.int64)
    n_cells = 2 * nx * ny
    if shuffle_seed is not None:
        perm = np.random.default_rng(shuffle_seed).permutation(n_cells)
        a = perm[a]
        b = np.where(b >= 0, perm[np.maximum(b, 0)], -1)
    patch_defs = [dict(name="inlet", type="patch"), dict(name="outlet", type="patch"), dict(name="walls", type="wall"),
                  dict(name="frontAndBack", type="empty")]
    return assemble(points, faces, a, b, pid, patch_defs, n_cells)


# ------------------------------------------------------------------------------------------------
# renumbering (RCM stop-gap for hpathRenumber; reference: examples/*/constant/renumberMeshDict)
# ------------------------------------------------------------------------------------------------

def _morton3(x, y, z):
    """Z-order key of non-negative integer triples (up to 21 bits each)."""
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return spread(x) | (spread(y) << np.uint64(1)) | (spread(z) << np.uint64(2))


def blocked_order(ni, nj, nk, tile, offset=(0, 0, 0), brick_order="lex"):
    """Cell permutation new_of_old for a logical (ni,nj,nk) block: cells are numbered tile by tile (tiles of
    tile=(ti,tj,tk) cells, i fastest inside a tile and across tiles), the structured analogue of the reference's
    renumberMesh pre-processing (hpathRenumber / CuthillMcKee): consecutive cell ids form compact 3D bricks, which
    is what the shared-memory tile kernels want.  offset shifts the tile lattice (offset=(1,1,1) aligns the bricks
    with the interior submesh, whose first cell is (1,1,1)).  brick_order: "lex" = bricks in lexicographic order,
    "morton" = bricks along a Z-order curve, which keeps the bricks across EVERY face of a brick close in the numbering
    (with "lex" the z-neighbour brick is a whole xy-slab away, beyond the reach of the L2 on large blocks)."""
    ti, tj, tk = tile
    c = np.arange(ni * nj * nk, dtype=np.int64)
    i, j, k = c % ni, (c // ni) % nj, c // (ni * nj)
    io, jo, ko = i + (ti - offset[0]) % ti, j + (tj - offset[1]) % tj, k + (tk - offset[2]) % tk
    if brick_order == "morton":
        order = np.lexsort((io % ti, jo % tj, ko % tk, _morton3(io // ti, jo // tj, ko // tk)))
    else:
        order = np.lexsort((io % ti, jo % tj, ko % tk, io // ti, jo // tj, ko // tk))   # last key is primary
    new_of_old = np.empty(len(c), dtype=np.int64)
    new_of_old[order] = c
    return new_of_old


def hex_block(n, blocks=(1, 1, 1), rank=0, cell_size=None, z_cyclic=None, tile=None, brick_order="lex"):
    """The processor mesh of ONE rank of a block-decomposed hex box, generated without ever building the global
    mesh (weak-scaling runs: 16.8 M cells per rank).  n = (nx,ny,nz) cells of this block, blocks = (bx,by,bz),
    rank = bi + bx*(bj + by*bk) as in block_assignment().  Physically identical to
    decompose(hex_box(n*blocks), block_assignment)[rank] (same points, same cell order, same patches; the
    processor-face ids are a different but consistent global numbering -- they only serve to match the two sides).
    z is a cyclic pair when bz == 1 (default), walls otherwise; z_cyclic=True with bz > 1 keeps the span periodic: the bottom
    plane of the lowest blocks and the top plane of the highest become `processorCyclic` patches (as in decompose()).
    tile=(ti,tj,tk): number the cells brick by brick (blocked_order, aligned with the interior submesh); the
    permutation is returned as m["new_of_old"] (fields given in lexicographic order must be permuted with it)."""
    nx, ny, nz = n
    bx, by, bz = blocks
    bi, bj, bk = rank % bx, (rank // bx) % by, rank // (bx * by)
    GX, GY, GZ = nx * bx, ny * by, nz * bz
    if cell_size is None:
        cell_size = (1.0 / GX, 1.0 / GX, 1.0 / GX)
    if z_cyclic is None:
        z_cyclic = bz == 1
    L = (cell_size[0] * GX, cell_size[1] * GY, cell_size[2] * GZ)
    faces, a, b, side, _ = _structured(nx, ny, nz)
    xs = (L[0] * np.linspace(0.0, 1.0, GX + 1))[bi * nx:(bi + 1) * nx + 1]
    ys = (L[1] * np.linspace(0.0, 1.0, GY + 1))[bj * ny:(bj + 1) * ny + 1]
    zs = (L[2] * np.linspace(0.0, 1.0, GZ + 1))[bk * nz:(bk + 1) * nz + 1]
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    points = np.stack([X.transpose(2, 1, 0).ravel(), Y.transpose(2, 1, 0).ravel(), Z.transpose(2, 1, 0).ravel()], axis=1)
    patch_defs = [dict(name="inlet", type="patch"), dict(name="outlet", type="patch"), dict(name="walls", type="wall")]
    phys = [0, 1, 2, 2, 2, 2]
    if z_cyclic:
        patch_defs.append(dict(name="periodic_m", type="cyclic", neighbourPatch="periodic_p"))
        patch_defs.append(dict(name="periodic_p", type="cyclic", neighbourPatch="periodic_m"))
        phys[4], phys[5] = 3, 4
    # neighbour rank across each logical side (-1: domain boundary)
    def rk(i, j, k):
        return i + bx * (j + by * k)
    nbr = [rk(bi - 1, bj, bk) if bi > 0 else -1, rk(bi + 1, bj, bk) if bi < bx - 1 else -1,
           rk(bi, bj - 1, bk) if bj > 0 else -1, rk(bi, bj + 1, bk) if bj < by - 1 else -1,
           rk(bi, bj, bk - 1) if bk > 0 else -1, rk(bi, bj, bk + 1) if bk < bz - 1 else -1]
    proc_ranks = sorted(set(r for r in nbr if r >= 0))
    n_phys = len(patch_defs)
    for r in proc_ranks:
        patch_defs.append(dict(name=f"procBoundary{rank}to{r}", type="processor", myProcNo=rank, neighbProcNo=r))
    side_patch = [phys[s] if nbr[s] < 0 else n_phys + proc_ranks.index(nbr[s]) for s in range(6)]
    if z_cyclic and bz > 1:
        # the periodic pair is cut by the decomposition: the cyclic patches stay (empty) and their faces move to
        # processorCyclic patches towards the rank that holds the twin plane
        through = []
        if bk == 0:
            through.append((rk(bi, bj, bz - 1), "periodic_m", 4))
        if bk == bz - 1:
            through.append((rk(bi, bj, 0), "periodic_p", 5))
        for r, ref, s_ in sorted(through):
            side_patch[s_] = len(patch_defs)
            patch_defs.append(dict(name=f"procBoundary{rank}to{r}through{ref}", type="processorCyclic", myProcNo=rank, neighbProcNo=r, referPatch=ref))
    # a global id for every boundary face of the block (unique per geometric face, the same on both ranks)
    nf = len(side)
    n_i, n_j = (nx + 1) * ny * nz, nx * (ny + 1) * nz
    idx = np.arange(nf, dtype=np.int64)
    gid = np.zeros(nf, dtype=np.int64)
    # i-faces: local (i,j,k) from meshgrid(indexing="ij").ravel() -> i slowest
    li = idx[:n_i]
    i, j, k = li // (ny * nz), (li // nz) % ny, li % nz
    gid[:n_i] = (bi * nx + i) + (GX + 1) * ((bj * ny + j) + GY * (bk * nz + k))
    lj = idx[n_i:n_i + n_j] - n_i
    i, j, k = lj // ((ny + 1) * nz), (lj // nz) % (ny + 1), lj % nz
    gid[n_i:n_i + n_j] = (GX + 1) * GY * GZ + (bi * nx + i) + GX * ((bj * ny + j) + (GY + 1) * (bk * nz + k))
    lk = idx[n_i + n_j:] - n_i - n_j
    i, j, k = lk // (ny * (nz + 1)), (lk // (nz + 1)) % ny, lk % (nz + 1)
    gid[n_i + n_j:] = (GX + 1) * GY * GZ + GX * (GY + 1) * GZ + (bi * nx + i) + GX * ((bj * ny + j) + GY * (bk * nz + k))
    patch_id = np.full(nf, -1, dtype=np.int64)
    for s in range(6):
        patch_id[side == s] = side_patch[s]
    new_of_old = None
    if isinstance(tile, str) and tile == "morton":
        # the numbering `python -m lfm_public_b200.tools.renumber --method morton` gives this block: cells along a Z-order curve
        # through their centres (on a uniform block that is the curve through the cell indices)
        c = np.arange(nx * ny * nz, dtype=np.int64)
        order = np.argsort(_morton3(c % nx, (c // nx) % ny, c // (nx * ny)), kind="stable")
        new_of_old = np.empty(len(c), dtype=np.int64)
        new_of_old[order] = c
        a = new_of_old[a]
        b = np.where(b >= 0, new_of_old[np.maximum(b, 0)], -1)
    elif tile is not None:
        new_of_old = blocked_order(nx, ny, nz, tile, offset=(1, 1, 1), brick_order=brick_order)
        a = new_of_old[a]
        b = np.where(b >= 0, new_of_old[np.maximum(b, 0)], -1)
    # cyclic twins must share the same offset inside their patches: order the two z planes by (i,j) == gid within a plane
    m = assemble(points, faces, a, b, patch_id, patch_defs, nx * ny * nz, cyclic_keys=gid)
    # recover the id of each assembled face (assemble sorts; redo its ordering on gid)
    internal = b >= 0
    own = np.where(internal, np.minimum(a, b), a)
    nei = np.where(internal, np.maximum(a, b), -1)
    idx_int = np.nonzero(internal)[0]
    order_int = idx_int[np.lexsort((nei[idx_int], own[idx_int]))]
    idx_bnd = np.nonzero(~internal)[0]
    order_bnd = idx_bnd[np.lexsort((gid[idx_bnd], patch_id[idx_bnd]))]
    order = np.concatenate([order_int, order_bnd])
    m["faceProcAddressing"] = (gid[order] + 1).astype(np.int64)
    assert m["faceProcAddressing"].max() < 2 ** 31
    m["faceProcAddressing"] = m["faceProcAddressing"].astype(np.int32)
    m["logical"] = (nx, ny, nz)
    m["block"] = (bi, bj, bk)
    if new_of_old is not None:
        m["new_of_old"] = new_of_old
    return m


def renumber_cells(m, new_of_old):
    """Applies a cell permutation and restores canonical (upper-triangular) face order."""
    new_of_old = np.asarray(new_of_old, dtype=np.int64)
    nif = len(m["neighbour"])
    a = new_of_old[m["owner"].astype(np.int64)]
    b = np.full(len(a), -1, dtype=np.int64)
    b[:nif] = new_of_old[m["neighbour"].astype(np.int64)]
    pid = np.full(len(a), -1, dtype=np.int64)
    for i, p in enumerate(m["patches"]):
        pid[p["startFace"]:p["startFace"] + p["nFaces"]] = i
    out = assemble(m["points"], m["faces"], a, b, pid, m["patches"], m["nCells"])
    return out


def rcm_order(m):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    nif = len(m["neighbour"])
    o = m["owner"][:nif].astype(np.int64)
    n = m["neighbour"].astype(np.int64)
    nc = m["nCells"]
    g = coo_matrix((np.ones(2 * nif, dtype=np.int8), (np.concatenate([o, n]), np.concatenate([n, o]))), shape=(nc, nc)).tocsr()
    perm = reverse_cuthill_mckee(g, symmetric_mode=True)   # perm[new] = old
    new_of_old = np.empty(nc, dtype=np.int64)
    new_of_old[perm] = np.arange(nc)
    return new_of_old


# ------------------------------------------------------------------------------------------------
# decomposition into processorN meshes
# ------------------------------------------------------------------------------------------------

def block_assignment(m, blocks):
    """cell -> rank for a logical (ni,nj,nk) mesh split into blocks (bi,bj,bk); rank = bi_idx + bi*(bj_idx + bj*bk_idx)."""
    ni, nj, nk = m["logical"]
    bi, bj, bk = blocks
    c = np.arange(m["nCells"])
    i = c % ni
    j = (c // ni) % nj
    k = c // (ni * nj)
    return ((i * bi) // ni + bi * ((j * bj) // nj + bj * ((k * bk) // nk))).astype(np.int32)


def decompose(m, cell_rank):
    """Returns a list of per-rank mesh dicts with processor patches (ascending neighbour rank, faces in
    ascending global face id on both sides), faceProcAddressing (1-based, negative when the local face
    is flipped w.r.t. the global one), cellProcAddressing, pointProcAddressing, boundaryProcAddressing.
    A cyclic pair whose two cells live on different ranks becomes a pair of `processorCyclic` patches
    (`procBoundaryAtoBthrough<cyclic patch>`, referPatch = the cyclic patch the faces came from, faces in the
    order of the cyclic patch on both sides), as decomposePar does."""
    cell_rank = np.asarray(cell_rank, dtype=np.int64)
    n_ranks = int(cell_rank.max()) + 1
    nf = len(m["owner"])
    nif = len(m["neighbour"])
    own = m["owner"].astype(np.int64)
    nei = np.full(nf, -1, dtype=np.int64)
    nei[:nif] = m["neighbour"]
    r_own = cell_rank[own]
    r_nei = np.where(nei >= 0, cell_rank[np.maximum(nei, 0)], -1)
    pid = np.full(nf, -1, dtype=np.int64)
    twin_rank = np.full(nf, -1, dtype=np.int64)     # rank of the cyclic twin's cell where it differs from the face's own rank
    for i, p in enumerate(m["patches"]):
        if p["type"] == "cyclic":
            q = [x for x in m["patches"] if x["name"] == p["neighbourPatch"]][0]
            assert q["nFaces"] == p["nFaces"]
            ra = cell_rank[own[p["startFace"]:p["startFace"] + p["nFaces"]]]
            rb = cell_rank[own[q["startFace"]:q["startFace"] + q["nFaces"]]]
            twin_rank[p["startFace"]:p["startFace"] + p["nFaces"]] = np.where(ra != rb, rb, -1)
        pid[p["startFace"]:p["startFace"] + p["nFaces"]] = i
    out = []
    gfid = np.arange(nf, dtype=np.int64)
    for r in range(n_ranks):
        cells = np.nonzero(cell_rank == r)[0]
        local_of_global = np.full(m["nCells"], -1, dtype=np.int64)
        local_of_global[cells] = np.arange(len(cells))
        sel_int = (r_own == r) & (r_nei == r)
        sel_bnd = (r_own == r) & (nei < 0) & (twin_rank < 0)
        sel_pc = (r_own == r) & (nei < 0) & (twin_rank >= 0)   # cyclic faces whose twin went to another rank
        sel_po = (r_own == r) & (nei >= 0) & (r_nei != r)      # we hold the global owner: keep orientation
        sel_pn = (r_nei == r) & (r_own != r)                   # we hold the global neighbour: flip
        f_int = gfid[sel_int]
        f_bnd = gfid[sel_bnd]
        f_proc = np.concatenate([gfid[sel_po], gfid[sel_pn]])
        flip_proc = np.concatenate([np.zeros(sel_po.sum(), bool), np.ones(sel_pn.sum(), bool)])
        other = np.concatenate([r_nei[sel_po], r_own[sel_pn]])
        o = np.lexsort((f_proc, other))
        f_proc, flip_proc, other = f_proc[o], flip_proc[o], other[o]
        f_pc = gfid[sel_pc]
        o = np.lexsort((f_pc, pid[f_pc], twin_rank[f_pc]))     # by neighbour rank, then cyclic patch, then position in it
        f_pc = f_pc[o]
        f_all = np.concatenate([f_int, f_bnd, f_proc, f_pc])
        flip = np.concatenate([np.zeros(len(f_int) + len(f_bnd), bool), flip_proc, np.zeros(len(f_pc), bool)])
        faces = _flip(m["faces"][f_all], flip)
        l_own = np.where(flip, local_of_global[np.maximum(nei[f_all], 0)], local_of_global[own[f_all]])
        l_nei = local_of_global[nei[f_int]]
        # points
        used = np.unique(faces[faces >= 0])
        lp = np.full(len(m["points"]), -1, dtype=np.int64)
        lp[used] = np.arange(len(used))
        faces_l = np.where(faces >= 0, lp[np.maximum(faces, 0)], -1).astype(np.int32)
        patches = []
        start = len(f_int)
        for i, p in enumerate(m["patches"]):
            n = int(np.count_nonzero(pid[f_bnd] == i))
            d = {k: v for k, v in p.items() if k not in ("nFaces", "startFace")}
            d["nFaces"], d["startFace"] = n, start
            start += n
            patches.append(d)
        bpa = list(range(len(m["patches"])))
        for nb in np.unique(other):
            n = int(np.count_nonzero(other == nb))
            patches.append(dict(name=f"procBoundary{r}to{int(nb)}", type="processor", myProcNo=r, neighbProcNo=int(nb),
                                nFaces=n, startFace=start))
            start += n
            bpa.append(-1)
        for nb in np.unique(twin_rank[f_pc]):
            for i in np.unique(pid[f_pc][twin_rank[f_pc] == nb]):
                n = int(np.count_nonzero((twin_rank[f_pc] == nb) & (pid[f_pc] == i)))
                ref = m["patches"][int(i)]["name"]
                patches.append(dict(name=f"procBoundary{r}to{int(nb)}through{ref}", type="processorCyclic", myProcNo=r,
                                    neighbProcNo=int(nb), referPatch=ref, nFaces=n, startFace=start))
                start += n
                bpa.append(-1)
        # canonical order inside rank: internal faces are already (owner, neighbour)-sorted because the
        # local numbering preserves the global cell order
        assert (np.diff(l_own[:len(f_int)]) >= 0).all()
        out.append(dict(points=m["points"][used], faces=faces_l, owner=l_own.astype(np.int32), neighbour=l_nei.astype(np.int32),
                        patches=patches, nCells=len(cells),
                        faceProcAddressing=np.where(flip, -(f_all + 1), f_all + 1).astype(np.int32),
                        cellProcAddressing=cells.astype(np.int32), pointProcAddressing=used.astype(np.int32),
                        boundaryProcAddressing=np.asarray(bpa, dtype=np.int32)))
    return out


# ------------------------------------------------------------------------------------------------
# geometry helpers used for initial fields (plain numpy restatement for generators only)
# ------------------------------------------------------------------------------------------------

def cell_centres_estimate(m):
    """Average of the cell's face-centre averages (good enough for analytic initial fields)."""
    f = m["faces"]
    valid = f >= 0
    cnt = valid.sum(1)
    fc = (m["points"][np.maximum(f, 0)] * valid[..., None]).sum(1) / cnt[:, None]
    nc = m["nCells"]
    acc = np.zeros((nc, 3))
    num = np.zeros(nc)
    np.add.at(acc, m["owner"], fc)
    np.add.at(num, m["owner"], 1)
    nif = len(m["neighbour"])
    np.add.at(acc, m["neighbour"], fc[:nif])
    np.add.at(num, m["neighbour"], 1)
    return acc / num[:, None]


# ------------------------------------------------------------------------------------------------
# writers
# ------------------------------------------------------------------------------------------------
_HDR = "FoamFile\n{{\n    version     2.0;\n    format      ascii;\n    class       {cls};\n    object      {obj};\n}}\n\n"


def _write_labels(path, obj, arr):
    with open(path, "w") as f:
        f.write(_HDR.format(cls="labelList", obj=obj))
        f.write(f"{len(arr)}\n(\n")
        f.write("\n".join(map(str, np.asarray(arr).tolist())))
        f.write("\n)\n")


def write_polymesh(m, mesh_dir):
    os.makedirs(mesh_dir, exist_ok=True)
    with open(os.path.join(mesh_dir, "points"), "w") as f:
        f.write(_HDR.format(cls="vectorField", obj="points"))
        f.write(f"{len(m['points'])}\n(\n")
        f.write("\n".join("(%.17g %.17g %.17g)" % tuple(p) for p in m["points"].tolist()))
        f.write("\n)\n")
    with open(os.path.join(mesh_dir, "faces"), "w") as f:
        f.write(_HDR.format(cls="faceList", obj="faces"))
        f.write(f"{len(m['faces'])}\n(\n")
        lines = []
        for row in m["faces"].tolist():
            if row[3] < 0:
                lines.append("3(%d %d %d)" % tuple(row[:3]))
            else:
                lines.append("4(%d %d %d %d)" % tuple(row))
        f.write("\n".join(lines))
        f.write("\n)\n")
    _write_labels(os.path.join(mesh_dir, "owner"), "owner", m["owner"])
    _write_labels(os.path.join(mesh_dir, "neighbour"), "neighbour", m["neighbour"])
    with open(os.path.join(mesh_dir, "boundary"), "w") as f:
        f.write(_HDR.format(cls="polyBoundaryMesh", obj="boundary"))
        f.write(f"{len(m['patches'])}\n(\n")
        for p in m["patches"]:
            f.write(f"    {p['name']}\n    {{\n        type            {p['type']};\n")
            if p["type"] == "cyclic":
                f.write(f"        neighbourPatch  {p['neighbourPatch']};\n")
            if p["type"] in ("processor", "processorCyclic"):
                f.write(f"        myProcNo        {p['myProcNo']};\n        neighbProcNo    {p['neighbProcNo']};\n")
            if p["type"] == "processorCyclic":
                f.write(f"        referPatch      {p['referPatch']};\n")
            f.write(f"        nFaces          {p['nFaces']};\n        startFace       {p['startFace']};\n    }}\n")
        f.write(")\n")
    for key in ("faceProcAddressing", "cellProcAddressing", "pointProcAddressing", "boundaryProcAddressing", "cellSubmesh"):
        if key in m:
            _write_labels(os.path.join(mesh_dir, key), key, m[key])


def write_field(path, name, m, values):
    values = np.asarray(values, dtype=np.float64)
    vec = values.ndim == 2
    with open(path, "w") as f:
        f.write(_HDR.format(cls="volVectorField" if vec else "volScalarField", obj=name))
        f.write("dimensions      [0 0 0 0 0 0 0];\n\n")
        f.write(f"internalField   nonuniform List<{'vector' if vec else 'scalar'}> \n{len(values)}\n(\n")
        if vec:
            f.write("\n".join("(%.17g %.17g %.17g)" % tuple(v) for v in values.tolist()))
        else:
            f.write("\n".join("%.17g" % v for v in values.tolist()))
        f.write("\n)\n;\n\nboundaryField\n{\n")
        for p in m["patches"]:
            t = {"empty": "empty", "cyclic": "cyclic", "processor": "processor", "processorCyclic": "processorCyclic"}.get(p["type"], "zeroGradient")
            f.write(f"    {p['name']}\n    {{\n        type            {t};\n    }}\n")
        f.write("}\n")
