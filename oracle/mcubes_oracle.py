"""CPU restatement of the meshing tail of the decode path — TEST INFRASTRUCTURE ONLY (never imported by the product).

  * marching_cubes(volume, isovalue): the algorithm PyMCubes runs for the reference (`mcubes.marching_cubes(pred, 0)`,
    /root/reference/triplane_decoder/visualize.py:100): Lorensen-Cline marching cubes on Bourke's corner / edge
    numbering, one shared vertex per crossed grid edge by linear interpolation, vertices in index coordinates.
    PyMCubes is an un-pinned third-party dependency (triplane_decoder/environment.yml:19) that is not installed here,
    and the reference ships no mesh fixture: PARITY WITH THE BINARY PACKAGE IS UNPINNED.  What is pinned:
      - the vertex SET is table-independent and is checked against an independent count of crossed grid edges;
      - the case table is DERIVED (ishapediting_b200/triplane_decoder/mc_table.py) and checked here by invariants
        (check_table): every triangle corner lies on a crossed edge, every crossed edge is used, complementary /
        face-adjacent cases agree on shared faces (watertightness), orientation is consistent;
      - meshes of analytic volumes are closed 2-manifolds with the right Euler characteristic.
  * filter_smooth_simple(vertices, triangles, iterations): Open3D 0.18 TriangleMesh::FilterSmoothSimple
    (/root/reference/drag_utils.py:300, `filter_smooth_simple(number_of_iterations=10)`): per iteration
    v_i <- (v_i + sum_{j in N(i)} v_j) / (1 + |N(i)|) with N(i) the vertices sharing a triangle edge with i
    (unique), all vertices updated from the previous iterate, float64.
"""
from __future__ import annotations

import numpy as np

from ishapediting_b200.triplane_decoder.mc_table import CORNERS, EDGES, EDGE_OWNER, FACES, TRI_COUNT, TRI_TABLE


def check_table():
    """Invariants of the derived case table; raises AssertionError with the offending case."""
    edge_id = {frozenset(e): i for i, e in enumerate(EDGES)}
    for case in range(256):
        below = [(case >> i) & 1 for i in range(8)]
        crossed = {i for i, (a, b) in enumerate(EDGES) if below[a] != below[b]}
        tris = [int(e) for e in TRI_TABLE[case] if e >= 0]
        assert len(tris) == 3 * int(TRI_COUNT[case]), case
        assert set(tris) == crossed, (case, sorted(set(tris)), sorted(crossed))
        # every directed triangle edge appears once; its reverse appears once too unless it lies in a cube face
        half = {}
        for t in range(0, len(tris), 3):
            a, b, c = tris[t:t + 3]
            assert len({a, b, c}) == 3, case
            for u, v in ((a, b), (b, c), (c, a)):
                assert (u, v) not in half, (case, u, v)
                half[(u, v)] = True
        face_segments = {f: set() for f in range(6)}
        for (u, v) in half:
            if (v, u) in half:
                continue                                   # interior edge of the cell's patch
            shared = [f for f, face in enumerate(FACES)
                      if {*EDGES[u]} <= set(face) and {*EDGES[v]} <= set(face)]
            assert len(shared) == 1, (case, u, v, shared)   # an open edge must lie in exactly one cube face
            face_segments[shared[0]].add((u, v))
        # the face rule: a face's open segments depend only on the four corner states of that face, and the cell on
        # the other side (same face pattern) must produce the same segments reversed -> checked through `face_rule`
        for f, face in enumerate(FACES):
            pat = tuple(below[c] for c in face)
            segs = frozenset(frozenset((u, v)) for u, v in face_segments[f])
            fe = [edge_id[frozenset((face[k], face[(k + 1) % 4]))] for k in range(4)]
            local = frozenset(frozenset((fe.index(u), fe.index(v))) for u, v in (tuple(s) for s in segs))
            key = pat
            if key in _FACE_RULE:
                assert _FACE_RULE[key] == local, (case, f, pat)
            else:
                _FACE_RULE[key] = local
    return True


_FACE_RULE: dict = {}


def marching_cubes(volume, isovalue=0.0):
    """volume (X,Y,Z) float32 -> (vertices (V,3) float32 in index coordinates, triangles (T,3) int32).
    Vertex order: by grid point (x-major linear index), then by axis x,y,z of the edge it owns; triangle order: by
    cell (x-major), then table order — the order the CUDA kernels emit, so results can be compared element-wise."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    X, Y, Z = vol.shape
    iso = np.float32(isovalue)
    below = vol < iso
    flags = np.zeros((X, Y, Z, 3), dtype=bool)
    flags[:-1, :, :, 0] = below[:-1] != below[1:]
    flags[:, :-1, :, 1] = below[:, :-1] != below[:, 1:]
    flags[:, :, :-1, 2] = below[:, :, :-1] != below[:, :, 1:]
    flat = flags.reshape(-1)
    vid = np.cumsum(flat, dtype=np.int64) - flat          # exclusive scan in (point, axis) order
    vid = vid.reshape(X, Y, Z, 3)
    nv = int(flat.sum())
    verts = np.zeros((nv, 3), dtype=np.float32)
    for ax in range(3):
        idx = np.argwhere(flags[..., ax])
        if idx.size == 0:
            continue
        p0 = vol[idx[:, 0], idx[:, 1], idx[:, 2]]
        q = idx.copy()
        q[:, ax] += 1
        p1 = vol[q[:, 0], q[:, 1], q[:, 2]]
        t = ((iso - p0) / (p1 - p0)).astype(np.float32)
        pos = idx.astype(np.float32)
        pos[:, ax] = pos[:, ax] + t
        verts[vid[idx[:, 0], idx[:, 1], idx[:, 2], ax]] = pos
    # cells
    b = below.astype(np.int32)
    case = np.zeros((X - 1, Y - 1, Z - 1), dtype=np.int32)
    for i, (dx, dy, dz) in enumerate(CORNERS):
        case |= b[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz] << i
    cells = np.argwhere((case != 0) & (case != 255))
    ccase = case[cells[:, 0], cells[:, 1], cells[:, 2]]
    cnt = TRI_COUNT[ccase].astype(np.int64)
    toff = np.cumsum(cnt) - cnt
    nt = int(cnt.sum())
    tris = np.zeros((nt, 3), dtype=np.int32)
    width = TRI_TABLE.shape[1] // 3
    owner = np.array(EDGE_OWNER)
    for k in range(width):
        sel = cnt > k
        if not sel.any():
            break
        cc, cs = cells[sel], ccase[sel]
        for j in range(3):
            e = TRI_TABLE[cs, 3 * k + j].astype(np.int64)
            corner, axis = owner[e, 0], owner[e, 1]
            off = CORNERS[corner]
            tris[toff[sel] + k, j] = vid[cc[:, 0] + off[:, 0], cc[:, 1] + off[:, 1], cc[:, 2] + off[:, 2], axis]
    return verts, tris


def crossed_edge_count(volume, isovalue=0.0):
    """Independent count of the marching-cubes vertex set: grid edges whose end points lie on different sides."""
    below = np.asarray(volume) < isovalue
    return int((below[:-1] != below[1:]).sum() + (below[:, :-1] != below[:, 1:]).sum() + (below[:, :, :-1] != below[:, :, 1:]).sum())


def filter_smooth_simple(vertices, triangles, iterations=1):
    """Open3D TriangleMesh::FilterSmoothSimple (see module docstring).  float64 like Open3D."""
    v = np.asarray(vertices, dtype=np.float64).copy()
    t = np.asarray(triangles, dtype=np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]], t[:, [1, 0]], t[:, [2, 1]], t[:, [0, 2]]], axis=0)
    e = np.unique(e, axis=0)                              # unique directed pairs (i -> neighbour j)
    deg = np.bincount(e[:, 0], minlength=len(v)).astype(np.float64)
    for _ in range(iterations):
        s = v.copy()
        np.add.at(s, e[:, 0], v[e[:, 1]])
        v = s / (1.0 + deg)[:, None]
    return v


def mesh_report(vertices, triangles):
    """Topology numbers of a triangle mesh: (closed manifold?, consistently oriented?, Euler characteristic, volume)."""
    t = np.asarray(triangles, dtype=np.int64)
    he = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
    key = he[:, 0] * (int(t.max()) + 1) + he[:, 1]
    uniq_directed = len(np.unique(key)) == len(key)
    und = np.sort(he, axis=1)
    _, counts = np.unique(und, axis=0, return_counts=True)
    closed = bool((counts == 2).all())
    rev = he[:, 1] * (int(t.max()) + 1) + he[:, 0]
    oriented = uniq_directed and bool(np.isin(rev, key).all())
    V, E, Fc = len(np.unique(t)), len(counts), len(t)
    v = np.asarray(vertices, dtype=np.float64)
    vol6 = np.einsum("ij,ij->i", v[t[:, 0]], np.cross(v[t[:, 1]], v[t[:, 2]])).sum()
    return dict(closed=closed, oriented=oriented, euler=V - E + Fc, volume=vol6 / 6.0)
