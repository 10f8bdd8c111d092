"""Marching-cubes case table, DERIVED (not transcribed) from the cube's topology.

The reference meshes the logit volume with PyMCubes (`mcubes.marching_cubes(pred, 0)`, triplane_decoder/
visualize.py:100; the package is an un-pinned dependency in triplane_decoder/environment.yml and is not installed
offline).  PyMCubes implements Lorensen-Cline marching cubes on Bourke's corner / edge numbering, which is what
this table follows:

    corner i at (x,y,z) offsets  0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1)
    edge   e joins corners       0:(0,1) 1:(1,2) 2:(2,3) 3:(3,0) 4:(4,5) 5:(5,6) 6:(6,7) 7:(7,4) 8:(0,4) 9:(1,5) 10:(2,6) 11:(3,7)
    case bit i is set when value(corner i) < isovalue

The vertex set of marching cubes does not depend on the table at all (one vertex on every grid edge whose end points
lie on different sides); only the triangulation inside a cell does.  Bourke's 256x16 table itself is not available
offline, so the triangulation is generated from first principles and is guaranteed watertight by construction:
  * on every cube FACE the crossed edges are joined pairwise; a face with four crossed edges (two diagonal corners
    below the isovalue) is resolved by cutting off each BELOW corner — a rule that only looks at the face, so the two
    cells sharing the face always agree (no holes, unlike the original 15-case table);
  * the face segments chain into closed loops of crossed edges; every loop is triangulated (a fan where possible)
    so that no triangle lies inside a cube face (the neighbouring cell would emit its mirror image there);
  * loops are oriented so that the normal points to the BELOW side (towards smaller values: out of the solid for
    occupancy logits, which are positive inside).
Triangle COUNTS in ambiguous cells may therefore differ from PyMCubes' (parity with the binary package is unpinned —
DESIGN.md); vertex positions and the surface topology class do not.
"""
from __future__ import annotations

import numpy as np

CORNERS = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# each edge as (owning corner, axis): the grid edge starts at that corner and runs along +axis
EDGE_OWNER = [(0, 0), (1, 1), (3, 0), (0, 1), (4, 0), (5, 1), (7, 0), (4, 1), (0, 2), (1, 2), (2, 2), (3, 2)]
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]   # cyclic corner order
_EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}


def _case_loops(case: int):
    below = [(case >> i) & 1 for i in range(8)]
    nbr = {}

    def link(a, b):
        nbr.setdefault(a, []).append(b)
        nbr.setdefault(b, []).append(a)

    for face in FACES:
        fe = [_EDGE_ID[frozenset((face[k], face[(k + 1) % 4]))] for k in range(4)]     # edge k joins corner k, k+1
        crossed = [k for k in range(4) if below[face[k]] != below[face[(k + 1) % 4]]]
        if len(crossed) == 2:
            link(fe[crossed[0]], fe[crossed[1]])
        elif len(crossed) == 4:
            for k in range(4):                       # cut off every BELOW corner: its two face edges are joined
                if below[face[k]]:
                    link(fe[(k - 1) % 4], fe[k])
    loops, seen = [], set()
    for start in sorted(nbr):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = nbr[cur]
            nxt = a if a != prev else b
            if len(loop) > 1 and nxt == start:
                break
            if nxt in seen:                          # two-edge degenerate cannot happen on a cube; guard anyway
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        loops.append(loop)
    # orientation: normal (Newell, edge mid-points) towards the below end points of the loop's own edges
    mids = np.array([(CORNERS[a] + CORNERS[b]) / 2.0 for a, b in EDGES])
    out = []
    for loop in loops:
        p = mids[loop]
        n = np.zeros(3)
        for k in range(len(loop)):
            a, b = p[k], p[(k + 1) % len(loop)]
            n += np.cross(a, b)
        ends = np.array([CORNERS[a] if below[a] else CORNERS[b] for a, b in (EDGES[e] for e in loop)], dtype=float)
        if np.dot(n, ends.mean(0) - p.mean(0)) < 0:
            loop = [loop[0]] + loop[:0:-1]
        out.append(loop)
    return out


_FACE_EDGES = [frozenset(_EDGE_ID[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)) for f in FACES]


def _in_one_face(a, b, c):
    return any({a, b, c} <= fe for fe in _FACE_EDGES)


def _triangulations(poly):
    """All triangulations of a convex-position polygon (vertex list), fans first."""
    n = len(poly)
    if n < 3:
        yield []
        return
    if n == 3:
        yield [tuple(poly)]
        return
    # split on the triangle (poly[0], poly[k], poly[-1])
    for k in range(1, n - 1):
        for left in _triangulations(poly[:k + 1]):
            for right in _triangulations(poly[k:]):
                yield left + [(poly[0], poly[k], poly[-1])] + right


def _triangulate(loop):
    """A triangulation of the loop with NO triangle lying inside a cube face: such a triangle has zero thickness
    against the neighbouring cell (which would emit its mirror image) and makes the edge 4-valent.  Fans are tried
    first (every apex), then every other triangulation; all orientations follow the loop."""
    n = len(loop)
    for apex in range(n):
        rot = loop[apex:] + loop[:apex]
        tris = [(rot[0], rot[k], rot[k + 1]) for k in range(1, n - 1)]
        if not any(_in_one_face(*t) for t in tris):
            return tris
    for tris in _triangulations(loop):
        if not any(_in_one_face(*t) for t in tris):
            return tris
    raise AssertionError(f"no face-free triangulation for loop {loop}")


def build_table():
    """(tri_table int8 [256, 3*max_tris] padded with -1, tri_count uint8 [256])."""
    rows = []
    for case in range(256):
        tris = []
        for loop in _case_loops(case):
            for t in _triangulate(loop):
                tris += list(t)
        rows.append(tris)
    width = max(len(r) for r in rows)
    table = -np.ones((256, width), dtype=np.int8)
    for c, r in enumerate(rows):
        table[c, :len(r)] = r
    counts = np.array([len(r) // 3 for r in rows], dtype=np.uint8)
    return table, counts


TRI_TABLE, TRI_COUNT = build_table()
