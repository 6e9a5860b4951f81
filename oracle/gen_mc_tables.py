"""Generate the marching-cubes case tables programmatically (no table could be downloaded:
scikit-image / Lewiner's LUTs are not in this image - SURVEY.md §8(c), "parity unpinned").

Writes rgb-d-pifuhd_b200/csrc/mc_tables.h, shared by the CUDA kernels and the C oracle.

Conventions (as SURVEY §8(c) recalls scikit-image's `marching_cubes_lewiner`):
  * volume im[a0, a1, a2]; cube corner i at offsets (d2, d1, d0) along (axis2, axis1, axis0):
      v0 (0,0,0) v1 (1,0,0) v2 (1,1,0) v3 (0,1,0) v4 (0,0,1) v5 (1,0,1) v6 (1,1,1) v7 (0,1,1)
  * case index bit i set iff value(v_i) > level (strict)
  * edges e0..e11 = v0v1 v1v2 v2v3 v3v0 v4v5 v5v6 v6v7 v7v4 v0v4 v1v5 v2v6 v3v7

Topology rule: on every cube face the cut edges are joined by segments; a face with two
diagonal inside corners (ambiguous) always separates the inside corners.  The rule depends
only on the face's own corner states, so two cells sharing a face draw the same segments
and the surface is watertight (unlike the classic 15-case table with complement symmetry).
Segments chain into closed loops, each loop is triangulated (fan from its lowest edge id unless a
diagonal would lie inside a cube face) and oriented so the normal points towards decreasing values ('descent').
"""
import itertools
import os

import numpy as np

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]  # (d2, d1, d0)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}


def face_segments(case, face):
    inside = [(case >> c) & 1 for c in face]
    cuts = []  # (edge id, index k of the cycle edge between face[k] and face[k+1])
    for k in range(4):
        a, b = face[k], face[(k + 1) % 4]
        if inside[k] != inside[(k + 1) % 4]:
            cuts.append((EDGE_OF[frozenset((a, b))], k))
    if not cuts:
        return []
    if len(cuts) == 2:
        return [(cuts[0][0], cuts[1][0])]
    # ambiguous: corners alternate.  Separate the inside corners: each inside corner face[k]
    # is cut off by the segment joining cycle edges k-1 and k.
    segs = []
    by_k = {k: e for e, k in cuts}
    for k in range(4):
        if inside[k]:
            segs.append((by_k[(k - 1) % 4], by_k[k]))
    return segs


def corner_xyz(c):
    """Corner position in the OUTPUT frame (axis0, axis1, axis2) - the reverse of (d2, d1, d0).
    Orientation is decided in this frame; the reversed one is its mirror image."""
    return np.array(CORNERS[c][::-1], float)


def edge_mid(e):
    a, b = EDGES[e]
    return (corner_xyz(a) + corner_xyz(b)) / 2


def loops_of(case):
    segs = []
    for f in FACES:
        segs += face_segments(case, f)
    adj = {}
    for a, b in segs:
        adj.setdefault(a, []).append(b)
        adj.setdefault(b, []).append(a)
    assert all(len(v) == 2 for v in adj.values()), (case, adj)
    seen, loops = set(), []
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [n for n in adj[cur] if n != prev]
            n = nxt[0] if nxt else adj[cur][0]
            if adj[cur][0] == adj[cur][1]:
                n = adj[cur][0]
            if n == start:
                break
            loop.append(n)
            seen.add(n)
            prev, cur = cur, n
        loops.append(loop)
    return loops


def orient(case, loop):
    pts = np.array([edge_mid(e) for e in loop])
    cen = pts.mean(0)
    nrm = np.zeros(3)
    for i in range(len(loop)):
        nrm += np.cross(pts[i] - cen, pts[(i + 1) % len(loop)] - cen)
    # every loop vertex sits on an edge with one inside and one outside end: the normal must
    # point from the inside end to the outside end
    s = 0.0
    for e in loop:
        a, b = EDGES[e]
        if (case >> a) & 1:
            a, b = b, a                     # a outside, b inside
        s += float(np.dot(nrm, corner_xyz(a) - corner_xyz(b)))
    assert abs(s) > 1e-9, (case, loop)
    # normal must point from inside (high) to outside (low): s > 0
    if s < 0:
        loop = [loop[0]] + loop[:0:-1]
    return loop


FACE_EDGES = [set(EDGE_OF[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)) for f in FACES]


def in_face(e1, e2):
    return any(e1 in fe and e2 in fe for fe in FACE_EDGES)


def triangulations(poly):
    """All triangulations of a convex polygon given as a vertex list (orientation preserved)."""
    if len(poly) < 3:
        return [[]]
    if len(poly) == 3:
        return [[tuple(poly)]]
    out = []
    a, b = poly[0], poly[-1]
    for m in range(1, len(poly) - 1):
        for left in triangulations(poly[:m + 1]):
            for right in triangulations(poly[m:]):
                out.append(left + [(a, poly[m], b)] + right)
    return out


def triangulate(loop):
    """Fan from the lowest edge id when none of its diagonals lies inside a cube face (a
    diagonal in a face would coincide with the neighbour cell's and make a non-manifold
    edge); otherwise the first triangulation (in enumeration order) free of such diagonals."""
    def bad(tris):
        n = 0
        boundary = set()
        for i in range(len(loop)):
            boundary.add(frozenset((loop[i], loop[(i + 1) % len(loop)])))
        for t in tris:
            for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if frozenset((x, y)) not in boundary and in_face(x, y):
                    n += 1
        return n
    fan = [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
    if bad(fan) == 0:
        return fan
    best = None
    for cand in triangulations(loop):
        b = bad(cand)
        if b == 0:
            return cand
        if best is None or b < best[0]:
            best = (b, cand)
    return best[1]


def build():
    tris, verts = [], []
    for case in range(256):
        t = []
        for loop in loops_of(case):
            loop = orient(case, loop)
            for tri in triangulate(loop):
                t += list(tri)
        order = []
        for e in t:
            if e not in order:
                order.append(e)
        tris.append(t)
        verts.append(order)
    return tris, verts


def edge_geometry():
    """Per edge: axis (0 = volume axis 0 ... 2 = axis 2), corner offsets of its lower end as
    (o0, o1, o2) along (axis0, axis1, axis2), low mask (perpendicular axes at offset 0)."""
    geo = []
    for a, b in EDGES:
        ca, cb = CORNERS[a], CORNERS[b]
        lo = tuple(min(x, y) for x, y in zip(ca, cb))       # (d2, d1, d0)
        diff = [abs(x - y) for x, y in zip(ca, cb)]
        d_idx = diff.index(1)                               # 0 -> axis2, 1 -> axis1, 2 -> axis0
        axis = 2 - d_idx
        o = (lo[2], lo[1], lo[0])                           # along (axis0, axis1, axis2)
        low_mask = 0
        for ax in range(3):
            if ax != axis and o[ax] == 0:
                low_mask |= 1 << ax
        geo.append((axis, o, low_mask))
    return geo


def shifted_edge(geo, e, mask):
    """Edge id of the same lattice edge seen from the cell shifted by -1 on the axes in mask."""
    axis, o, _ = geo[e]
    o2 = tuple(o[ax] + (1 if (mask >> ax) & 1 else 0) for ax in range(3))
    for i, (ax, oo, _) in enumerate(geo):
        if ax == axis and oo == o2:
            return i
    return -1


def main():
    tris, verts = build()
    geo = edge_geometry()
    max_t = max(len(t) for t in tris) // 3
    max_v = max(len(v) for v in verts)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "rgb-d-pifuhd_b200", "csrc", "mc_tables.h")
    L = []
    L.append("// GENERATED by oracle/gen_mc_tables.py - do not edit.  See that file for the conventions.")
    L.append("#pragma once")
    L.append("#define MC_MAX_TRIS %d" % max_t)
    L.append("#define MC_MAX_VERTS %d" % max_v)
    L.append("#ifndef MC_TABLE_QUALIFIER\n#define MC_TABLE_QUALIFIER static const\n#endif")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_NTRI[256] = {%s};" % ",".join(str(len(t) // 3) for t in tris))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_NVERT[256] = {%s};" % ",".join(str(len(v)) for v in verts))
    rows = []
    for t in tris:
        rows.append("{%s}" % ",".join(str(x) for x in t + [-1] * (3 * max_t - len(t))))
    L.append("MC_TABLE_QUALIFIER signed char MC_TRIS[256][%d] = {\n%s};" % (3 * max_t, ",\n".join(rows)))
    rows = []
    for v in verts:
        rows.append("{%s}" % ",".join(str(x) for x in v + [-1] * (max_v - len(v))))
    L.append("MC_TABLE_QUALIFIER signed char MC_VERTS[256][%d] = {\n%s};" % (max_v, ",\n".join(rows)))
    L.append("// corner offsets along (axis0, axis1, axis2)")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_CORNER[8][3] = {%s};" %
             ",".join("{%d,%d,%d}" % (c[2], c[1], c[0]) for c in CORNERS))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_CORNERS[12][2] = {%s};" %
             ",".join("{%d,%d}" % e for e in EDGES))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_AXIS[12] = {%s};" % ",".join(str(g[0]) for g in geo))
    L.append("// perpendicular axes on which the edge sits at offset 0 (shared with the previous cell)")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_LOWMASK[12] = {%s};" % ",".join(str(g[2]) for g in geo))
    rows = []
    for e in range(12):
        rows.append("{%s}" % ",".join(str(shifted_edge(geo, e, m)) for m in range(8)))
    L.append("// id of edge e as seen from the cell shifted by -1 on the axes of mask m (-1: impossible)")
    L.append("MC_TABLE_QUALIFIER signed char MC_EDGE_SHIFT[12][8] = {\n%s};" % ",\n".join(rows))
    with open(out, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", out, "max tris", max_t, "max verts", max_v)
    return tris, verts


if __name__ == "__main__":
    main()
