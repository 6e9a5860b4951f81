"""Generate the marching-cubes tables of the CUDA kernels (rgb-d-pifuhd_b200/csrc/mc_tables.h).

scikit-image's `marching_cubes_lewiner` (`mesh_util.py:84`) and Lewiner's lookup tables are not available in
this image ("parity unpinned"), so the tables are derived from a rule, stated here once and
implemented twice, independently: by this generator (the tables the CUDA kernels read) and, at run time and from
the corner values alone, by the C oracle (oracle/mc_ref.c - it includes no table).  A wrong table entry therefore
fails the bit-exact GPU-vs-oracle tests.

THE RULE
  * volume im[a0, a1, a2]; cube corner i at offsets (d2, d1, d0) along (axis2, axis1, axis0):
      v0 (0,0,0) v1 (1,0,0) v2 (1,1,0) v3 (0,1,0) v4 (0,0,1) v5 (1,0,1) v6 (1,1,1) v7 (0,1,1)
    case bit i set iff value(v_i) > level (strict); edges e0..e11 =
      v0v1 v1v2 v2v3 v3v0 v4v5 v5v6 v6v7 v7v4 v0v4 v1v5 v2v6 v3v7;
    faces, as corner cycles, f0..f5 = (0,1,2,3) (4,5,6,7) (0,1,5,4) (3,2,6,7) (0,3,7,4) (1,2,6,5).
  * On every face the cut edges are joined by segments.  Two cuts: one segment.  Four cuts (inside corners on a
    diagonal - the ambiguous face): Lewiner's face test (Lewiner et al. 2003, `test_face`; the asymptotic decider):
    with a, c the inside and b, d the outside corner values minus the level, the inside corners are JOINED
    across the face when a*c > b*d (the bilinear interpolant's saddle value is inside), otherwise SEPARATED
    (ties: separated).  Joined: each OUTSIDE corner is cut off by a segment; separated: each INSIDE corner is.
    The decision depends only on the face's four values, so the two cells that share the face draw the same
    segments and the surface is watertight.
  * Segments chain into closed loops.  A loop starts at its lowest edge id and runs in the direction that makes
    its normal point from the inside (high values) to the outside ('descent').  Loops are taken in ascending
    order of their first edge.
  * A loop is triangulated as the fan from its first vertex unless one of the fan's diagonals lies inside a cube
    face; then the first triangulation, in the enumeration order of `triangulations`, with the fewest such
    diagonals.
  * What is NOT done (deviations from Lewiner's 33 cases, listed in DESIGN.md): the interior test (tunnels
    through the cell, sub-cases 4.1.2, 6.1.2, 7.4.2, 10.1.2, 12.1.2, 13.5.x) and the centre vertex.

A cell's table row is selected by (case, decisions of its ambiguous faces): sub = MC_SUB_BASE[case] + the
decision bits of the faces in MC_AMB[case], packed in ascending face order.
"""
import os

import numpy as np

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]  # (d2, d1, d0)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}


def face_ambiguous(case, face):
    ins = [(case >> c) & 1 for c in face]
    return ins[0] == ins[2] and ins[1] == ins[3] and ins[0] != ins[1]


def face_segments(case, face, joined):
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
    # ambiguous: corners alternate.  Cut off the outside corners (inside joined) or the inside corners (separated):
    # corner face[k] is cut off by the segment joining cycle edges k-1 and k.
    by_k = {k: e for e, k in cuts}
    return [(by_k[(k - 1) % 4], by_k[k]) for k in range(4) if bool(inside[k]) != bool(joined)]


def corner_xyz(c):
    """Corner position in the OUTPUT frame (axis0, axis1, axis2) - the reverse of (d2, d1, d0)."""
    return np.array(CORNERS[c][::-1], float)


def edge_mid(e):
    a, b = EDGES[e]
    return (corner_xyz(a) + corner_xyz(b)) / 2


def loops_of(case, joined_mask):
    segs = []
    for fi, f in enumerate(FACES):
        segs += face_segments(case, f, (joined_mask >> fi) & 1)
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
            n = adj[cur][0] if adj[cur][0] != prev else adj[cur][1]
            if prev is None:
                n = adj[cur][0]
            if n == start and len(loop) > 1:
                break
            loop.append(n)
            seen.add(n)
            prev, cur = cur, n
        loops.append(loop)
    return loops


FACE_EDGES = [set(EDGE_OF[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)) for f in FACES]


def orient(case, loop):
    """Direction of a loop, decided on its first segment p = loop[0] -> q = loop[1] (both on one cube face f, outward
    normal N): the surface leaves the segment towards the cube's interior (-N), so its normal there is d x (-N) with
    d = mid(q) - mid(p); it must point to the outside end of edge p (w = outside corner - inside corner of p).
    Reverse the loop, keeping its first vertex, when dot(d x (-N), w) < 0.  (A centroid-based normal vanishes by
    symmetry on the 8-edge loops of the joined cases.)"""
    p, q = loop[0], loop[1]
    f = [fi for fi, fe in enumerate(FACE_EDGES) if p in fe and q in fe]
    assert len(f) == 1, (case, loop)
    cen = sum(corner_xyz(c) for c in FACES[f[0]]) / 4.0
    nrm = 2.0 * (cen - np.array([0.5, 0.5, 0.5]))
    d = edge_mid(q) - edge_mid(p)
    a, b = EDGES[p]
    if (case >> a) & 1:
        a, b = b, a                         # a outside, b inside
    w = corner_xyz(a) - corner_xyz(b)
    s = float(np.dot(np.cross(d, -nrm), w))
    assert abs(s) > 1e-9, (case, loop)
    if s < 0:
        loop = [loop[0]] + loop[:0:-1]
    return loop




def in_face(e1, e2):
    return any(e1 in fe and e2 in fe for fe in FACE_EDGES)


def triangulations(poly):
    """All triangulations of a polygon given as a vertex list (orientation preserved), in a fixed order: the
    triangle on the closing side (poly[0], poly[m], poly[-1]) for m ascending, left part before right part."""
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


def bad_diagonals(loop, tris):
    boundary = set(frozenset((loop[i], loop[(i + 1) % len(loop)])) for i in range(len(loop)))
    n = 0
    for t in tris:
        for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            if frozenset((x, y)) not in boundary and in_face(x, y):
                n += 1
    return n


def triangulate(loop):
    fan = [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
    if bad_diagonals(loop, fan) == 0:
        return fan
    best = None
    for cand in triangulations(loop):
        b = bad_diagonals(loop, cand)
        if b == 0:
            return cand
        if best is None or b < best[0]:
            best = (b, cand)
    return best[1]


def cell_triangles(case, joined_mask):
    """Edge ids of the cell's triangles (flat list) for a case and the joined/separated decision of each face."""
    t = []
    for loop in loops_of(case, joined_mask):
        for tri in triangulate(orient(case, loop)):
            t += list(tri)
    return t


def build():
    """-> (amb masks [256], sub base [256], rows): rows[sub] = (triangle edge list, first-use edge order)."""
    amb, base, rows = [], [], []
    for case in range(256):
        faces = [fi for fi, f in enumerate(FACES) if face_ambiguous(case, f)]
        amb.append(sum(1 << fi for fi in faces))
        base.append(len(rows))
        for bits in range(1 << len(faces)):
            mask = sum(((bits >> q) & 1) << fi for q, fi in enumerate(faces))
            t = cell_triangles(case, mask)
            order = []
            for e in t:
                if e not in order:
                    order.append(e)
            rows.append((t, order))
    return amb, base, rows


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
    amb, base, rows = build()
    geo = edge_geometry()
    nsub = len(rows)
    max_t = max(len(t) for t, _ in rows) // 3
    max_v = max(len(v) for _, v in rows)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "mc_tables.h")
    L = []
    L.append("// GENERATED by rgb-d-pifuhd_b200/tools/gen_mc_tables.py - do not edit.  See that file for the rule.")
    L.append("#pragma once")
    L.append("#define MC_MAX_TRIS %d" % max_t)
    L.append("#define MC_MAX_VERTS %d" % max_v)
    L.append("#define MC_NSUB %d" % nsub)
    L.append("#ifndef MC_TABLE_QUALIFIER\n#define MC_TABLE_QUALIFIER static const\n#endif")
    L.append("// faces (bit f = face f of the cycle list below) whose inside corners sit on a diagonal")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_AMB[256] = {%s};" % ",".join(str(a) for a in amb))
    L.append("// first table row of a case; + the joined/separated bits of its ambiguous faces, ascending face order")
    L.append("MC_TABLE_QUALIFIER unsigned short MC_SUB_BASE[256] = {%s};" % ",".join(str(b) for b in base))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_FACE_CORNERS[6][4] = {%s};" %
             ",".join("{%d,%d,%d,%d}" % f for f in FACES))
    case_of_row = []
    for case in range(256):
        case_of_row += [case] * ((base[case + 1] if case < 255 else nsub) - base[case])
    L.append("// the case a table row belongs to")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_SUB_CASE[MC_NSUB] = {%s};" % ",".join(str(c) for c in case_of_row))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_NTRI[MC_NSUB] = {%s};" % ",".join(str(len(t) // 3) for t, _ in rows))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_NVERT[MC_NSUB] = {%s};" % ",".join(str(len(v)) for _, v in rows))
    body = ["{%s}" % ",".join(str(x) for x in t + [-1] * (3 * max_t - len(t))) for t, _ in rows]
    L.append("MC_TABLE_QUALIFIER signed char MC_TRIS[MC_NSUB][%d] = {\n%s};" % (3 * max_t, ",\n".join(body)))
    body = ["{%s}" % ",".join(str(x) for x in v + [-1] * (max_v - len(v))) for _, v in rows]
    L.append("MC_TABLE_QUALIFIER signed char MC_VERTS[MC_NSUB][%d] = {\n%s};" % (max_v, ",\n".join(body)))
    L.append("// corner offsets along (axis0, axis1, axis2)")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_CORNER[8][3] = {%s};" %
             ",".join("{%d,%d,%d}" % (c[2], c[1], c[0]) for c in CORNERS))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_CORNERS[12][2] = {%s};" %
             ",".join("{%d,%d}" % e for e in EDGES))
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_AXIS[12] = {%s};" % ",".join(str(g[0]) for g in geo))
    L.append("// perpendicular axes on which the edge sits at offset 0 (shared with the previous cell)")
    L.append("MC_TABLE_QUALIFIER unsigned char MC_EDGE_LOWMASK[12] = {%s};" % ",".join(str(g[2]) for g in geo))
    body = ["{%s}" % ",".join(str(shifted_edge(geo, e, m)) for m in range(8)) for e in range(12)]
    L.append("// id of edge e as seen from the cell shifted by -1 on the axes of mask m (-1: impossible)")
    L.append("MC_TABLE_QUALIFIER signed char MC_EDGE_SHIFT[12][8] = {\n%s};" % ",\n".join(body))
    with open(out, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", out, "rows", nsub, "max tris", max_t, "max verts", max_v)
    return amb, base, rows


if __name__ == "__main__":
    main()
