"""CPU ORACLE (test infrastructure - NOT a product path): the component `meshcleaning` keeps
(`reconstruction.py:325-344`), restated with scipy.

trimesh is third-party and absent from this image ("parity unpinned" for its exact behaviour); what is restated, from
the documented behaviour of `Trimesh.split()` / `graph.split` / `graph.connected_components`:
  * face adjacency = pairs of faces that share an edge used by exactly two faces;
  * components of that graph, in the order of their lowest face; `split()` defaults to only_watertight=True: a
    component is kept when every edge of the SUBMESH is used by exactly two of its faces, and it has >= 4 faces;
  * meshcleaning: out = cc[0]; replaced by a later component only when its extent along axis 0 is strictly greater.
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def largest_component(verts, faces, colors=None, only_watertight=True):
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    nf = len(faces)
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0), 1)
    owner = np.tile(np.arange(nf), 3)
    key = e[:, 0] * (len(verts) + 1) + e[:, 1]
    order = np.argsort(key, kind="stable")
    key, owner = key[order], owner[order]
    uniq, start, cnt = np.unique(key, return_index=True, return_counts=True)
    two = start[cnt == 2]
    a, b = owner[two], owner[two + 1]
    graph = coo_matrix((np.ones(len(a)), (a, b)), shape=(nf, nf))
    _, label = connected_components(graph, directed=False)
    # order of components = order of their lowest face
    firsts = {}
    for f, l in enumerate(label):
        firsts.setdefault(l, f)
    comps = sorted(firsts, key=lambda l: firsts[l])
    best, best_h = None, None
    for l in comps:
        fs = np.flatnonzero(label == l)
        if only_watertight:
            if len(fs) < 4:
                continue
            sub = np.sort(np.concatenate([faces[fs][:, [0, 1]], faces[fs][:, [1, 2]], faces[fs][:, [2, 0]]], 0), 1)
            _, c = np.unique(sub[:, 0] * (len(verts) + 1) + sub[:, 1], return_counts=True)
            if not (c == 2).all():
                continue
        x = verts[np.unique(faces[fs]), 0]
        h = x.max() - x.min()
        if best is None or h > best_h:
            best, best_h = fs, h
    if best is None:
        raise IndexError("no component")            # trimesh: cc[0] on an empty list
    keep_v = np.unique(faces[best])
    newid = -np.ones(len(verts), dtype=np.int64)
    newid[keep_v] = np.arange(len(keep_v))
    out_f = newid[faces[best]].astype(np.int32)
    return verts[keep_v], out_f, (np.asarray(colors)[keep_v] if colors is not None else None)
