"""Synthetic scene of BASELINE.json configs[3] ("C4"): the five Cornell walls plus two tessellated icospheres
(subdivision 9 -> 2 x 5,242,880 = 10,485,760 triangles), SURVEY.md 8d.  Harness code (numpy), not part of the hot path.

  unit-sphere vertices scaled 0.45, centres (-0.45, -0.55, -3.2) and (0.5, -0.55, -2.8); vertex coordinates rounded to 6
  decimals (what an OBJ written with %.6f would hold); vertex normal = unit position; material 'teapot' (index 3).
"""
import numpy as np

from .api import TRI_DTYPE


def icosphere(subdiv):
    """Vertices (unit sphere) and faces of an icosahedron subdivided `subdiv` times (each triangle -> 4)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e.sort(axis=1)
        key = e[:, 0] * (v.shape[0] + 1) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (v.shape[0] + 1), uniq % (v.shape[0] + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid])
        n = f.shape[0]
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return v, f


def cornell_walls(cornellbox_tris):
    """The wall triangles of geometry/cornellbox.obj: the file lists the two boxes first (objects shortBox, longBox,
    geometry/cornellbox.obj:4-79 = triangles 0-19), then ceiling, leftWall, backWall, rightWall, floor (:80-187 = 20-59)."""
    assert cornellbox_tris.size == 60
    return cornellbox_tris[20:]


def synthetic_c4(cornellbox_tris, subdiv=9, material=3):
    """Triangles (TRI_DTYPE) of the C4 scene at the given subdivision level."""
    parts = [cornell_walls(cornellbox_tris)]
    v, f = icosphere(subdiv)
    for centre in ((-0.45, -0.55, -3.2), (0.5, -0.55, -2.8)):
        pos = np.round(v * 0.45 + np.asarray(centre), 6).astype(np.float32)
        nrm = np.round(v, 6).astype(np.float32)
        t = np.zeros(f.shape[0], TRI_DTYPE)
        for k, name in enumerate(("v1", "v2", "v3")):
            t[name][:, :3] = pos[f[:, k]]; t[name][:, 3] = 1.0
        for k, name in enumerate(("vn1", "vn2", "vn3")):
            t[name][:, :3] = nrm[f[:, k]]
        t["matID"] = material
        parts.append(t)
    return np.concatenate(parts)
