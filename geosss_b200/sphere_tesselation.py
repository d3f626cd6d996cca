"""
Rotation grids from the 600-cell for exhaustive registration search
(geosss/sphere_tesselation.py:36-201; used by scripts/protein_reg3d3d.py:629-648).

The reference builds the grid on the host and then loops ``log_prob`` over it in Python (~1.5 ms per
rotation).  Here the grid construction stays a host-side set-up step (numpy, vectorised over all
tetrahedra) and the evaluation is one batched GPU call: ``systematic_search(pdf, n_discretize)``.
"""

import itertools

import numpy as np

from .pointcloud import matrix2quat, quat2matrix

_EVEN_PERMS = ((0, 1, 2, 3), (0, 2, 3, 1), (0, 3, 1, 2), (1, 0, 3, 2), (1, 2, 0, 3), (1, 3, 2, 0),
               (2, 0, 1, 3), (2, 1, 3, 0), (2, 3, 0, 1), (3, 0, 2, 1), (3, 1, 0, 2), (3, 2, 1, 0))


def c600_vertices(upper_sphere=True):
    """The 120 vertices of the 600-cell (unit quaternions of the binary icosahedral group), optionally only
    those within 120 degrees of (1,0,0,0)  (sphere_tesselation.py:64-98)."""
    phi = (1 + np.sqrt(5)) / 2
    pts = [np.array(s, dtype=float) for s in itertools.product((-1, 1), repeat=4)]
    for j in range(4):
        for s in (-2.0, 2.0):
            e = np.zeros(4)
            e[j] = s
            pts.append(e)
    for perm in _EVEN_PERMS:
        for a, b, c in itertools.product((-1, 1), repeat=3):
            v = np.zeros(4)
            v[perm[0]], v[perm[1]], v[perm[2]] = a * phi, b, c / phi
            pts.append(v)
    nodes = 0.5 * np.array(pts)
    if upper_sphere:
        nodes = nodes[np.arccos(nodes[:, 0]) <= np.deg2rad(120)]
    return nodes


def c600_tetrahedra(vertices):
    """All 4-cliques of the 36-degree neighbour graph, in lexicographic vertex order like the reference's
    itertools.combinations scan (sphere_tesselation.py:100-117), found by growing cliques along edges."""
    angles = np.round(np.rad2deg(np.arccos(np.clip(vertices @ vertices.T, -1, 1))))
    adj = angles == 36.0
    n = len(vertices)
    cells = []
    for i in range(n):
        ni = [j for j in range(i + 1, n) if adj[i, j]]
        for a, j in enumerate(ni):
            nij = [k for k in ni[a + 1:] if adj[j, k]]
            for b, k in enumerate(nij):
                for m in nij[b + 1:]:
                    if adj[k, m]:
                        cells.append((i, j, k, m))
    return vertices[np.array(cells)]


def split_tetrahedra(tets):
    """Every tetrahedron (T, 4, 4) -> its eight children (8T, 4, 4), edge midpoints pushed onto S^3, the
    inner octahedron cut along its shortest diagonal  (sphere_tesselation.py:120-171)."""
    v = tets
    edges = ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))
    mid = np.stack([v[:, i] + v[:, j] for i, j in edges], axis=1)
    # The cells of the 600-cell are regular, so the three diagonals tie up to rounding and the cut is decided
    # by the last bit.  Norms and dot products are therefore taken vector by vector with the same numpy calls
    # the reference makes (1-D norm, 1-D @), which reproduces its grid exactly.
    flat = mid.reshape(-1, 4)
    flat /= np.array([np.linalg.norm(row) for row in flat])[:, None]
    dots = np.array([[m[0] @ m[5], m[2] @ m[4], m[3] @ m[1]] for m in mid])
    which = np.argmax(dots, axis=1)
    V = lambda i: v[:, i]  # noqa: E731
    N = lambda i: mid[:, i]  # noqa: E731
    corner = [(V(0), N(0), N(2), N(3)), (V(1), N(0), N(1), N(4)), (V(2), N(1), N(2), N(5)), (V(3), N(3), N(4), N(5))]
    inner = {
        0: [(0, 5, 3, 2), (0, 5, 4, 3), (5, 0, 1, 4), (5, 0, 1, 2)],
        1: [(0, 4, 3, 2), (0, 1, 4, 2), (5, 2, 1, 4), (5, 3, 2, 4)],
        2: [(3, 1, 0, 2), (3, 1, 4, 0), (3, 1, 5, 4), (1, 3, 2, 5)],
    }
    out = np.empty((len(v), 8, 4, 4))
    for c, quad in enumerate(corner):
        out[:, c] = np.stack(quad, axis=1)
    for case, quads in inner.items():
        sel = which == case
        for c, quad in enumerate(quads):
            out[sel, 4 + c] = np.stack([mid[sel, q] for q in quad], axis=1)
    return out.reshape(-1, 4, 4)


def tessellate_rotations(n_discretize=2):
    """Unit quaternions (x, y, z, w) covering rotation space: centroids of the 600-cell's tetrahedra after
    ``n_discretize`` subdivisions, duplicates (as rotations, to 10 decimals) removed  (sphere_tesselation.py:174-201)."""
    tets = c600_tetrahedra(c600_vertices())
    for _ in range(n_discretize):
        tets = split_tetrahedra(tets)
    q = tets.mean(axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rot = np.array([quat2matrix(row) for row in q]).reshape(len(q), -1)
    rot = np.unique(np.round(rot, decimals=10), axis=0).reshape(-1, 3, 3)
    return np.array([matrix2quat(R) for R in rot])


def systematic_search(pdf, n_discretize=2, quaternions=None):
    """Evaluate ``pdf.log_prob`` on the whole grid in one batched GPU call and return
    (best quaternion, best log-probability, quaternions, log-probabilities)."""
    quats = tessellate_rotations(n_discretize) if quaternions is None else np.asarray(quaternions, dtype=float)
    logp = np.asarray(pdf.log_prob(quats))
    best = int(np.argmax(logp))
    return quats[best], float(logp[best]), quats, logp
