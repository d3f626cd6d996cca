"""Weighted point clouds for the registration target (geosss/pointcloud.py:101-115, 206-270)."""

import numpy as np


def quat2matrix(q):
    """Unit quaternion (x, y, z, w), scalar last, normalised first -> 3x3 rotation matrix
    (what scipy's Rotation.from_quat(q).as_matrix() returns, pointcloud.py:101-115)."""
    x, y, z, w = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def matrix2quat(R):
    """3x3 rotation matrix -> unit quaternion, scalar last (Shepperd's method)."""
    R = np.asarray(R, dtype=float)
    tr = np.trace(R)
    cand = np.array([R[0, 0], R[1, 1], R[2, 2], tr])
    i = int(np.argmax(cand))
    if i == 3:
        q = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], 1 + tr])
    else:
        j, k = (i + 1) % 3, (i + 2) % 3
        q = np.empty(4)
        q[i] = 1 + 2 * R[i, i] - tr
        q[j] = R[j, i] + R[i, j]
        q[k] = R[k, i] + R[i, k]
        q[3] = R[k, j] - R[j, k]
    return q / np.linalg.norm(q)


class PointCloud:
    """Positions (n, 2|3) with optional per-point weights (pointcloud.py:206-270)."""

    def __init__(self, positions, weights=None):
        if np.ndim(positions) != 2:
            raise ValueError("Expected rank-2 array")
        if np.shape(positions)[1] not in (2, 3):
            raise ValueError("Expected 2d or 3d point cloud")
        self.positions = np.array(positions, dtype=float)
        self.weights = np.ones(len(self.positions)) if weights is None else np.array(weights, dtype=float)

    @property
    def dim(self):
        return self.positions.shape[1]

    @property
    def size(self):
        return self.positions.shape[0]

    @property
    def center_of_mass(self):
        return self.weights @ self.positions / self.weights.sum()

    def transform_positions(self, rotation, translation=None):
        R = quat2matrix(rotation) if np.shape(rotation)[-1] == 4 else np.asarray(rotation)
        shift = np.zeros(len(R)) if translation is None else translation
        return self.positions @ R.T + shift

    def transform(self, rotation, translation=None):
        self.positions = self.transform_positions(rotation, translation)
