"""Weighted point clouds and rotation helpers for the registration target (geosss/pointcloud.py)."""

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


class RotationProjection(PointCloud):
    """3-D cloud that is rotated and then projected onto the xy-plane (parallel-beam geometry,
    pointcloud.py:273-293); the source of a 3D->2D registration."""

    def transform_positions(self, rotation, translation=None):
        R = quat2matrix(rotation) if np.shape(rotation)[-1] == 4 else np.asarray(rotation)
        shift = np.zeros(self.dim - 1) if translation is None else translation
        return self.positions @ R[:-1].T + shift


class RotationMatrix:
    """2x2 / 3x3 rotation matrices from an angle / intrinsic z-y-z Euler angles (pointcloud.py:9-98)."""

    def __init__(self, degree=False):
        self.degree = degree
        self.rot_mat = None

    def create(self, angle):
        angle = np.asarray(angle, dtype=float)
        if self.degree:
            angle = np.deg2rad(angle)
        if angle.size == 1:
            return self.rotation2d(float(angle.reshape(-1)[0]))
        if angle.size == 3:
            return self.rotation3d(angle)
        raise ValueError("Angle vector should be of size 1 or 3")

    def rotation2d(self, angle):
        c, s = np.cos(angle), np.sin(angle)
        self.rot_mat = np.array([[c, -s], [s, c]])
        return self.rot_mat

    def rotation3d(self, euler_angles):
        def about_z(t):
            return np.array([[np.cos(t), -np.sin(t), 0.0], [np.sin(t), np.cos(t), 0.0], [0.0, 0.0, 1.0]])

        alpha, beta, gamma = euler_angles
        about_y = np.array([[np.cos(beta), 0.0, np.sin(beta)], [0.0, 1.0, 0.0], [-np.sin(beta), 0.0, np.cos(beta)]])
        self.rot_mat = about_z(alpha) @ about_y @ about_z(gamma)
        return self.rot_mat


def jacobian_rotation_matrix(q):
    """MINUS dR/dq, shape (3, 3, 4), for R(q) the rotation of q / |q|, q = (x, y, z, w) not necessarily
    normalised (pointcloud.py:135-203; the sign is upstream's).  Quotient rule on R = N / r,
    r = |q|^2 + 1e-300, the same form the device gradient uses (csrc/targets.cuh: cpd_grad_finish)."""
    x, y, z, w = np.asarray(q, dtype=float)
    r = w * w + x * x + y * y + z * z + 1e-300
    N = r * quat2matrix_unnormalised(x, y, z, w)
    dN = 2 * np.stack([np.array([[x, y, z], [y, -x, -w], [z, w, -x]]), np.array([[-y, x, w], [x, y, z], [-w, z, -y]]),
                       np.array([[-z, -w, x], [w, -z, y], [x, y, z]]), np.array([[w, -z, y], [z, w, -x], [-y, x, w]])], axis=-1)
    dr = 2 * np.array([x, y, z, w])
    return -(dN / r - N[..., None] * dr / r**2)


def quat2matrix_unnormalised(x, y, z, w):
    """N(q) / r: the rotation written with the quaternion's squared norm r left explicit."""
    r = w * w + x * x + y * y + z * z + 1e-300
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]]) / r
