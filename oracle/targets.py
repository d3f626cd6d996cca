"""
Target log-densities on the sphere (oracle; test infrastructure only).

Float64 numpy/scipy restatement of the ``log_prob`` methods on the geosss hot
path.  Each class cites the reference lines it follows.  The classes are plain
parameter holders: ``log_prob(x)`` accepts one point ``(d,)`` or rows ``(N, d)``.
"""

import numpy as np
from scipy.special import i0, logsumexp
from scipy.spatial import KDTree
from scipy.spatial.transform import Rotation

from . import sphere


class VMFMixture:
    """Mixture of von Mises-Fisher "densities".

    Component: ``x @ mu - log(2 pi) - log(i0(|mu|))`` (geosss/distributions.py:156-157;
    the normaliser uses i0 whatever d is and overflows to inf for |mu| >~ 713).
    Mixture: per-component values are stacked, ``log w`` added and reduced with
    scipy's logsumexp (distributions.py:218-221); weights are normalised to sum 1
    and default to uniform (distributions.py:211-216).
    """

    def __init__(self, mus, weights=None):
        self.mus = np.array(mus, dtype=float)
        if self.mus.ndim == 1:
            self.mus = self.mus[None, :]
        K = len(self.mus)
        w = np.ones(K) if weights is None else np.array(weights, dtype=float)
        self.weights = w / w.sum()

    @property
    def d(self):
        return self.mus.shape[1]

    def component_log_prob(self, k, x):
        mu = self.mus[k]
        return x @ mu - np.log(2 * np.pi) - np.log(i0(np.linalg.norm(mu)))

    def log_prob(self, x):
        x = np.asarray(x, dtype=float)
        p = np.transpose([self.component_log_prob(k, x) for k in range(len(self.mus))])
        p = p + np.log(self.weights)
        return logsumexp(p, axis=-1)

    def gradient(self, x):
        """sum_k exp(p_k) mu_k / exp(logsumexp(p)) (distributions.py:223-227; component gradient = mu,
        :159-160).  The un-shifted exp underflows to 0/0 far from every mode, as upstream."""
        x = np.asarray(x, dtype=float)
        p = np.array([self.component_log_prob(k, x) for k in range(len(self.mus))]) + np.log(self.weights)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sum(np.exp(p) * self.mus.T, axis=1) / np.exp(logsumexp(p, axis=-1))

    def log_consts(self):
        """log w_k - log 2pi - log i0(kappa_k): what the device kernels precompute."""
        with np.errstate(over="ignore"):
            kap = np.linalg.norm(self.mus, axis=1)
            return np.log(self.weights) - np.log(2 * np.pi) - np.log(i0(kap))


class SingleVMF(VMFMixture):
    """A lone VonMisesFisher (distributions.py:117-160): no logsumexp, no weights."""

    def __init__(self, mu):
        super().__init__(np.asarray(mu, dtype=float)[None, :])

    def log_prob(self, x):
        return self.component_log_prob(0, np.asarray(x, dtype=float))

    def gradient(self, x):
        return self.mus[0]


class Bingham:
    """log p(x) = x^T A x [+ x.b]  (distributions.py:78-86; Fisher term :113-114)."""

    def __init__(self, A, b=None):
        A = np.asarray(A, dtype=float)
        assert np.allclose(A, A.T)  # distributions.py:68
        self.A = A
        self.b = None if b is None else np.asarray(b, dtype=float)
        lam, vec = np.linalg.eigh(A)  # ascending (distributions.py:73-76)
        self.v, self.U = lam[::-1], vec[:, ::-1]

    @property
    def d(self):
        return len(self.A)

    @property
    def mode(self):
        return self.U[:, 0]

    def gradient(self, x):
        """2 A x (distributions.py:88-89).  BinghamFisher inherits it unchanged upstream, so the
        Fisher term b does NOT enter the gradient; kept."""
        return 2 * self.A @ np.asarray(x, dtype=float)

    def log_prob(self, x):
        x = np.asarray(x, dtype=float)
        assert x.ndim in (1, 2) and x.shape[-1] == self.d  # distributions.py:84
        out = np.sum((x @ self.A) * x, axis=-1)
        if self.b is not None:
            out = out + x @ self.b
        return out


def random_bingham_matrix(d=2, vmax=None, vmin=None, eigensystem=False, seed=None):
    """Precision matrix of ``random_bingham`` (distributions.py:230-258)."""
    gen = np.random.default_rng(seed)
    G = gen.standard_normal((d, d))
    S = G.T @ G
    lam, vec = np.linalg.eigh(S)
    if vmin is not None:
        lam += vmin - lam.min()
    if vmax is not None:
        lam *= vmax / lam.max()
    if eigensystem:
        vec = np.eye(d)
    return (vec * lam) @ vec.T


def brownian_knots(n_points=100, dimension=6, step_size=0.05, seed=1234):
    """Knots of ``brownian_curve`` (geosss/spherical_curve.py:105-129)."""
    gen = np.random.default_rng(seed)
    pts = np.zeros((n_points, dimension))
    pts[0] = sphere.unit(gen.standard_normal(dimension))
    for i in range(1, n_points):
        pts[i] = sphere.unit(pts[i - 1] + gen.normal(size=dimension) * step_size)
    return pts


class CurvedVMF:
    """kappa * x . (nearest point on the piecewise-slerp curve through the knots).

    distributions.py:272-275 -> SlerpCurve.find_nearest (spherical_curve.py:95-102)
    -> distance_slerp (spherical_curve.py:10-32).  Segment loop, ``sin(theta)+1e-10``
    and first-index argmin are kept literally.
    """

    def __init__(self, knots, kappa=100.0):
        self.knots = np.asarray(knots, dtype=float)
        self.kappa = float(kappa)

    @property
    def d(self):
        return self.knots.shape[-1]

    @staticmethod
    def _nearest_on_segment(x, a, b):
        ang = sphere.distance(a, b)
        t = np.arctan2(b @ x - a @ x * np.cos(ang), a @ x * np.sin(ang))
        t = np.clip(t, 0.0, ang)
        y = (np.sin(ang - t) * a + np.sin(t) * b) / (np.sin(ang) + 1e-10)
        return sphere.distance(x, y), y

    def nearest(self, x):
        best, best_y = None, None
        for a, b in zip(self.knots, self.knots[1:]):
            dist, y = self._nearest_on_segment(x, a, b)
            if best is None or dist < best:  # np.argmin keeps the first minimum
                best, best_y = dist, y
        return best_y

    def log_prob(self, x):
        x = np.asarray(x, dtype=float)
        if x.ndim == 1:
            return self.kappa * (x @ self.nearest(x))
        return np.array([self.log_prob(row) for row in x])

    def gradient(self, x):
        """kappa * nearest curve point (distributions.py:277-278)."""
        return self.kappa * self.nearest(np.asarray(x, dtype=float))


def quat_to_matrix(q):
    """scipy Rotation.from_quat(q).as_matrix(): scalar-last, q normalised first
    (geosss/pointcloud.py:101-115)."""
    return Rotation.from_quat(q).as_matrix()


def quat_to_matrix_closed_form(q):
    """Closed form of the same map; what the device kernels evaluate."""
    x, y, z, w = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def rotation_jacobian(q):
    """MINUS dR/dq for R(q) = the rotation of q / |q|, scalar-last, shape (3, 3, 4)
    (pointcloud.py:135-203; the sign flip is upstream's and cancels the one in ``grad_R``).
    Written as the quotient rule on R = N / r, N the homogeneous quadratic form, r = |q|^2 + 1e-300."""
    x, y, z, w = np.asarray(q, dtype=float)
    r = w**2 + x**2 + y**2 + z**2 + 1e-300
    N = np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )
    dN = np.zeros((3, 3, 4))
    dN[..., 0] = 2 * np.array([[x, y, z], [y, -x, -w], [z, w, -x]])
    dN[..., 1] = 2 * np.array([[-y, x, w], [x, y, z], [-w, z, -y]])
    dN[..., 2] = 2 * np.array([[-z, -w, x], [w, -z, y], [x, y, z]])
    dN[..., 3] = 2 * np.array([[w, -z, y], [z, w, -x], [-y, x, w]])
    dr = 2 * np.array([x, y, z, w])
    return -(dN / r - N[..., None] * dr / r**2)


class CPD:
    """Coherent-point-drift log-likelihood of a rigid rotation (unit quaternion).

    registration.py:47-53 (quaternion -> R, times beta), :101-109 (kNN query of
    the rotated source against each target point, KDTree rebuilt per call),
    :197-259 (Gaussian terms + uniform outlier column, logsumexp per target row,
    weighted sum over rows).  ``k <= 0`` or ``k >= n_source`` sums all source points.
    """

    def __init__(self, target, source, sigma=1.0, k=20, beta=1.0, omega=0.0,
                 target_weights=None, source_weights=None, translation=None):
        # a 2-D target scores the rotated source projected onto xy (RotationProjection,
        # pointcloud.py:273-293); translation: fixed shift after the rotation (pointcloud.py:260-264)
        self.translation = None if translation is None else np.asarray(translation, dtype=float)
        self.target = np.asarray(target, dtype=float)
        self.source = np.asarray(source, dtype=float)
        self.w_t = (np.ones(len(self.target)) if target_weights is None
                    else np.asarray(target_weights, dtype=float))
        self.w_s = (np.ones(len(self.source)) if source_weights is None
                    else np.asarray(source_weights, dtype=float))
        self.sigma, self.k = float(sigma), int(k)
        self.beta, self.omega = float(beta), float(omega)
        self.d = 4

    @property
    def log_outlier(self):
        # registration.py:205-209, 237-238
        with np.errstate(divide="ignore"):
            return np.log(self.omega + 1e-308) - np.sum(np.log(np.ptp(self.target, 0)))

    @property
    def log_constant(self):
        # registration.py:211-215
        D = self.target.shape[1]
        with np.errstate(divide="ignore"):
            return np.log(1 - self.omega) - D / 2 * np.log(2 * np.pi * self.sigma**2)

    def _neighbours(self, moved):
        k = self.k if 0 < self.k < len(moved) else len(moved)
        dist, ind = KDTree(moved).query(self.target, k=k)
        if k == 1:
            dist, ind = dist[:, None], ind[:, None]
        return dist, ind

    def _move(self, R):
        D = self.target.shape[1]
        moved = self.source @ R[:D].T
        return moved if self.translation is None else moved + self.translation

    def log_prob_R(self, R):
        moved = self._move(R)
        dist, ind = self._neighbours(moved)
        log_phi = -0.5 * dist**2 / self.sigma**2
        log_p = np.log(self.w_s[ind]) + log_phi + self.log_constant
        col = np.full((len(self.target), 1), self.log_outlier)
        rows = logsumexp(np.hstack((log_p, col)), axis=1)
        return rows @ self.w_t

    def log_prob(self, q):
        q = np.asarray(q, dtype=float)
        if q.ndim == 2 and q.shape != (3, 3):
            return np.array([self.log_prob(row) for row in q])
        R = q if q.shape == (3, 3) else quat_to_matrix(q)
        return self.beta * self.log_prob_R(R)

    def _row_log_norm(self, log_p):
        """log p_l that the responsibilities are taken against (registration.py:217-247: outlier
        column included)."""
        col = np.full((len(self.target), 1), self.log_outlier)
        return logsumexp(np.hstack((log_p, col)), axis=1)

    def _log_terms(self, dist, ind):
        return np.log(self.w_s[ind]) + (-0.5 * dist**2 / self.sigma**2) + self.log_constant

    def grad_R(self, q):
        """d log p / d R as a 3x3 array (registration.py:261-293): responsibilities
        gamma = exp(clip(log p_lk - log p_l, -20, 0)), residuals y_l - R x_k, outer products with
        the un-rotated source points."""
        moved = self._move(quat_to_matrix(q))
        dist, ind = self._neighbours(moved)
        log_p = self._log_terms(dist, ind)
        gamma = np.exp(np.clip(log_p - self._row_log_norm(log_p)[:, None], -20, 0))
        resid = self.target[:, None, :] - moved[ind]
        pad = self.source.shape[1] - self.target.shape[1]  # registration.py:277-283: zero-pad to the source dimension
        resid = np.pad(resid, ((0, 0), (0, 0), (0, pad)))
        coeff = (self.w_t[:, None] * gamma) / self.sigma**2
        return -np.einsum("lk,lkj,lki->ji", coeff, resid, self.source[ind])

    def gradient(self, q):
        """Chain rule through the (sign-flipped) Jacobian of R(q) (registration.py:55-59,
        pointcloud.py:135-203)."""
        q = np.asarray(q, dtype=float)
        g = (self.beta * self.grad_R(q)).ravel()
        return rotation_jacobian(q).reshape(9, 4).T @ g


class GMM(CPD):
    """Parent model without the outlier column / (1-omega) factor
    (registration.py:111-119)."""

    def __init__(self, target, source, sigma=1.0, k=20, beta=1.0,
                 target_weights=None, source_weights=None, translation=None):
        super().__init__(target, source, sigma, k, beta, 0.0, target_weights, source_weights, translation)

    def log_prob_R(self, R):
        moved = self._move(R)
        dist, ind = self._neighbours(moved)
        return logsumexp(self._log_terms(dist, ind), axis=1) @ self.w_t

    def _log_terms(self, dist, ind):
        D = self.target.shape[1]
        log_phi = -0.5 * dist**2 / self.sigma**2
        return log_phi - D / 2 * np.log(2 * np.pi * self.sigma**2) + np.log(self.w_s[ind])

    def _row_log_norm(self, log_p):
        # registration.py:134-135: no outlier column in the parent model
        return logsumexp(log_p, axis=1)
