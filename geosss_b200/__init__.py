"""
geosss_b200 -- B200-native batched engine for geosss's geodesic slice-sampling hot path.

Import-compatible names for the path (``import geosss_b200 as gs``)::

    pdf = gs.MixtureModel([gs.VonMisesFisher(80.0 * mu) for mu in mus])
    samples = gs.ShrinkageSphericalSliceSampler(pdf, init_state, seed).sample(n_samples, burnin)

All density evaluation and sampling runs in hand-written sm_100a kernels behind the C ABI of
``libgeosss_cuda.so`` (include/geosss_cuda.h).  There is no CPU fallback.
"""

from .distributions import (  # noqa: F401
    Bingham,
    BinghamFisher,
    CurvedVonMisesFisher,
    Distribution,
    MixtureModel,
    VonMisesFisher,
    random_bingham,
)
from .mcmc import (  # noqa: F401
    IndependenceSampler,
    MetropolisHastings,
    MixtureRWMHIndependenceSampler,
    RejectionSphericalSliceSampler,
    SphericalHMC,
    ShrinkageSphericalSliceSampler,
    determine_burnin,
)
from .pointcloud import (  # noqa: F401
    PointCloud,
    RotationMatrix,
    RotationProjection,
    jacobian_rotation_matrix,
    matrix2quat,
    quat2matrix,
)
from .rand import bfind, sample_bingham, sample_bingham_2d, sample_bingham_3d, sample_vMF  # noqa: F401
from .registration import CoherentPointDrift, GaussianMixtureModel  # noqa: F401
from . import io  # noqa: F401
from .sphere import (  # noqa: F401
    cartesian2polar,
    cartesian2spherical,
    distance,
    givens,
    polar2cartesian,
    spherical2cartesian,
    orthogonal_projection,
    radial_projection,
    sample_sphere,
    sample_subsphere,
    spherical_projection,
)
from .sphere_tesselation import systematic_search, tessellate_rotations  # noqa: F401
from .spherical_curve import SlerpCurve, SphericalCurve, brownian_curve  # noqa: F401
from .utils import IAT, SamplerLauncher, acf, acf_fft, colors, counter, n_eff, take_time  # noqa: F401

__version__ = "0.1.0"
