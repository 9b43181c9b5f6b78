"""fgq_b200 — B200-native Gauss quadrature node/weight generation.

Host-side mirror of FastGaussQuadrature.jl's public functions for the per-node generation
path (reference src/FastGaussQuadrature.jl:7-19), over the C ABI of ``libfgq_cuda.so``
(``include/fgq.h``).  Import it as ``import fgq_b200`` (see ``fgq_b200.py`` at the repo root).
"""
from ._lib import DomainError, FGQError, NotImplementedOnDevice, build, lib  # noqa: F401
from . import multi  # noqa: F401  (single-process multi-GPU face of the C ABI)
from .api import (  # noqa: F401
    approx_besselroots,
    empty_pinned,
    besselroots,
    gausschebyshev,
    gausschebyshevt,
    gausschebyshevu,
    gausschebyshevv,
    gausschebyshevw,
    gaussjacobi,
    gaussjacobi_device,
    gausslobatto,
    gaussradau,
    gausslaguerre,
    gausslaguerre_device,
    gausshermite,
    gausshermite_device,
    gausshermite_reduced,
    gausshermite_reduced_device,
    gausslegendre,
    gausslegendre_batch,
    gaussjacobi_batch,
    gausshermite_batch,
    gausslaguerre_batch,
    gausslegendre_device,
    gausslegendre_host_pair,
    gausslegendre_moments_fused,
    gausslegendre_integrate,
    rule_integrate,
    half_index_partition,
    index_partition,
    LegendreShard,
    rule_moments_device,
    measure_peaks,
    launch_count,
)
