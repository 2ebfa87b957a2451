"""site_analysis_b200 -- B200 (sm_100a) backend for the per-frame site-assignment hot path of
bjmorgan/site-analysis, behind the reference's own Python API.

    from site_analysis_b200 import TrajectoryBuilder        # instead of: from site_analysis import ...

The CUDA backend (``libsab200.so``, C ABI in ``include/sab200.h``) is required; there is no CPU
fallback.
"""
from .atom import Atom, atoms_from_indices, atoms_from_species_string, atoms_from_structure  # noqa: F401
from .builders import (TrajectoryBuilder, create_trajectory_with_dynamic_voronoi_sites,  # noqa: F401
                       create_trajectory_with_polyhedral_sites, create_trajectory_with_spherical_sites,
                       create_trajectory_with_voronoi_sites)
from .dynamic_voronoi_site import DynamicVoronoiSite  # noqa: F401
from .engine import AssignmentEngine, GuardError  # noqa: F401
from .polyhedral_site import PolyhedralSite  # noqa: F401
from .site import Site  # noqa: F401
from .site_collection import (DynamicVoronoiSiteCollection, PolyhedralSiteCollection,  # noqa: F401
                              SiteCollection, SphericalSiteCollection, VoronoiSiteCollection)
from .spherical_site import SphericalSite  # noqa: F401
from .structure import SimpleStructure  # noqa: F401
from .trajectory import Trajectory  # noqa: F401
from .transition_table import TransitionTable  # noqa: F401
from .voronoi_site import VoronoiSite  # noqa: F401
from .folds import fold_rows  # noqa: F401

__version__ = "0.1.0"
