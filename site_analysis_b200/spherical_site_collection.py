"""Reference module name kept for drop-in imports (``site_analysis.spherical_site_collection``)."""
from .site_collection import SphericalSiteCollection  # noqa: F401
