"""Reference module name kept for drop-in imports (``site_analysis.dynamic_voronoi_site_collection``)."""
from .site_collection import DynamicVoronoiSiteCollection  # noqa: F401
