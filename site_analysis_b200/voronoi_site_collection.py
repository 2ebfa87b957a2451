"""Reference module name kept for drop-in imports (``site_analysis.voronoi_site_collection``)."""
from .site_collection import VoronoiSiteCollection  # noqa: F401
