"""Reference module name kept for drop-in imports (``site_analysis.polyhedral_site_collection``)."""
from .site_collection import PolyhedralSiteCollection  # noqa: F401
