"""Test-infrastructure stand-in for pymatgen (see pymatgen/core/__init__.py)."""
