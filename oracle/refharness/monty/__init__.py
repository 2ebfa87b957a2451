"""Test-infrastructure stand-in for monty (only ``monty.io.zopen`` is imported by the reference)."""
