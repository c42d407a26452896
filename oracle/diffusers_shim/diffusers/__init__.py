"""Test-only stand-in for diffusers==0.31.0 (see ../README.md)."""
__version__ = "0.31.0+shim"
