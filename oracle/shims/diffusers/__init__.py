"""minimal stand-in for diffusers==0.16.0 (see ../README.md)"""
__version__ = "0.16.0-shim"
