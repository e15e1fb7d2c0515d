"""Stand-in package for the missing pip dependency `diffusers` (oracle/shims: used only to import the unmodified reference when minting golden vectors)."""
