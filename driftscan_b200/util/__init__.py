"""Small host-side helpers (naming patterns, sky geometry, HEALPix pixel grid)."""
