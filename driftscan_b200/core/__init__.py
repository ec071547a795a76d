"""Telescope model, beam-transfer products and the product manager
(mirrors the ``drift.core`` package layout of the reference)."""
