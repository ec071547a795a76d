"""Telescope implementations (mirrors ``drift.telescope``)."""
