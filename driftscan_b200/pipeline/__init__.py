"""Timestream side of the path: simulated visibilities and their m-mode transform."""
