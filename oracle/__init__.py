"""CPU fp64 oracle for the driftscan beam-transfer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``driftscan_b200/`` may import this
package: it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function cites the reference file:line (relative to the upstream
``radiocosmology/driftscan`` tree) that it restates.  Parts of the path live
in third-party packages that are absent from the reference tree (``cora``,
``healpy``/libsharp, ``caput``): those are restated from their published
definitions and flagged ``EXTERNAL``.

Parity status
-------------
* in-tree arithmetic (fringe, Stokes maps, baseline bookkeeping, noise, SVD
  chain, projections, blockla): PINNED against outputs of the reference's own
  code executed under dependency stubs (``tests/golden/make_golden.py``).
* spherical-harmonic transform (healpy.map2alm / libsharp) and HEALPix pixel
  geometry: **parity unpinned** -- neither healpy nor the reference's golden
  products (downloaded tarball, tests/test_functional.py:121-127) are
  available offline.  The restatement is pinned instead by analytic
  known-answer tests (tests/test_oracle_sht.py).
"""
