"""ORACLE package -- test infrastructure, not product code.

CPU restatement of the reference (hyperbolic-optics 0.3.0) reflection path used
as the parity checker.  Importers allowed: ``tests/``, ``__graft_entry__.smoke()``,
``bench.py`` (``cpu_baseline`` leg and ``--impl reference``).  The product package
``hyperbolic_optics_b200`` must never import from here.
"""
