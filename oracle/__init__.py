"""Parity oracle for the mft6.py log-posterior path.  TEST INFRASTRUCTURE ONLY.

Importable only from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs -- never from ``mcmc_spec_b200`` (the product).  PARITY UNPINNED for the
restated third-party numerics: see oracle/thirdparty.py.
"""
