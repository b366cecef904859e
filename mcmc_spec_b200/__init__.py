"""mcmc_spec_b200: B200-native log-posterior evaluation for unresolved binary stars.

A from-scratch implementation of the per-walker ``logposterior`` path of
kendallsullivan/mcmc_spec (``mft6.py``), behind the same parameter-file interface and the same
``lnprob(theta) -> float`` contract, computed by hand-written sm_100a CUDA kernels through a
C-ABI shared library (``libmcmcspec.so``; declarations in ``include/mcmcspec.h``).
"""
from .problem import Problem  # noqa: F401

__all__ = ["Problem"]
