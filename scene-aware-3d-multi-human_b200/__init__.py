"""B200-native scene-constrained multi-person SMPL optimisation loop.

Drop-in for the hot path of dluvizon/scene-aware-3d-multi-human (``mhmocap/optimizer.py`` driving ``smpl.py``,
``transforms.py``, ``losses.py``): Python host code over ``libmhopt.so`` (hand-written sm_100a CUDA kernels behind the
C ABI of ``include/mhopt.h``).  Importing the package loads the library and fails loudly if it is missing.
"""
from . import _lib                                         # noqa: F401  (raises when libmhopt.so is absent)
from .optimizer import SMPLDepthSequenceOptimizer, SMPLOptimizerBase       # noqa: F401

__all__ = ['SMPLDepthSequenceOptimizer', 'SMPLOptimizerBase']
