"""B200-native scene-constrained multi-person SMPL optimisation loop.

Drop-in for the hot path of dluvizon/scene-aware-3d-multi-human (``mhmocap/optimizer.py`` driving ``smpl.py``,
``transforms.py``, ``losses.py``): Python host code over ``libmhopt.so`` (hand-written sm_100a CUDA kernels behind the
C ABI of ``include/mhopt.h``).  Importing the package loads the library and fails loudly if it is missing.
"""
from . import _lib                                         # noqa: F401  (raises when libmhopt.so is absent)
from .optimizer import SMPLDepthSequenceOptimizer, SMPLOptimizerBase       # noqa: F401
from . import evaluation                                   # noqa: F401  (mhmocap/evaluate.py mirror; SMPL joints on the device)



def install_as_mhmocap_optimizer():
    """Register this package's optimiser module as ``mhmocap.optimizer`` so that the reference's ``mhmocap.predict``
    (``predict.py:12``: ``from .optimizer import SMPLDepthSequenceOptimizer``) binds the B200 implementation.  Must run
    before ``mhmocap.predict`` is imported; also provides an empty ``matplotlib.pyplot`` when matplotlib is absent
    (``predict.py:6``; plots are only drawn with ``save_visualizations``)."""
    import sys
    import types
    from . import optimizer
    sys.modules['mhmocap.optimizer'] = optimizer
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType('matplotlib')
        mpl.pyplot = types.ModuleType('matplotlib.pyplot')
        sys.modules.setdefault('matplotlib', mpl)
        sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
    return optimizer


__all__ = ['SMPLDepthSequenceOptimizer', 'SMPLOptimizerBase', 'install_as_mhmocap_optimizer', 'evaluation']
