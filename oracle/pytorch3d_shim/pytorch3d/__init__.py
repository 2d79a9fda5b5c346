"""Stand-in ``pytorch3d`` package (TEST INFRASTRUCTURE ONLY).

Exposes just the five renderer symbols and ``Meshes`` that the UNMODIFIED
reference optimiser imports (``mhmocap/optimizer.py:7-14``), backed by
``oracle.raster``.  Put this directory ahead of ``/root/reference`` on
``sys.path`` to run the reference in the build container.  Not PyTorch3D.
"""
__version__ = "0.0-oracle-shim"
