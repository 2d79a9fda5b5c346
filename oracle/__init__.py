"""CPU oracle for the scene-aware multi-human SMPL optimisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
CPU baseline being timed -- never as the thing shipped.  The product path
(``scene-aware-3d-multi-human_b200``) fails loudly when ``libmhopt.so`` is missing.

What is here
------------
* ``synth``      synthetic SMPL-shaped body model (V=6890, F=13776) and synthetic
                 multi-person sequences -- the licensed
                 ``SMPL_NEUTRAL.pkl`` and MuPoTS data are not available.
* ``refmath``    torch-CPU fp32 restatement of the reference-owned arithmetic
                 (``mhmocap/smpl.py``, ``transforms.py``, ``losses.py``,
                 ``morphology.py``, ``one_euro_filter.py``), each function citing
                 the reference file:line it follows.
* ``raster``     torch-CPU restatement of the PyTorch3D mesh rasteriser /
                 soft-silhouette shader semantics the reference calls
                 (``optimizer.py:210-232, 428-430, 447-448``).
* ``fit_ref``    restatement of ``SMPLDepthSequenceOptimizer`` (``optimizer.py``)
                 on torch autograd + ``torch.optim`` -- the "port" timed as the CPU
                 baseline and the teacher-forced parity checker.
* ``pytorch3d_shim``  stand-in ``pytorch3d`` package (backed by ``raster``) that
                 lets the UNMODIFIED reference optimiser run in the build
                 container so that ``refmath``/``fit_ref`` can be pinned against it
                 (``tests/golden/make_golden.py``).

Parity status
-------------
* Reference-owned arithmetic (SMPL, projection, loss algebra, priors, temporal
  terms, RMSprop/Adam): PINNED -- checked against the reference's own functions
  imported from ``/root/reference`` by the generators ``tests/golden/make_golden.py``,
  ``make_eval_golden.py`` and ``make_init_w17_golden.py`` (they run only where the
  reference is mounted) and, at test time, against the committed golden vectors
  they wrote to ``tests/golden/`` (``tests/test_oracle_golden.py``).
* Rasteriser / silhouette shader: PARITY UNPINNED.  The algorithm lives in the
  third-party dependency ``pytorch3d`` (conda channel ``pytorch3d``, unpinned in
  the reference's ``environment.yml:13``; ~0.5-0.7), whose source is not under
  ``/root/reference`` and which is not installable here.  ``raster`` restates its
  published naive-rasterisation semantics; the CUDA
  kernels are checked against that restatement.
"""
