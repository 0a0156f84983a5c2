"""CPU oracle for the rt3d geometric path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``rt-3d-reconstruction_b200/``) never does, and fails loudly when its CUDA
library is missing rather than falling back to anything in here.

Layout
  rt3d_oracle.c   C restatement of every stage (each function cites the reference
                  file:line it follows); built by ``oracle/Makefile`` into
                  ``librt3d_oracle.so``.
  oracle.py       ctypes/numpy front-end + the Python restatement of the fusion
                  control flow (``2cam/vision_pipeline_utils.py:88-165``).
  ref_import.py   imports the UNMODIFIED reference module from /root/reference with
                  stub modules for its absent dependencies (build container only:
                  /root/reference does not exist on the GPU box).

Parity status: B0/B1/T1/C1/W0/V1/S1/centroid/F1-control-flow are pinned against the
reference's own source run on torch-CPU (tests/golden + tests/test_oracle_vs_reference.py).
O1 (Open3D remove_statistical_outlier) and V2 (Open3D tensor voxel_down_sample) are
"parity unpinned": open3d==0.18.0 is an un-vendored dependency that cannot be installed
here, so they restate its published algorithm and are cross-checked against an
independent scipy.cKDTree restatement only.
"""
from .oracle import *  # noqa: F401,F403
