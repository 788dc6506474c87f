"""B200-native drop-in for the reference's per-frame velocity-field solve.

Mirrors ``utils/compute_optical_flow.py`` and ``utils/find_singularity_point.py``
of SEU-dynamical-models/Manifold-based-optical-flow-method:

    from manifold_based_optical_flow_method_b200 import compute_optical_flow, find_singularity_point

replaces the reference's ``from utils import compute_optical_flow, find_singularity_point``
(S3_compute_v_and_detection_singularity.py:11).  All numerics run in hand-written
sm_100a CUDA kernels behind the C-ABI library ``csrc/libmof_b200.so``
(include/mof_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
