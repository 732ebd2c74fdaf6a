"""ungar_b200 — B200-native batched derivative evaluation for Ungar NMPC problems.

The product path is: this package -> ctypes -> C ABI (include/ungar_b200.h) -> hand-written sm_100a CUDA
kernels (ungar_b200/csrc).  It fails loudly when the CUDA library or a CUDA device is missing.
"""
from .function import Function, Model, SoftSQPOptimizer  # noqa: F401
from .workloads import MODEL_IDS, MODEL_NAMES, QUADROTOR, QUADRUPED, RC_CAR  # noqa: F401

# RelaxedPolyBarrierFunction (stiffness, epsilon) each reference example passes to SoftSQPOptimizer
# (quadrotor.example.cpp:370 defaults of soft_sqp.hpp:44-50; rc_car.example.cpp:363; quadruped.example.cpp:444).
EXAMPLE_BARRIER = {QUADROTOR: (100.0, 2e-5), RC_CAR: (100.0, 1e-2), QUADRUPED: (1.0, 1.0)}
