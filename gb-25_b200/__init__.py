"""gb25_b200 — host-side mirror of the GordonBell25 time-stepping interface over libgb25cuda.

Only what the hot path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI,
``include/gb25cuda.h``), the ctypes binding, and the Python mirror of the reference's
constructors / ``first_time_step!`` / ``time_step!`` / ``loop!`` / workloads / ``compare_states``
(/root/reference/src/*.jl).  There is no CPU fallback: every compute entry point raises if the
CUDA library is missing or no device is present.
"""
from . import grids, splitexplicit, sharding  # noqa: F401
from .config import PhysicsConfig  # noqa: F401
