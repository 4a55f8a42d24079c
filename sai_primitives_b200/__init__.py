"""B200-native batched operational-space controller (drop-in for the hot path of
manips-sai-org/sai-primitives): C ABI in include/sai_b200_osc.h, kernels in csrc/,
Python host mirror in batched.py.  No CPU fallback: importing the host mirror loads
libsai_b200_osc.so or fails."""
from . import capi  # noqa: F401
from .batched import (  # noqa: F401
    BOUNDED_INERTIA_ESTIMATES,
    FULL_DYNAMIC_DECOUPLING,
    IMPEDANCE,
    BatchedRobot,
    BatchedSimulation,
    JointTask,
    MotionForceTask,
    RobotController,
    registerUrdf,
)
