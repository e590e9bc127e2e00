"""B200-native batched implementation of LongTermPlanner's planning hot path.

    from longtermplanner_b200 import LongTermPlanner, Trajectory

``LongTermPlanner`` mirrors the reference class and adds ``planTrajectories``; it drives the
hand-written sm_100a kernels in ``csrc/`` through the C ABI of ``include/ltp_b200.h``.
There is no CPU fallback: importing the planner without the built CUDA library raises.
``EnvBatch`` keeps a vectorised environment's problem-major state on the device and replans it
(zero-copy torch tensors / any DLPack producer in, torch tensors out).
``longtermplanner_b200.workloads`` (synthetic inputs) is plain numpy and always importable.
"""
from . import workloads  # noqa: F401

__all__ = ["LongTermPlanner", "Trajectory", "BatchSolution", "BatchTrajectories", "EnvBatch", "workloads"]


def __getattr__(name):
    if name in ("LongTermPlanner", "Trajectory", "BatchSolution", "BatchTrajectories"):
        from . import planner  # raises ImportError if lib/libltp_b200.so has not been built
        return getattr(planner, name)
    if name == "EnvBatch":
        from . import envs
        return envs.EnvBatch
    raise AttributeError(name)
