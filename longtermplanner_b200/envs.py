"""Receding-horizon planning for a batch of environments whose state lives problem-major on the
GPU -- the reference's stated use (README.md:10-13: an RL agent emits sparse goals, the planner
turns each into a jerk-limited trajectory and is re-invoked before the previous goal is reached).

    envs = EnvBatch(planner, q, v, a)          # [n, dof] float64 CUDA tensors (or DLPack producers)
    traj = envs.replan(goals, horizon=2001)    # traj.q / v / a / j: (horizon, n, dof)
    envs.advance(10)                           # the state 10 samples on becomes the new start state

Everything stays on the device and on the caller's current stream; the only kernels outside
libltp_b200.so are the caller's own.
"""
from __future__ import annotations

from typing import Optional

import torch

from .planner import BatchSolution, BatchTrajectories, LongTermPlanner, as_tensor


class EnvBatch:
    def __init__(self, planner: LongTermPlanner, q, v, a):
        self.planner = planner
        q, v, a = (as_tensor(x) for x in (q, v, a))
        n, dof = q.shape
        if dof != planner.dof_:
            raise ValueError(f"state has {dof} joints, the planner {planner.dof_}")
        self.n, self.dof = n, dof
        # joint-major working copies (what the kernels read and ltp_advance_batch writes)
        self._state = [planner.transpose(x.contiguous()) for x in (q, v, a)]
        self._goal = torch.empty(dof, n, dtype=torch.float64, device=q.device)
        self.solution: Optional[BatchSolution] = None
        self.trajectories: Optional[BatchTrajectories] = None
        planner.reserve(n)

    def state(self):
        """(q, v, a) problem-major [n, dof] (fresh tensors)"""
        return tuple(self.planner.transpose(x) for x in self._state)

    def replan(self, q_goal, horizon: int) -> BatchTrajectories:
        """solve + sample from the current state towards q_goal ([n, dof]); fixed horizon, time-major"""
        self.planner.transpose(as_tensor(q_goal).contiguous(), out=self._goal)
        self.solution = self.planner.solve(self._goal, *self._state, out=self.solution)
        self.trajectories = self.planner.sample(*self._state, self.solution, horizon=horizon,
                                                out=self.trajectories if self.trajectories is not None
                                                and self.trajectories.stride == horizon else None)
        return self.trajectories

    def advance(self, tick: int) -> None:
        """the state at sample index `tick` of the current plans becomes the start state; environments
        whose last plan was rejected keep theirs"""
        if self.trajectories is None:
            raise RuntimeError("advance() before replan()")
        self.planner.advance(self.trajectories, tick, *self._state, valid=self.solution.reached)
