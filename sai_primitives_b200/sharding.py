"""Host-side partitioning for the batched path.

Robots are independent, so the only multi-GPU logic is bookkeeping:
  * shard_range: contiguous block of robot indices per rank (one process per GPU, no collective);
  * ModelGroups: a mixed-DoF batch is split by robot model into homogeneous groups, one handle (and one
    kernel specialisation) per group, and results are scattered back to the caller's robot order
    (BASELINE config 4).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world_size: int):
    """[lo, hi) of the robots owned by `rank`: robot i -> rank floor(i * world_size / n_total), i.e. contiguous,
    ordered blocks whose sizes differ by at most one.  Same arithmetic as the C entry point osc_shard_range."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    n_total, world_size = int(n_total), int(world_size)
    lo = -((-rank * n_total) // world_size)
    hi = -((-(rank + 1) * n_total) // world_size)
    return lo, hi


def shard_of(index: int, n_total: int, world_size: int) -> int:
    """rank that owns robot `index` under shard_range"""
    return (int(index) * int(world_size)) // int(n_total)


class ModelGroups:
    """Groups the robots of a mixed batch by model name.  `build(name, n)` is called once per distinct model
    and must return an object with .set_state(q, dq) and .cycle() -> tau [n, dof] (e.g. a small wrapper around
    BatchedRobot + RobotController); rows of the per-robot inputs are routed to their group and the torques
    come back in the caller's order, zero-padded to the largest dof."""

    def __init__(self, model_of_robot, build):
        self.model_of_robot = list(model_of_robot)
        names = sorted(set(self.model_of_robot))
        self.index = {nm: np.array([i for i, m in enumerate(self.model_of_robot) if m == nm], dtype=np.int64) for nm in names}
        self.groups = {nm: build(nm, len(self.index[nm])) for nm in names}

    def n_robots(self):
        return len(self.model_of_robot)

    def set_state(self, q_rows, dq_rows):
        """q_rows[i], dq_rows[i]: 1-D arrays of robot i's dof"""
        for nm, idx in self.index.items():
            q = np.array([q_rows[i] for i in idx], dtype=np.float64)
            dq = np.array([dq_rows[i] for i in idx], dtype=np.float64)
            self.groups[nm].set_state(q, dq)

    def cycle(self):
        out = [None] * self.n_robots()
        for nm, idx in self.index.items():
            tau = self.groups[nm].cycle()
            for row, i in enumerate(idx):
                out[i] = tau[row]
        return out
