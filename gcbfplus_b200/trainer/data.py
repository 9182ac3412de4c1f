"""Rollout container -- gcbfplus/trainer/data.py:8-31, compact: the dense per-step
GraphsTuples of the reference are replaced by the states they are a function of
(agent states, goals, obstacles, LiDAR hit nodes); topology is rebuilt on demand."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class Rollout:
    agent: torch.Tensor        # [b, T+1, N, sd]  agent[:, t] = graph_t, agent[:, t+1] = next_graph_t
    goal: torch.Tensor         # [b, N, sd]
    hits: torch.Tensor         # [b, T+1, N, R, pd]
    obstacle: object           # batched over b
    actions: torch.Tensor      # [b, T, N, nu]
    rewards: torch.Tensor      # [b, T]
    costs: torch.Tensor        # [b, T]
    dones: torch.Tensor        # [b, T]
    log_pis: Optional[torch.Tensor] = None
    n_edges: Optional[torch.Tensor] = None   # [T+1] real edge count per step (all envs)

    @property
    def length(self) -> int:
        return int(self.rewards.shape[0])

    @property
    def time_horizon(self) -> int:
        return int(self.rewards.shape[1])

    @property
    def num_agents(self) -> int:
        return int(self.agent.shape[2])

    @property
    def n_data(self) -> int:
        return self.length * self.time_horizon
