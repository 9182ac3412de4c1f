"""MultiAgentController ABC -- gcbfplus/algo/base.py:10-68."""
from abc import ABC, abstractmethod
from typing import Optional, Tuple

import torch

from ..env.base import MultiAgentEnv
from ..utils.graph import SwarmGraph


class MultiAgentController(ABC):

    def __init__(self, env: MultiAgentEnv, node_dim: int, edge_dim: int, action_dim: int, n_agents: int):
        self._env = env
        self._node_dim = node_dim
        self._edge_dim = edge_dim
        self._action_dim = action_dim
        self._n_agents = n_agents

    @property
    def node_dim(self) -> int:
        return self._node_dim

    @property
    def edge_dim(self) -> int:
        return self._edge_dim

    @property
    def action_dim(self) -> int:
        return self._action_dim

    @property
    def n_agents(self) -> int:
        return self._n_agents

    @property
    @abstractmethod
    def config(self) -> dict:
        pass

    @property
    @abstractmethod
    def actor_params(self):
        pass

    @abstractmethod
    def act(self, graph: SwarmGraph, params=None) -> torch.Tensor:
        pass

    @abstractmethod
    def step(self, graph: SwarmGraph, key, params=None) -> Tuple[torch.Tensor, torch.Tensor]:
        pass

    @abstractmethod
    def update(self, rollout, step: int) -> dict:
        pass

    @abstractmethod
    def save(self, save_dir: str, step: int):
        pass

    @abstractmethod
    def load(self, load_dir: str, step: int):
        pass
