"""SwarmGraph: the batched, sparse replacement of the reference's GraphsTuple
(gcbfplus/utils/graph.py:47-186) -- struct of arrays over G graphs x N agents, with
receiver-grouped edge lists instead of dense padded N x N blocks (see include/gcbf_b200.h).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import TYPE_CHECKING, Optional

import torch

if TYPE_CHECKING:  # pragma: no cover
    from ..env.base import MultiAgentEnv


@dataclass
class EnvState:
    """gcbfplus/env/double_integrator.py:24-31 EnvState, batched over graphs."""
    agent: torch.Tensor          # [G, N, sd]
    goal: torch.Tensor           # [G, N, sd]
    obstacle: object             # Obstacle container (packed device tensor [G, O, w])

    @property
    def n_agent(self) -> int:
        return self.agent.shape[-2]


@dataclass
class SwarmGraph:
    env: "MultiAgentEnv"
    agent: torch.Tensor          # [G, N, sd] fp32
    goal: torch.Tensor           # [G, N, sd]
    obstacle: object
    hits: torch.Tensor           # [G, N, R, pd]  LiDAR hit nodes (position part; rest of the state is 0)
    row_start: torch.Tensor      # [G*N] int32
    row_deg: torch.Tensor        # [G*N] int32
    edge_recv: torch.Tensor      # [edge_cap] int32
    edge_src: torch.Tensor       # [edge_cap] int32
    counters: torch.Tensor       # [4] int32: n_edges, overflow flag
    clip_all: bool = False       # True for graphs made by add_edge_feats / forward_graph

    # ---- reference-compatible views ----
    @property
    def n_graphs(self) -> int:
        return int(self.agent.shape[0])

    @property
    def env_states(self) -> EnvState:
        return EnvState(self.agent, self.goal, self.obstacle)

    @property
    def n_edge(self) -> int:
        """Number of real edges (device sync)."""
        return int(self.counters[0].item())

    @property
    def states(self) -> torch.Tensor:
        """[G, 2N + N R, sd]: [agents | goals | hit nodes] like the reference (no pad node)."""
        G, N, sd = self.agent.shape
        R, pd = self.hits.shape[2], self.hits.shape[3]
        hit = torch.zeros(G, N * R, sd, device=self.agent.device, dtype=self.agent.dtype)
        hit[..., :pd] = self.hits.reshape(G, N * R, pd)
        return torch.cat([self.agent, self.goal, hit], dim=1)

    def type_states(self, type_idx: int, n_type: Optional[int] = None) -> torch.Tensor:
        """gcbfplus/utils/graph.py:126-139: 0 agents, 1 goals, 2 hit nodes."""
        if type_idx == 0:
            return self.agent
        if type_idx == 1:
            return self.goal
        G, N, sd = self.agent.shape
        return self.states[:, 2 * N:]

    def check_overflow(self) -> None:
        if int(self.counters[1].item()) != 0:
            raise RuntimeError(
                f"edge capacity overflow: {int(self.counters[0].item())} edges needed > edge_cap="
                f"{self.edge_recv.numel()}; raise env.edge_cap_per_agent")

    def _replace(self, **kw) -> "SwarmGraph":
        return replace(self, **kw)

    def index(self, idx) -> "SwarmGraph":
        """Sub-batch of graphs (re-building topology lazily is the caller's job)."""
        raise NotImplementedError("slice the state tensors and call env.get_graph / env.topology")
