"""Batched closed-loop rollout -- replaces jit(vmap(rollout)) of gcbfplus/trainer/utils.py:25-55
and trainer/trainer.py:81-87.  One CUDA graph holds the whole T-step loop:
per step {actor GNN forward -> act + clip + Euler + reward/cost -> LiDAR + neighbour lists},
3 C-ABI calls / 14 kernel launches, no host sync inside the loop.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from ..algo.params import NetParams
from ..utils.graph import SwarmGraph
from .data import Rollout


class RolloutEngine:
    def __init__(self, env, n_envs: int, T: Optional[int] = None, n_obs: Optional[int] = None,
                 use_cuda_graph: bool = True, policy: str = "actor"):
        """policy: 'actor' (a = 2 pi + u_ref, algo.step) or 'u_ref' (test.py --u-ref)."""
        self.env = env
        self.E = n_envs
        self.T = T or env.max_episode_steps
        self.O = env.params["n_obs"] if n_obs is None else n_obs
        self.policy = policy
        self.use_cuda_graph = use_cuda_graph
        dev = env.device
        E, T, N = self.E, self.T, env.num_agents
        sd, nu, R, pd = env.state_dim, env.action_dim, env.n_hits, env.pos_dim
        f32, i32 = torch.float32, torch.int32
        self.agent = torch.zeros(T + 1, E, N, sd, dtype=f32, device=dev)
        self.hits = torch.zeros(T + 1, E, N, R, pd, dtype=f32, device=dev)
        self.goal = torch.zeros(E, N, sd, dtype=f32, device=dev)
        self.obs_w = 16 if pd == 2 else 4
        self.obstacles = torch.zeros(E, max(self.O, 1), self.obs_w, dtype=f32, device=dev)
        self.actions = torch.zeros(T, E, N, nu, dtype=f32, device=dev)
        self.rewards = torch.zeros(T, E, dtype=f32, device=dev)
        self.costs = torch.zeros(T, E, dtype=f32, device=dev)
        self.pi = torch.zeros(E, N, nu, dtype=f32, device=dev)
        self.desc = env.desc(E, self.O)
        cap = self.desc.edge_cap
        self.row_start = torch.zeros(E * N, dtype=i32, device=dev)
        self.row_deg = torch.zeros(E * N, dtype=i32, device=dev)
        self.edge_recv = torch.zeros(cap, dtype=i32, device=dev)
        self.edge_src = torch.zeros(cap, dtype=i32, device=dev)
        self.counters = torch.zeros(T + 1, 4, dtype=i32, device=dev)
        n_ws = env.lib.gcbf_gnn_workspace_floats(C.byref(self.desc), nu)
        self.ws = torch.empty(int(n_ws), dtype=f32, device=dev)
        self.params_buf = torch.zeros(_lib.param_count(env.edge_dim, nu), dtype=f32, device=dev)
        # folded inference weights (gcbf_prepare_infer), rebuilt by set_params()
        self.infer_blob = torch.zeros(int(env.lib.gcbf_infer_count(env.edge_dim, nu)), dtype=f32, device=dev)
        self.use_tc = 1 if _lib.USE_TC else 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_run = 0
        self._obstacle_obj = None

    # ------------------------------------------------------------------ one env step (enqueue only)
    def _build(self, t: int, stream: int) -> None:
        env, d = self.env, self.desc
        rc = env.lib.gcbf_graph_build(C.byref(d), self.agent[t].data_ptr(),
                                      self.obstacles.data_ptr() if self.O > 0 else None, env.ray_table.data_ptr(),
                                      self.hits[t].data_ptr(), self.row_start.data_ptr(), self.row_deg.data_ptr(),
                                      self.edge_recv.data_ptr(), self.edge_src.data_ptr(),
                                      self.counters[t].data_ptr(), 1, stream)
        _lib.check(rc, "gcbf_graph_build")

    def _step(self, t: int, stream: int) -> None:
        env, d = self.env, self.desc
        nu = env.action_dim
        mode = 2
        if self.policy == "actor":
            rc = env.lib.gcbf_gnn_infer(C.byref(d), _lib.NET_ACTOR, nu, self.params_buf.data_ptr(),
                                        self.infer_blob.data_ptr(), self.use_tc, self.agent[t].data_ptr(),
                                        self.goal.data_ptr(), self.hits[t].data_ptr(),
                                          self.row_start.data_ptr(), self.row_deg.data_ptr(),
                                          self.edge_recv.data_ptr(), self.edge_src.data_ptr(),
                                          self.counters[t].data_ptr(), 0, self.pi.data_ptr(), self.ws.data_ptr(),
                                          self.ws.numel(), stream)
            _lib.check(rc, "gcbf_gnn_infer")
            mode = 0
        rc = env.lib.gcbf_env_step(C.byref(d), self.agent[t].data_ptr(), self.goal.data_ptr(),
                                   self.obstacles.data_ptr() if self.O > 0 else None, self.pi.data_ptr(),
                                   self.row_start.data_ptr(), self.row_deg.data_ptr(), self.edge_src.data_ptr(),
                                   self.actions[t].data_ptr(), self.agent[t + 1].data_ptr(),
                                   self.rewards[t].data_ptr(), self.costs[t].data_ptr(), mode, stream)
        _lib.check(rc, "gcbf_env_step")
        self._build(t + 1, stream)

    def _enqueue_all(self, stream: int) -> None:
        self._build(0, stream)
        for t in range(self.T):
            self._step(t, stream)

    # ------------------------------------------------------------------ public
    def set_initial(self, agent0: torch.Tensor, goal: torch.Tensor, obstacle) -> None:
        """Initial conditions: [E,N,sd] x2 (device or pinned host) + obstacle container."""
        self.agent[0].copy_(agent0.reshape(self.agent[0].shape), non_blocking=True)
        self.goal.copy_(goal.reshape(self.goal.shape), non_blocking=True)
        if self.O > 0:
            packed = obstacle.packed if hasattr(obstacle, "packed") else obstacle
            self.obstacles.copy_(packed.reshape(self.obstacles.shape), non_blocking=True)
        self._obstacle_obj = obstacle

    def set_params(self, params: NetParams) -> None:
        self.params_buf.copy_(params.flat, non_blocking=True)
        env = self.env
        _lib.check(env.lib.gcbf_prepare_infer(env.edge_dim, env.action_dim, _lib.ptr(self.params_buf),
                                              _lib.ptr(self.infer_blob), env._stream()), "gcbf_prepare_infer")

    def run(self, check: bool = True) -> None:
        """Run the T-step rollout from the current initial conditions (async)."""
        dev = self.env.device
        self.counters.zero_()
        if self.use_cuda_graph:
            if self._graph is None:
                lib = self.env.lib
                # warm-up outside capture (lazy module load, function attributes)
                self._build(0, torch.cuda.current_stream(dev).cuda_stream)
                self._step(0, torch.cuda.current_stream(dev).cuda_stream)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                n0 = lib.gcbf_launch_count()
                with torch.cuda.graph(g):
                    self._enqueue_all(torch.cuda.current_stream(dev).cuda_stream)
                self.launches_per_run = int(lib.gcbf_launch_count() - n0)
                self._graph = g
            self._graph.replay()
        else:
            n0 = self.env.lib.gcbf_launch_count()
            self._enqueue_all(torch.cuda.current_stream(dev).cuda_stream)
            self.launches_per_run = int(self.env.lib.gcbf_launch_count() - n0)
        if check:
            self.check_overflow()

    def check_overflow(self) -> None:
        c = self.counters.cpu()
        if int(c[:, 1].max()) != 0:
            raise RuntimeError(f"edge capacity overflow during rollout: up to {int(c[:, 0].max())} edges > edge_cap="
                               f"{self.desc.edge_cap}; raise env.edge_cap_per_agent")

    def result(self) -> Rollout:
        """trainer/data.py Rollout in the reference's (b, T) order (views/transposes of the record)."""
        dones = torch.zeros(self.E, self.T, dtype=torch.bool, device=self.env.device)
        return Rollout(agent=self.agent.transpose(0, 1), goal=self.goal, hits=self.hits.transpose(0, 1),
                       obstacle=self._obstacle_obj, actions=self.actions.transpose(0, 1),
                       rewards=self.rewards.transpose(0, 1), costs=self.costs.transpose(0, 1), dones=dones,
                       log_pis=None, n_edges=self.counters[:, 0])

    def final_graph(self) -> SwarmGraph:
        return SwarmGraph(self.env, self.agent[self.T], self.goal, self._obstacle_obj, self.hits[self.T],
                          self.row_start, self.row_deg, self.edge_recv, self.edge_src, self.counters[self.T])
